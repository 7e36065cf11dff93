"""Data-parallel training: one process per GPU, ONE flat gradient bucket, one all-reduce per step.

The reference trains under nn.DataParallel (train.py:93-96: one process, a thread per GPU, parameters re-broadcast and
outputs gathered every forward).  The model has 343 685 parameters = 1.37 MB of fp32 gradients: bucketing
and overlap machinery buys nothing at that size, launch count does.  So:

  * `FlatBucketDDP(model)` broadcasts rank 0's parameters and buffers once, then makes every parameter's `.grad` a
    view into one contiguous fp32 buffer -- backward writes the bucket in place, no flatten copy;
  * `reduce_gradients()` is a single `all_reduce(SUM)` over that buffer (NCCL over NVLink on GPUs, gloo in the CPU
    tests) followed by one in-place scale; with `grad_dtype=torch.bfloat16` the bucket is rounded to bf16 for the
    collective (0.69 MB on the wire; the cross-rank SUM itself then runs in bf16 inside NCCL / gloo -- about 3 significant
    digits per gradient element, fine under clip_grad_norm_ + Adam) while backward's accumulation into the bucket, the
    scale, the clipping and the optimizer stay fp32 (BASELINE config 4);
  * parameters that take no part in the forward (feature_net.inner3, net.py:25) simply keep a zero gradient: no
    unused-parameter search.  With Adam and weight_decay = 0 (train.py:98 defaults) a zero gradient leaves the
    parameter where `grad is None` would.

`train_step` mirrors train.py:194-215 (zero_grad, forward, full_loss, backward, [all-reduce], clip_grad_norm_ 2.0,
optimizer.step).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


class FlatBucketDDP(nn.Module):
    def __init__(self, module: nn.Module, process_group=None, grad_dtype: Optional[torch.dtype] = None,
                 broadcast_from_rank0: bool = True):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.grad_dtype = grad_dtype
        self._params = [p for p in module.parameters() if p.requires_grad]
        if not self._params:
            raise ValueError("FlatBucketDDP: the module has no trainable parameter")
        dev, dt = self._params[0].device, self._params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self._params):
            raise ValueError("FlatBucketDDP: all trainable parameters must share one device and dtype")
        self._bucket = torch.zeros(sum(p.numel() for p in self._params), device=dev, dtype=dt)
        self._wire = torch.empty_like(self._bucket, dtype=grad_dtype) if grad_dtype not in (None, dt) else None
        self._attach()
        if broadcast_from_rank0 and self._active():
            self._broadcast_state()

    # -- plumbing --------------------------------------------------------------------------------------------
    def _active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1

    def _attach(self) -> None:
        off = 0
        for p in self._params:
            n = p.numel()
            p.grad = self._bucket[off:off + n].view_as(p)
            off += n

    @torch.no_grad()
    def _broadcast_state(self) -> None:
        # copy_ on the parameters / buffers themselves (not on `.data`): the version counters move, so packed-weight
        # caches keyed on them (estimator._pack_key) see the new values
        tensors = list(self.module.parameters()) + list(self.module.buffers())
        for dtype in sorted({t.dtype for t in tensors}, key=str):       # same order on every rank
            group = [t for t in tensors if t.dtype == dtype]
            flat = torch.cat([t.detach().reshape(-1) for t in group])
            dist.broadcast(flat, src=0, group=self.process_group)
            off = 0
            for t in group:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()

    # -- the module interface --------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def state_dict(self, *args, **kwargs):
        """Keys carry the 'module.' prefix, like the checkpoints the reference saves from nn.DataParallel
        (train.py:153-157) -- interchangeable with them."""
        return super().state_dict(*args, **kwargs)

    def _views_intact(self) -> bool:
        lo = self._bucket.data_ptr()
        hi = lo + self._bucket.numel() * self._bucket.element_size()
        return all(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self._params)

    def zero_grad(self, set_to_none: bool = False) -> None:      # noqa: ARG002  (the views must survive)
        self._bucket.zero_()
        if not self._views_intact():          # someone replaced a .grad (e.g. optimizer.zero_grad(set_to_none=True))
            self._attach()

    @property
    def gradient_bucket(self) -> torch.Tensor:
        return self._bucket

    def reduce_gradients(self) -> None:
        """Average the gradient bucket over the ranks: one collective."""
        if not self._active():
            return
        world = dist.get_world_size(self.process_group)
        if self._wire is not None:
            self._wire.copy_(self._bucket)
            dist.all_reduce(self._wire, op=dist.ReduceOp.SUM, group=self.process_group)
            self._bucket.copy_(self._wire)
        else:
            dist.all_reduce(self._bucket, op=dist.ReduceOp.SUM, group=self.process_group)
        self._bucket.div_(world)

    def average_buffers(self) -> None:
        """Average the floating-point buffers (BatchNorm running statistics) over the ranks -- call before saving a
        checkpoint; every rank normalises with its own batch statistics during training, as under DataParallel."""
        if not self._active():
            return
        world = dist.get_world_size(self.process_group)
        with torch.no_grad():                  # in-place ops on the buffers themselves: their version counters move,
            for b in self.module.buffers():    # so cached BN-folded weight packs (estimator._pack_key) are rebuilt
                if b.is_floating_point():
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.process_group)
                    b.div_(world)


def train_step(model: nn.Module, optimizer: torch.optim.Optimizer, sample: Dict, loss_fn, regress: bool = True,
               clip_norm: float = 2.0):
    """train.py:194-215 for a sample already on the model's device.  `model` is a Pipeline in any wrapping
    (plain, FlatBucketDDP).  Returns the detached loss and the model outputs."""
    model.train()
    if isinstance(model, FlatBucketDDP):
        model.zero_grad()
    else:
        optimizer.zero_grad()
    outputs = model(sample["imgs"], sample["proj_matrices"], sample["depth_min"], sample["depth_max"])
    loss = loss_fn(outputs["depths"], outputs["depths_upsampled"], outputs["confidences"], sample["depth"], sample["mask"],
                   sample["depth_min"], sample["depth_max"], regress)
    loss.backward()
    if isinstance(model, FlatBucketDDP):
        model.reduce_gradients()
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip_norm)
    optimizer.step()
    return loss.detach(), outputs
