"""FlatBucketDDP / train_step host logic on CPU: two processes, gloo backend (the GPU path runs the same code on
nccl).  The wrapped module here is a small conv net with BatchNorm and an unused parameter (the shape of the
problem: feature_net.inner3 takes no part in the reference's forward, net.py:25) -- the wrapper is model-agnostic;
the Pipeline itself needs the CUDA library."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from itermvs_b200.ddp import FlatBucketDDP


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Conv2d(3, 8, 3, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(8)
        self.b = nn.Conv2d(8, 1, 1)
        self.unused = nn.Conv2d(8, 8, 1)

    def forward(self, x):
        return self.b(torch.relu(self.bn(self.a(x))))


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(4, 3, 8, 8, generator=g), torch.randn(4, 1, 8, 8, generator=g)


def _worker(rank, world, port, q, wire_bf16):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)                       # different initial weights per rank: the broadcast must fix that
        net = Net()
        ddp = FlatBucketDDP(net, grad_dtype=torch.bfloat16 if wire_bf16 else None)
        start = {k: v.clone() for k, v in net.state_dict().items()}
        opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=1e-2)
        x, y = _data(rank)
        for step in range(2):
            ddp.zero_grad()
            loss = ((ddp(x) - y) ** 2).mean()
            loss.backward()
            if step == 0:
                local = {k: p.grad.clone() for k, p in net.named_parameters()}
                in_bucket = all(p.grad.data_ptr() >= ddp.gradient_bucket.data_ptr() for p in net.parameters())
            ddp.reduce_gradients()
            if step == 0:
                reduced = {k: p.grad.clone() for k, p in net.named_parameters()}
            torch.nn.utils.clip_grad_norm_(net.parameters(), 2.0)
            opt.step()
            opt.zero_grad(set_to_none=True)           # a caller that drops the views: zero_grad() must re-attach
        end = {k: v.clone() for k, v in net.state_dict().items()}
        # numpy over the queue: torch tensors travel as shared-memory handles that die with this process
        q.put((rank,) + tuple({k: v.detach().numpy().copy() for k, v in d.items()} for d in (start, local, reduced, end)) + (in_bucket,))
    finally:
        dist.destroy_process_group()


def _run(wire_bf16):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, wire_bf16)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return [tuple({k: torch.from_numpy(v) for k, v in item.items()} if isinstance(item, dict) else item for item in t) for t in out]


def test_flat_bucket_allreduce_two_ranks():
    (_, start0, local0, red0, end0, inb0), (_, start1, local1, red1, end1, inb1) = _run(False)
    assert inb0 and inb1                                            # gradients were written into the bucket views
    for k in start0:
        assert torch.equal(start0[k], start1[k]), k                 # rank 0's parameters and buffers everywhere
    for k in local0:
        want = (local0[k] + local1[k]) / 2
        assert torch.allclose(red0[k], want, rtol=1e-6, atol=1e-7) and torch.equal(red0[k], red1[k]), k
    assert float(red0["unused.weight"].abs().max()) == 0.0          # no gradient, no search for unused parameters
    for k in end0:
        if "running" in k or "num_batches" in k:
            continue                                                # per-rank batch statistics, as under DataParallel
        assert torch.equal(end0[k], end1[k]), k                     # replicas stay in lock-step
    assert not torch.equal(end0["a.weight"], start0["a.weight"])
    assert torch.equal(end0["unused.weight"], start0["unused.weight"])


def test_bf16_wire_format():
    (_, _, local0, red0, _, _), (_, _, local1, red1, _, _) = _run(True)
    for k in local0:
        want = (local0[k] + local1[k]) / 2
        assert torch.equal(red0[k], red1[k]), k
        scale = max(float(local0[k].abs().max()), float(local1[k].abs().max()), 1e-12)
        assert float((red0[k] - want).abs().max()) <= 2 ** -7 * scale, k     # bf16 on the wire: 8 bits of mantissa per rank


def test_single_process_is_a_no_op():
    net = Net()
    ddp = FlatBucketDDP(net)
    x, y = _data(0)
    ((ddp(x) - y) ** 2).mean().backward()
    before = ddp.gradient_bucket.clone()
    ddp.reduce_gradients()
    assert torch.equal(before, ddp.gradient_bucket) and float(before.abs().sum()) > 0
    assert all(k.startswith("module.") for k in ddp.state_dict())   # DataParallel-style checkpoint keys (train.py:153-157)
