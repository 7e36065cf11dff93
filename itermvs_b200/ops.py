"""Tensor-level wrappers over the C ABI (include/itermvs_b200.h) + the reference's module-level
functions of models/module.py re-exposed with the same names and argument meaning:

    differentiable_warping(src_fea, src_proj, ref_proj, depth_samples, return_mask=False)
    upsample(x, upsample_weight, scale=4)            (kept in torch: only used by the training path)
    depth_normalization / depth_unnormalization

Everything here runs on the CUDA library; there is no CPU fallback (a CPU tensor raises).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib

Tensor = torch.Tensor


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"itermvs_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"itermvs_b200: {name} must be float32, got {t.dtype}")
    return t.contiguous()


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


class NanFlag:
    """Deferred version of the reference's `assert not isnan(proj)` (module.py:83,87): kernels raise a
    device flag; `raise_if_set()` is called where the host synchronises anyway."""

    def __init__(self, device):
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)

    def ptr(self):
        return self.flag.data_ptr()

    def raise_if_set(self):
        if int(self.flag.item()) != 0:
            self.flag.zero_()
            raise AssertionError("nan in proj")          # same message as module.py:87


# ------------------------------------------------------------------------------------------------
def compose_projections(proj: Tensor, nan_flag: Optional[NanFlag] = None) -> Tensor:
    """proj [B,V,4,4] (view 0 = reference) -> [B,V-1,12] rot|trans of src @ inverse(ref)."""
    proj = _chk(proj, "proj")
    b, v = proj.shape[:2]
    out = torch.empty(b, v - 1, 12, device=proj.device, dtype=torch.float32)
    _lib.check(_lib.lib().imvs_compose_projections(proj.data_ptr(), b, v, out.data_ptr(),
                                                   nan_flag.ptr() if nan_flag else None, _stream()), "compose_projections")
    return out


def _warp_forward(src_fea: Tensor, src_proj: Tensor, ref_proj: Tensor, depth_samples: Tensor):
    b, c, h1, w1 = src_fea.shape
    _, d, h, w = depth_samples.shape
    out = torch.empty(b, c, d, h, w, device=src_fea.device, dtype=torch.float32)
    rt = torch.empty(b, 12, device=src_fea.device, dtype=torch.float32)
    flag = NanFlag(src_fea.device)
    _lib.check(_lib.lib().imvs_differentiable_warping(src_fea.data_ptr(), src_proj.data_ptr(), ref_proj.data_ptr(),
                                                      depth_samples.data_ptr(), out.data_ptr(), b, c, h1, w1, d, h, w,
                                                      rt.data_ptr(), flag.ptr(), _stream()), "differentiable_warping")
    flag.raise_if_set()     # the reference asserts (and therefore synchronises) on every call as well
    return out, rt


class _WarpFn(torch.autograd.Function):
    """differentiable_warping with its CUDA backward.  Only src_fea receives a gradient: the reference builds the
    sampling grid under torch.no_grad() (module.py:77), so projections and depth samples are constants."""

    @staticmethod
    def forward(ctx, src_fea, src_proj, ref_proj, depth_samples):
        out, _ = _warp_forward(src_fea, src_proj, ref_proj, depth_samples)
        ctx.save_for_backward(src_proj, ref_proj, depth_samples)
        ctx.fea_shape = tuple(src_fea.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        src_proj, ref_proj, depth_samples = ctx.saved_tensors
        b, c, h1, w1 = ctx.fea_shape
        _, d, h, w = depth_samples.shape
        grad_out = _chk(grad_out.float(), "grad_out")
        grad_fea = torch.empty(b, c, h1, w1, device=grad_out.device, dtype=torch.float32)
        rt = torch.empty(b, 12, device=grad_out.device, dtype=torch.float32)
        _lib.check(_lib.lib().imvs_differentiable_warping_backward(
            grad_out.data_ptr(), src_proj.data_ptr(), ref_proj.data_ptr(), depth_samples.data_ptr(), grad_fea.data_ptr(),
            b, c, h1, w1, d, h, w, rt.data_ptr(), None, _stream()), "differentiable_warping_backward")
        return grad_fea, None, None, None


def differentiable_warping(src_fea: Tensor, src_proj: Tensor, ref_proj: Tensor, depth_samples: Tensor,
                           return_mask: bool = False):
    """Drop-in for reference models/module.py:68.  src_fea [B,C,H1,W1], projections [B,4,4],
    depth_samples [B,D,H,W] -> [B,C,D,H,W] (and the validity mask when return_mask=True).
    Differentiable with respect to src_fea (CUDA backward), like the reference."""
    src_fea = _chk(src_fea, "src_fea")
    depth_samples = _chk(depth_samples.detach(), "depth_samples")
    src_proj = _chk(src_proj.detach().float(), "src_proj")
    ref_proj = _chk(ref_proj.detach().float(), "ref_proj")
    b, c, h1, w1 = src_fea.shape
    _, d, h, w = depth_samples.shape
    if torch.is_grad_enabled() and src_fea.requires_grad:
        out = _WarpFn.apply(src_fea, src_proj, ref_proj, depth_samples)
        rt = compose_projections(torch.stack([ref_proj, src_proj], dim=1))[:, 0] if return_mask else None
    else:
        out, rt = _warp_forward(src_fea, src_proj, ref_proj, depth_samples)
    if not return_mask:
        return out
    # module.py:104-111 (no caller in the reference passes return_mask=True; small torch epilogue)
    rot, trans = rt[:, :9].view(b, 3, 3), rt[:, 9:].view(b, 3, 1)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=out.device),
                            torch.arange(w, dtype=torch.float32, device=out.device), indexing="ij")
    xyz = torch.stack((xs.reshape(-1) * (w1 / w), ys.reshape(-1) * (h1 / h), torch.ones(h * w, device=out.device)))
    p = torch.matmul(rot, xyz.unsqueeze(0)).unsqueeze(2) * depth_samples.view(b, 1, d, h * w) + trans.view(b, 3, 1, 1)
    valid = p[:, 2:] > 1e-2
    px = torch.where(valid, p[:, 0:1], torch.full_like(p[:, 0:1], float(w))) / torch.where(valid, p[:, 2:3], torch.ones_like(p[:, 2:3]))
    py = torch.where(valid, p[:, 1:2], torch.full_like(p[:, 1:2], float(h))) / torch.where(valid, p[:, 2:3], torch.ones_like(p[:, 2:3]))
    valid = valid & (px >= 0) & (px < w) & (py >= 0) & (py < h)
    return out, valid.view(b, d, h, w)


def nchw_to_nhwc(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    x = _chk(x, "x")
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty(n, h, w, c, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().imvs_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), n, c, h, w, _stream()), "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x: Tensor) -> Tensor:
    x = _chk(x, "x")
    n, h, w, c = x.shape
    out = torch.empty(n, c, h, w, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().imvs_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), n, c, h, w, _stream()), "nhwc_to_nchw")
    return out


# ------------------------------------------------------------------------------------------------
def upsample(x: Tensor, upsample_weight: Tensor, scale: int = 4) -> Tensor:
    """Reference models/module.py:127 (convex combination upsampling) for callers that hold the
    materialised weight tensor (the training path).  The inference path never builds that tensor:
    see imvs_upsample_outputs."""
    batch, _, height, width = x.shape
    xp = F.pad(x, (1, 1, 1, 1), mode="replicate")
    nb = torch.stack([xp[:, :, ky:ky + height, kx:kx + width] for ky in range(3) for kx in range(3)], dim=2)
    up = (nb.view(batch, -1, 9, 1, 1, height, width) * upsample_weight).sum(dim=2)
    return up.permute(0, 1, 4, 2, 5, 3).reshape(batch, -1, scale * height, scale * width)


def depth_normalization(depth: Tensor, inverse_depth_min: Tensor, inverse_depth_max: Tensor) -> Tensor:
    """module.py:142-146."""
    return (1.0 / (depth + 1e-5) - inverse_depth_max) / (inverse_depth_min - inverse_depth_max)


def depth_unnormalization(normalized_depth: Tensor, inverse_depth_min: Tensor, inverse_depth_max: Tensor) -> Tensor:
    """module.py:148-152."""
    return 1.0 / (inverse_depth_max + normalized_depth * (inverse_depth_min - inverse_depth_max))
