"""Pack parameters (reference state_dict names) into the kernel layouts of include/itermvs_b200.h.

Tensor-core convolutions take weights as [tap][CinP][CoutP] (channels zero-padded to multiples of
8), split into hi = TF32(w) and lo = TF32(w - hi) (round-to-nearest, ties away from zero -- the
same rounding as `cvt.rna.tf32.f32`), so the kernels need no conversion on the weight operand and
the 3-pass mode recovers fp32-grade products.  FeatureNet's BatchNorm (eval mode, running
statistics, eps 1e-5) is folded into the preceding convolution.

Packed tensors are plain device tensors owned by the Packed* objects; `.struct` is the C struct
pointing at them.  Repack whenever parameters change -- the modules key their cache on the
parameters' version counters.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from . import _lib

Tensor = torch.Tensor


def round_tf32(x: Tensor) -> Tensor:
    """cvt.rna.tf32.f32: keep 10 mantissa bits, round to nearest, ties away from zero."""
    bits = x.contiguous().view(torch.int32)
    r = (bits + 0x1000) & ~0x1FFF
    # (two's-complement add on the sign-magnitude pattern raises the magnitude for both signs)
    return r.view(torch.float32)


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def split_tf32(w: Tensor) -> Tuple[Tensor, Tensor]:
    """(TF32-rounded copy for the 1-pass mode, plain fp32 for the 3-pass mode)."""
    return round_tf32(w).contiguous(), w.contiguous().clone()


def pack_umma(w_tf32: Tensor) -> Tensor:
    """[tap][CinP][CoutP] (TF32) -> tcgen05 K-major canonical order [cout block][tap][CinP/4][NB][4],
    NB = min(CoutP, 64): element (n, k) of a tap at (k/4)*NB*16 B + n*16 B + (k%4)*4 B."""
    taps, cinp, coutp = w_tf32.shape
    nb = min(coutp, 64)
    if coutp % nb != 0 or nb % 16 != 0 or cinp % 8 != 0:
        return None                                  # shape not served by the tcgen05 kernels
    x = w_tf32.reshape(taps, cinp // 4, 4, coutp // nb, nb)          # [tap][kc][k4][cb][n]
    return x.permute(3, 0, 1, 4, 2).contiguous()                     # [cb][tap][kc][n][k4]


def pack_f16x3(w: Tensor) -> Tensor:
    """[tap][CinP][CoutP] fp32 -> [tap][CinK/2][CoutP][2] int32: per channel pair (k, k+1) and cout the
    half2 of the fp16 roundings (k in the low 16 bits) and the half2 of the fp16-rounded remainders.
    CinK = CinP rounded up to 16 (zero rows).  Values beyond +-65504 saturate like the kernel's split."""
    taps, cinp, coutp = w.shape
    cink = (cinp + 15) // 16 * 16
    x = torch.zeros(taps, cink, coutp, device=w.device, dtype=torch.float32)
    x[:, :cinp] = w
    hi = x.clamp(-65504.0, 65504.0).half()
    lo = (x - hi.float()).clamp(-65504.0, 65504.0).half()

    def pairs(h):   # [tap][CinK][CoutP] fp16 -> [tap][CinK/2][CoutP] int32 (k even in the low half)
        return h.reshape(taps, cink // 2, 2, coutp).permute(0, 1, 3, 2).contiguous().view(torch.int32).squeeze(-1)

    return torch.stack([pairs(hi), pairs(lo)], dim=-1).contiguous()


def pack_umma_f16(w: Tensor):
    """[tap][CinP][CoutP] fp32 -> fp16 hi / lo in the tcgen05 K-major canonical order [tap][hi | lo][CinK/8][CoutN][8 halves]
    (element (n, k) of a tap at ((k/8) * CoutN + n) * 16 B + (k%8) * 2 B), CinK = CinP rounded up to 16, CoutN = CoutP
    rounded up to 16 (the UMMA N granularity at M = 128; the extra output channels have zero weights)."""
    taps, cinp, coutp = w.shape
    coutn = (coutp + 15) // 16 * 16
    if coutn > 64:
        return None
    cink = (cinp + 15) // 16 * 16
    x = torch.zeros(taps, cink, coutn, device=w.device, dtype=torch.float32)
    x[:, :cinp, :coutp] = w
    hi = x.clamp(-65504.0, 65504.0).half()
    lo = (x - hi.float()).clamp(-65504.0, 65504.0).half()

    def canon(h):    # [tap][CinK][CoutN] -> [tap][CinK/8][CoutN][8]
        return h.reshape(taps, cink // 8, 8, coutn).permute(0, 1, 3, 2)

    return torch.stack([canon(hi), canon(lo)], dim=1).contiguous()


def pack_umma_f16i(w: Tensor):
    """pack_umma_f16 with hi and lo interleaved per K chunk: [tap][CinK/8][hi | lo][CoutP][8 halves].  One tap is then ONE
    contiguous block whose rows [0, CoutP) / [CoutP, 2 CoutP) of every chunk are B_hi / B_lo: the persistent TMA + tcgen05
    kernel (csrc/tc5pconv.cuh) issues A_hi x [B_hi | B_lo] as a single N = 2 CoutP instruction."""
    u = pack_umma_f16(w)
    if u is None:
        return None
    return u.permute(0, 2, 1, 3, 4).contiguous()          # [tap][hi|lo][KC][N][8] -> [tap][KC][hi|lo][N][8]


def _packs(out: Tensor):
    t, f = split_tf32(out)
    return t, f, pack_umma(t), pack_f16x3(f), pack_umma_f16(f), pack_umma_f16i(f)


def pack_mma_conv(w: Tensor, cinp: int = 0, coutp: int = 0):
    """nn.Conv2d weight [Cout,Cin,kh,kw] -> (tf32, fp32, umma, f16x3): [kh*kw][CinP][CoutP] twice, the
    tcgen05 order and the fp16 hi/lo split."""
    co, cin, kh, kw = w.shape
    cinp, coutp = cinp or _pad8(cin), coutp or _pad8(co)
    out = torch.zeros(kh * kw, cinp, coutp, device=w.device, dtype=torch.float32)
    out[:, :cin, :co] = w.detach().float().permute(2, 3, 1, 0).reshape(kh * kw, cin, co)
    return _packs(out)


def pack_mma_tconv(w: Tensor) -> Tuple[Tensor, Tensor]:
    """nn.ConvTranspose2d weight [Cin,Cout,kh,kw] -> (hi, lo) each [kh*kw][CinP][CoutP]."""
    cin, co, kh, kw = w.shape
    out = torch.zeros(kh * kw, _pad8(cin), _pad8(co), device=w.device, dtype=torch.float32)
    out[:, :cin, :co] = w.detach().float().permute(2, 3, 0, 1).reshape(kh * kw, cin, co)
    return _packs(out)


def pack_fc(w: Tensor) -> Tensor:
    """1x1 conv weight [Cout,Cin,1,1] -> fp32 [Cin][Cout] (FFMA epilogue kernels)."""
    return w.detach().float().reshape(w.shape[0], w.shape[1]).t().contiguous()


def _split_f16(x: Tensor) -> Tuple[Tensor, Tensor]:
    """x = hi + lo with hi = fp16(x) (saturating), lo = fp16(x - hi): the kernels' cvt.rn.satfinite split."""
    hi = x.clamp(-65504.0, 65504.0).half()
    lo = (x - hi.float()).clamp(-65504.0, 65504.0).half()
    return hi, lo


def _umma_kmajor(w_nk: Tensor) -> Tensor:
    """[N][K] fp16 (K a multiple of 8) -> the tcgen05 K-major / no-swizzle canonical order [K/8][N][8]."""
    n, k = w_nk.shape
    return w_nk.reshape(n, k // 8, 8).permute(1, 0, 2).contiguous()


def pack_head_fused(fc1_w: Tensor, fc2_w: Tensor, fc2_b: Tensor) -> Tensor:
    """depth_head.2.weight [64,32,1,1], depth_head.4.weight [256,64,1,1] + bias [256] -> the byte blob of
    csrc/headfused.cuh: W1 hi | W1 lo | W2 hi | W2 lo (fp16, UMMA canonical order) | bias (fp32)."""
    w1 = fc1_w.detach().float().reshape(64, 32)
    w2 = fc2_w.detach().float().reshape(256, 64)
    parts = []
    for w in (w1, w2):
        hi, lo = _split_f16(w)
        parts += [_umma_kmajor(hi).reshape(-1).view(torch.uint8), _umma_kmajor(lo).reshape(-1).view(torch.uint8)]
    parts.append(fc2_b.detach().float().contiguous().view(torch.uint8))
    blob = torch.cat(parts).contiguous()
    assert blob.numel() == 74752
    return blob


def _vec(t: Tensor) -> Tensor:
    return t.detach().float().reshape(-1).contiguous()


_CORR_CONVS = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5")


class _Holder:
    def __init__(self):
        self.keep = []

    def pair(self, hl) -> _lib.WPair:
        self.keep.extend(t for t in hl if t is not None)
        return _lib.WPair(hl[0].data_ptr(), hl[1].data_ptr(), hl[2].data_ptr() if hl[2] is not None else None, hl[3].data_ptr(),
                          hl[4].data_ptr() if hl[4] is not None else None, hl[5].data_ptr() if hl[5] is not None else None)

    def ptr(self, t: Tensor) -> int:
        self.keep.append(t)
        return t.data_ptr()


def fill_corrnet(h: _Holder, dst: _lib.CorrNetWeights, sd: Dict[str, Tensor], prefix: str, dev) -> None:
    g = lambda k: sd[prefix + k].to(dev)
    dst.conv0 = h.pair(pack_mma_conv(g("conv0.conv.weight")))
    dst.conv1 = h.pair(pack_mma_conv(g("conv1.conv.weight")))
    dst.conv2 = h.pair(pack_mma_conv(g("conv2.conv.weight")))
    dst.conv3 = h.pair(pack_mma_tconv(g("conv3.weight")))
    dst.conv4 = h.pair(pack_mma_tconv(g("conv4.weight")))
    dst.conv5 = h.pair(pack_mma_conv(g("conv5.weight")))
    dst.conv5_b = h.ptr(_vec(g("conv5.bias")))


def fill_pvw(h: _Holder, s: _lib.Weights, sd, prefix, dev) -> None:
    g = lambda k: sd[prefix + k].to(dev)
    s.pvw_conv0 = h.pair(pack_mma_conv(g("conv.0.conv.weight")))
    s.pvw_conv1 = h.ptr(_vec(g("conv.1.weight")))
    s.pvw_conv1_b = h.ptr(_vec(g("conv.1.bias")))


def fill_gru(h: _Holder, s: _lib.Weights, sd, prefix, dev) -> None:
    g = lambda k: sd[prefix + k].to(dev)
    # input channels [h(32), x(11)] -> stored [h(32), x(16)]: zero rows for the 5 padding channels
    s.gru_zr = h.pair(pack_mma_conv(torch.cat([g("convz.weight"), g("convr.weight")], 0), cinp=48))
    s.gru_zr_b = h.ptr(_vec(torch.cat([g("convz.bias"), g("convr.bias")], 0)))
    s.gru_q = h.pair(pack_mma_conv(g("convq.weight"), cinp=48))
    s.gru_q_b = h.ptr(_vec(g("convq.bias")))


def fill_update(h: _Holder, s: _lib.Weights, sd, prefix, dev) -> None:
    g = lambda k: sd[prefix + k].to(dev)
    fill_gru(h, s, sd, prefix + "gru.", dev)
    s.head_conv0 = h.pair(pack_mma_conv(torch.cat([g("depth_head.0.weight"), g("confidence_head.0.weight")], 0)))
    s.head_fc1 = h.pair(pack_mma_conv(g("depth_head.2.weight")))
    s.head_fc2 = h.pair(pack_mma_conv(g("depth_head.4.weight")))
    s.head_fc2_b = h.ptr(_vec(g("depth_head.4.bias")))
    s.head_fused = h.ptr(pack_head_fused(g("depth_head.2.weight"), g("depth_head.4.weight"), g("depth_head.4.bias")))
    s.conf_fc = h.ptr(_vec(g("confidence_head.2.weight")))
    s.conf_fc_b = h.ptr(_vec(g("confidence_head.2.bias")))
    s.hinit_conv0 = h.pair(pack_mma_conv(g("hidden_init_head.0.weight")))
    s.hinit_fc = h.pair(pack_mma_conv(g("hidden_init_head.2.weight")))
    s.hinit_fc_b = h.ptr(_vec(g("hidden_init_head.2.bias")))


class PackedWeights(_Holder):
    """sd: {name: tensor} with names relative to the IterMVS module
    ('evaluation.pixel_view_weight.conv.0.conv.weight', 'update.gru.convz.weight', ...)."""

    def __init__(self, sd: Dict[str, Tensor], device: torch.device):
        super().__init__()
        s = _lib.Weights()
        fill_pvw(self, s, sd, "evaluation.pixel_view_weight.", device)
        for i in range(3):
            fill_corrnet(self, s.corrnet[i], sd, f"evaluation.corr_conv1.{i}.", device)
        fill_update(self, s, sd, "update.", device)
        s.ups_conv0 = self.pair(pack_mma_conv(sd["upsample.0.weight"].to(device)))
        s.ups_fc = self.ptr(pack_fc(sd["upsample.2.weight"].to(device)))
        self.num_sample = int(sd["update.hidden_init_head.0.weight"].shape[1])
        self.struct = s

    @property
    def ref(self):
        return C.byref(self.struct)


# FeatureNet layer order of imvs_featurenet_weights (include/itermvs_b200.h):
#   (state_dict prefix, has_bn)
FNET_LAYERS = [("conv1.", True)]
for _l in (1, 2, 3):
    FNET_LAYERS += [(f"layer{_l}.0.conv1.", True), (f"layer{_l}.0.conv2.", True), (f"layer{_l}.0.downsample.", True),
                    (f"layer{_l}.1.conv1.", True), (f"layer{_l}.1.conv2.", True)]
FNET_LAYERS += [("output3.", False), ("inner2.", False), ("output2.", False), ("inner1.", False), ("output1.", False)]


def fold_bn(sd: Dict[str, Tensor], prefix: str, eps: float = 1e-5) -> Tuple[Tensor, Tensor]:
    """conv (no bias) + BatchNorm2d(eval) -> conv weight, bias  (module.py:6-29)."""
    w = sd[prefix + "conv.weight"].float()
    scale = sd[prefix + "bn.weight"].float() / torch.sqrt(sd[prefix + "bn.running_var"].float() + eps)
    return w * scale.view(-1, 1, 1, 1), sd[prefix + "bn.bias"].float() - sd[prefix + "bn.running_mean"].float() * scale


class PackedFeatureNet(_Holder):
    """sd: FeatureNet state_dict (names relative to the module), eval-mode semantics."""

    def __init__(self, sd: Dict[str, Tensor], device: torch.device):
        super().__init__()
        s = _lib.FeatureNetWeights()
        sd = {k: v.to(device) for k, v in sd.items()}
        for i, (prefix, has_bn) in enumerate(FNET_LAYERS):
            if has_bn:
                w, b = fold_bn(sd, prefix)
            else:
                w, b = sd[prefix + "weight"].float(), sd[prefix + "bias"].float()
            s.w[i] = self.pair(pack_mma_conv(w))
            s.b[i] = self.ptr(_vec(b))
        # slots 21..23: [layerK.0.conv1 | layerK.0.downsample] stacked on Cout (one GEMM over the shared input)
        for k in (1, 2, 3):
            w1, b1 = fold_bn(sd, f"layer{k}.0.conv1.")
            w2, b2 = fold_bn(sd, f"layer{k}.0.downsample.")
            s.w[20 + k] = self.pair(pack_mma_conv(torch.cat([w1, w2], 0)))
            s.b[20 + k] = self.ptr(_vec(torch.cat([b1, b2], 0)))
        self.struct = s

    @property
    def ref(self):
        return C.byref(self.struct)
