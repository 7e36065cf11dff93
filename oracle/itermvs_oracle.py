"""CPU oracle for the IterMVS hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch-on-CPU restatement of the reference algorithm
(FangjinhuaWang/IterMVS @ 453e9c7, files models/module.py, models/itermvs.py, models/net.py),
written functionally over a flat {state_dict key -> tensor} weight dict. Each function cites the
reference file:line it follows. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import this package; the product package
`itermvs_b200` never does (its CUDA path raises if the extension is missing).

PINNING. The reference has no tests or golden vectors. This oracle is
pinned against outputs of the reference itself, executed in the build container by
`tests/golden/make_golden.py` (which imports /root/reference/models and the shipped DTU
checkpoint) and committed as fixtures under `tests/golden/`; `tests/test_oracle_golden.py`
replays them. The sampling step is restated with explicit floor / four-tap gathers rather than
`F.grid_sample`, and the x2 / x0.5 / x4 resamplers with explicit index arithmetic rather than
`F.interpolate`, so those library semantics are themselves checked by the golden replay.

Layouts here are the reference's (NCHW, lists of views); the CUDA path uses its own layouts and
is compared after conversion.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Weights = Dict[str, Tensor]

GROUPS = 8            # itermvs.py:28
OUT_BINS = 256        # itermvs.py:134
RADIUS = 4            # itermvs.py:135
INTERVAL_SCALE = 1.0 / 256  # itermvs.py:229
CORR_INTERVAL = {     # itermvs.py:231-235
    1: (-2.0, -2.0 / 3, 2.0 / 3, 2.0),
    2: (-8.0, -8.0 / 3, 8.0 / 3, 8.0),
    3: (-32.0, 32.0),
}


def strip_module_prefix(state_dict: Dict[str, Tensor]) -> Weights:
    """Checkpoints are saved from nn.DataParallel (train.py:153-157): keys carry 'module.'."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}


# --------------------------------------------------------------------------------------------
# resamplers (semantics of F.interpolate(mode='bilinear', align_corners=False), restated)
# --------------------------------------------------------------------------------------------
def _lin_up_axis(x: Tensor, dim: int, factor: int) -> Tensor:
    """1-D linear upsampling by an integer factor, align_corners=False:
    src = (dst + 0.5) / factor - 0.5, clamped below at 0; i1 = min(i0 + 1, n - 1)."""
    n = x.shape[dim]
    dst = torch.arange(n * factor, dtype=torch.float32, device=x.device)
    src = ((dst + 0.5) / factor - 0.5).clamp_(min=0.0)
    i0 = src.floor().to(torch.long)
    i1 = (i0 + 1).clamp_(max=n - 1)
    lam = (src - i0.to(torch.float32))
    shape = [1] * x.dim()
    shape[dim] = -1
    lam = lam.view(shape)
    return x.index_select(dim, i0) * (1.0 - lam) + x.index_select(dim, i1) * lam


def bilinear_up(x: Tensor, factor: int) -> Tensor:
    """F.interpolate(x, scale_factor=factor, mode='bilinear') for integer factor (2 or 4).
    ATen interpolates rows (h) first for each output, result is separable; we apply h then w."""
    return _lin_up_axis(_lin_up_axis(x, -2, factor), -1, factor)


def mean_pool2(x: Tensor) -> Tensor:
    """F.interpolate(x, scale_factor=0.5, mode='bilinear'): src = 2*dst + 0.5 => 2x2 mean
    (itermvs.py:95-98 for level 1)."""
    a = x[..., 0::2, 0::2]
    b = x[..., 0::2, 1::2]
    c = x[..., 1::2, 0::2]
    d = x[..., 1::2, 1::2]
    # ATen: h0lambda*(w0lambda*a + w1lambda*b) + h1lambda*(w0lambda*c + w1lambda*d), lambdas = 0.5
    return 0.5 * (0.5 * a + 0.5 * b) + 0.5 * (0.5 * c + 0.5 * d)


# --------------------------------------------------------------------------------------------
# depth <-> normalized inverse depth                                      module.py:142-152
# --------------------------------------------------------------------------------------------
def depth_unnormalization(nd: Tensor, inv_min: Tensor, inv_max: Tensor) -> Tensor:
    return 1.0 / (inv_max + nd * (inv_min - inv_max))


def depth_normalization(depth: Tensor, inv_min: Tensor, inv_max: Tensor) -> Tensor:
    return (1.0 / (depth + 1e-5) - inv_max) / (inv_min - inv_max)


def initial_depth_samples(inv_min: Tensor, inv_max: Tensor, num: int, h: int, w: int) -> Tensor:
    """itermvs.py:11-19. inv_* are [B,1,1,1]; returns [B,num,h,w]."""
    b = inv_min.shape[0]
    idx = torch.arange(num, dtype=torch.float32, device=inv_min.device).view(1, num, 1, 1).repeat(b, 1, h, w) / (num - 1)
    return 1.0 / (inv_max + idx * (inv_min - inv_max))


def iteration_depth_samples(nd: Tensor, level: int, inv_min: Tensor, inv_max: Tensor) -> Tensor:
    """itermvs.py:289-293."""
    off = torch.tensor(CORR_INTERVAL[level], dtype=torch.float32, device=nd.device).view(1, -1, 1, 1)
    s = (nd + off * INTERVAL_SCALE).clamp(min=0, max=1)
    return depth_unnormalization(s, inv_min, inv_max)


# --------------------------------------------------------------------------------------------
# plane-sweep warp                                                          module.py:68-125
# --------------------------------------------------------------------------------------------
def compose_projection(src_proj: Tensor, ref_proj: Tensor) -> Tensor:
    """module.py:78-87: proj = src_proj @ inverse(ref_proj), [B,4,4]."""
    return torch.matmul(src_proj, torch.inverse(ref_proj))


def warp_sample_positions(proj: Tensor, depth_samples: Tensor, h1: int, w1: int):
    """module.py:89-109. Returns (u, v) in source-feature pixel units, each [B,D,H*W]."""
    b, d, h, w = depth_samples.shape
    rot = proj[:, :3, :3]
    trans = proj[:, :3, 3:4]
    dev = depth_samples.device
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev), torch.arange(w, dtype=torch.float32, device=dev),
                            indexing="ij")
    xs = xs.reshape(-1) * (w1 / w)                      # module.py:95-96 (no half-pixel offset)
    ys = ys.reshape(-1) * (h1 / h)
    xyz = torch.stack((xs, ys, torch.ones_like(xs))).unsqueeze(0).expand(b, 3, h * w)
    rot_xyz = torch.matmul(rot, xyz)                    # [B,3,HW]
    p = rot_xyz.unsqueeze(2) * depth_samples.reshape(b, 1, d, h * w) + trans.view(b, 3, 1, 1)
    px, py, pz = p[:, 0], p[:, 1], p[:, 2]
    bad = ~(pz > 1e-2)                                  # module.py:105-108 (note: NaN also -> bad)
    px = torch.where(bad, torch.full_like(px, float(w)), px)
    py = torch.where(bad, torch.full_like(py, float(h)), py)
    pz = torch.where(bad, torch.ones_like(pz), pz)
    return px / pz, py / pz


def differentiable_warping(src_fea: Tensor, src_proj: Tensor, ref_proj: Tensor, depth_samples: Tensor
                           ) -> Tensor:
    """module.py:68-125 with grid_sample(bilinear, zeros, align_corners=True) restated as
    floor + 4 taps, each tap zeroed when outside [0,W1-1]x[0,H1-1]. The reference normalises
    (u / ((W1-1)/2) - 1) and grid_sample un-normalises ((g + 1)/2 * (W1-1)); we apply the same
    two roundings so the sample position is the one ATen sees. Returns [B,C,D,H,W]."""
    b, c, h1, w1 = src_fea.shape
    _, d, h, w = depth_samples.shape
    proj = compose_projection(src_proj, ref_proj)
    u, v = warp_sample_positions(proj, depth_samples, h1, w1)
    gx = u / ((w1 - 1) / 2) - 1                         # module.py:112-113
    gy = v / ((h1 - 1) / 2) - 1
    ix = ((gx + 1) / 2) * (w1 - 1)                      # ATen grid_sampler_unnormalize, align_corners
    iy = ((gy + 1) / 2) * (h1 - 1)
    x0 = ix.floor()
    y0 = iy.floor()
    fx = ix - x0
    fy = iy - y0
    flat = src_fea.reshape(b, c, h1 * w1)
    out = torch.zeros(b, c, d * h * w, dtype=src_fea.dtype, device=src_fea.device)
    for dy, dx, wgt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                        (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
        xx = x0 + dx
        yy = y0 + dy
        ok = (xx >= 0) & (xx <= w1 - 1) & (yy >= 0) & (yy <= h1 - 1)
        idx = (yy.clamp(0, h1 - 1) * w1 + xx.clamp(0, w1 - 1)).to(torch.long).reshape(b, 1, -1)
        tap = torch.gather(flat, 2, idx.expand(b, c, -1))
        out += tap * (wgt * ok.to(wgt.dtype)).reshape(b, 1, -1)
    return out.view(b, c, d, h, w)


def group_correlation(warped: Tensor, ref_fea: Tensor) -> Tensor:
    """itermvs.py:49-51 / 103-104: mean over the C/G channels of each group. -> [B,G,D,H,W]."""
    b, c, d, h, w = warped.shape
    prod = warped.view(b, GROUPS, c // GROUPS, d, h, w) * ref_fea.view(b, GROUPS, c // GROUPS, 1, h, w)
    return prod.mean(dim=2)


# --------------------------------------------------------------------------------------------
# small conv nets of the evaluation stage
# --------------------------------------------------------------------------------------------
def pixel_view_weight(wts: Weights, corr: Tensor, prefix="iter_mvs.evaluation.pixel_view_weight.") -> Tensor:
    """itermvs.py:341-350. corr [B,G,D,H,W] -> [B,1,H,W]."""
    b, g, d, h, w = corr.shape
    x = corr.permute(0, 2, 1, 3, 4).reshape(b * d, g, h, w)
    x = F.relu(F.conv2d(x, wts[prefix + "conv.0.conv.weight"], padding=1))
    x = F.conv2d(x, wts[prefix + "conv.1.weight"], wts[prefix + "conv.1.bias"]).view(b, d, h, w)
    return torch.softmax(x, dim=1).max(dim=1, keepdim=True)[0]


def corr_net(wts: Weights, corr: Tensor, prefix: str) -> Tensor:
    """itermvs.py:367-381. corr [B,G,R,H,W] -> [B,R,H,W]."""
    b, g, r, h, w = corr.shape
    x = corr.permute(0, 2, 1, 3, 4).reshape(b * r, g, h, w)
    c0 = F.relu(F.conv2d(x, wts[prefix + "conv0.conv.weight"], padding=1))
    c1 = F.relu(F.conv2d(c0, wts[prefix + "conv1.conv.weight"], stride=2, padding=1))
    c2 = F.relu(F.conv2d(c1, wts[prefix + "conv2.conv.weight"], stride=2, padding=1))
    x = c1 + F.conv_transpose2d(c2, wts[prefix + "conv3.weight"], stride=2, padding=1, output_padding=1)
    x = c0 + F.conv_transpose2d(x, wts[prefix + "conv4.weight"], stride=2, padding=1, output_padding=1)
    x = F.conv2d(x, wts[prefix + "conv5.weight"], wts[prefix + "conv5.bias"], padding=1)
    return x.view(b, r, h, w)


# --------------------------------------------------------------------------------------------
# Evaluation                                                               itermvs.py:33-126
# --------------------------------------------------------------------------------------------
def evaluation_init(wts: Weights, ref_fea3: Tensor, src_feas3: Sequence[Tensor], ref_proj3: Tensor,
                    src_projs3: Sequence[Tensor], depth_sample: Tensor, inv_min: Tensor, inv_max: Tensor):
    """itermvs.py:36-82 (view_weights is None branch).
    Returns dict(view_weights [B,S,2H3,2W3], corr [B,D,H3,W3], depth [B,1,2H3,2W3],
                 per_view_corr list, per_view_weight list)."""
    ev = "iter_mvs.evaluation."
    corr_sum = 0
    vw_sum = 1e-5
    vws, per_corr, per_vw = [], [], []
    for src, sp in zip(src_feas3, src_projs3):
        corr = group_correlation(differentiable_warping(src, sp, ref_proj3, depth_sample), ref_fea3)
        vw = pixel_view_weight(wts, corr)
        per_corr.append(corr)
        per_vw.append(vw)
        vws.append(bilinear_up(vw, 2))
        corr_sum = corr_sum + corr * vw.unsqueeze(1)
        vw_sum = vw_sum + vw.unsqueeze(1)
    agg = corr_sum / vw_sum
    corr = corr_net(wts, agg, ev + "corr_conv1.2.")
    d = depth_sample.shape[1]
    prob = torch.softmax(corr, dim=1)
    idx = (torch.arange(d, dtype=torch.float32, device=corr.device).view(1, d, 1, 1) * prob).sum(dim=1, keepdim=True)
    depth = bilinear_up(depth_unnormalization(idx / (d - 1.0), inv_min, inv_max), 2)
    return {"view_weights": torch.cat(vws, dim=1), "corr": corr, "depth": depth, "aggregated": agg,
            "per_view_corr": per_corr, "per_view_weight": per_vw}


def resample_ref_feature(ref_fea: Tensor, level: int) -> Tensor:
    """itermvs.py:95-98: bring the level-l reference feature to level-2 resolution."""
    if level == 1:
        return mean_pool2(ref_fea)
    if level == 3:
        return bilinear_up(ref_fea, 2)
    return ref_fea


def evaluation_iter(wts: Weights, ref_feas: Dict[str, Tensor], src_feas: Dict[str, Sequence[Tensor]],
                    ref_projs: Dict[str, Tensor], src_projs: Dict[str, Sequence[Tensor]],
                    depth_samples: Dict[str, Tensor], view_weights: Tensor, return_aggregated: bool = False):
    """itermvs.py:84-126. Returns corr [B,10,H2,W2] (and the three aggregated volumes on request)."""
    outs, aggs = [], []
    for l in (1, 2, 3):
        key = f"level{l}"
        ref_l = resample_ref_feature(ref_feas[key], l)
        ds = depth_samples[key]
        b, r, h, w = ds.shape
        corr_sum = 0
        vw_sum = 1e-5
        for i, (src, sp) in enumerate(zip(src_feas[key], src_projs[key])):
            corr = group_correlation(differentiable_warping(src, sp, ref_projs[key], ds), ref_l)
            vw = view_weights[:, i].reshape(b, 1, 1, h, w)
            corr_sum = corr_sum + corr * vw
            vw_sum = vw_sum + vw
        agg = corr_sum / vw_sum
        aggs.append(agg)
        outs.append(corr_net(wts, agg, f"iter_mvs.evaluation.corr_conv1.{l - 1}."))
    corr = torch.cat(outs, dim=1)
    return (corr, aggs) if return_aggregated else corr


# --------------------------------------------------------------------------------------------
# Update                                                                  itermvs.py:129-220
# --------------------------------------------------------------------------------------------
def conv_gru(wts: Weights, h: Tensor, x: Tensor, prefix="iter_mvs.update.gru.") -> Tensor:
    """module.py:59-66 (3x3, dilation 2, padding 2, bias)."""
    hx = torch.cat([h, x], dim=1)
    z = torch.sigmoid(F.conv2d(hx, wts[prefix + "convz.weight"], wts[prefix + "convz.bias"], padding=2, dilation=2))
    r = torch.sigmoid(F.conv2d(hx, wts[prefix + "convr.weight"], wts[prefix + "convr.bias"], padding=2, dilation=2))
    q = torch.tanh(F.conv2d(torch.cat([r * h, x], dim=1), wts[prefix + "convq.weight"], wts[prefix + "convq.bias"],
                            padding=2, dilation=2))
    return (1 - z) * h + z * q


def depth_head_logits(wts: Weights, hidden: Tensor, prefix="iter_mvs.update.depth_head.") -> Tensor:
    """itermvs.py:139-145."""
    x = F.relu(F.conv2d(hidden, wts[prefix + "0.weight"], padding=2, dilation=2))
    x = F.relu(F.conv2d(x, wts[prefix + "2.weight"]))
    return F.conv2d(x, wts[prefix + "4.weight"], wts[prefix + "4.bias"])


def confidence_logit(wts: Weights, hidden: Tensor, prefix="iter_mvs.update.confidence_head.") -> Tensor:
    """itermvs.py:147-151."""
    x = F.relu(F.conv2d(hidden, wts[prefix + "0.weight"], padding=2, dilation=2))
    return F.conv2d(x, wts[prefix + "2.weight"], wts[prefix + "2.bias"])


def window_regression(prob: Tensor) -> Tensor:
    """itermvs.py:173-190 / 203-219: arg-max bin, clamped +-RADIUS window (edge bins are counted
    repeatedly, as in the reference), sum(idx*p)/(1e-6+sum p), divided by OUT_BINS-1."""
    n = prob.shape[1]
    top = torch.argmax(prob, dim=1, keepdim=True).to(torch.float32)
    num = 0
    den = 1e-6
    for i in range(2 * RADIUS + 1):
        idx = (top - RADIUS + i).clamp(min=0, max=n - 1).to(torch.long)
        p = torch.gather(prob, 1, idx)
        num = num + idx * p
        den = den + p
    return (num / den) / (n - 1.0)


def hidden_init(wts: Weights, corr: Tensor, prefix="iter_mvs.update.hidden_init_head.") -> Tensor:
    """itermvs.py:159-164."""
    x = F.relu(F.conv2d(corr, wts[prefix + "0.weight"], padding=1))
    x = F.conv2d(x, wts[prefix + "2.weight"], wts[prefix + "2.bias"])
    return torch.tanh(bilinear_up(x, 2))


def depth_init(wts: Weights, hidden: Tensor):
    """itermvs.py:171-190. Returns (normalized_depth, probability)."""
    prob = torch.softmax(depth_head_logits(wts, hidden), dim=1)
    return window_regression(prob), prob


def update_step(wts: Weights, hidden: Tensor, nd: Tensor, corr: Tensor, confidence_flag: bool):
    """itermvs.py:192-220. Returns (hidden, nd, probability, confidence, confidence_logit)."""
    hidden = conv_gru(wts, hidden, torch.cat([nd, corr], dim=1))
    c0 = confidence_logit(wts, hidden) if confidence_flag else None
    prob = torch.softmax(depth_head_logits(wts, hidden), dim=1)
    return hidden, window_regression(prob), prob, (torch.sigmoid(c0) if confidence_flag else None), c0


# --------------------------------------------------------------------------------------------
# output stage                                     module.py:127-140, itermvs.py:262-264,321-324
# --------------------------------------------------------------------------------------------
def upsample_weights(wts: Weights, ref_fea2: Tensor, prefix="iter_mvs.upsample.") -> Tensor:
    """itermvs.py:262-264 -> [B,1,9,4,4,H,W], softmax over the 9 taps."""
    b, _, h, w = ref_fea2.shape
    x = F.relu(F.conv2d(ref_fea2, wts[prefix + "0.weight"], padding=1))
    x = F.conv2d(x, wts[prefix + "2.weight"]).view(b, 1, 9, 4, 4, h, w)
    return torch.softmax(x, dim=2)


def convex_upsample(x: Tensor, weight: Tensor, scale: int = 4) -> Tensor:
    """module.py:127-140. x [B,1,H,W]; replicate pad; 3x3 neighbourhood (tap k = ky*3+kx)."""
    b, _, h, w = x.shape
    xp = F.pad(x, (1, 1, 1, 1), mode="replicate")
    nb = torch.stack([xp[:, :, ky:ky + h, kx:kx + w] for ky in range(3) for kx in range(3)], dim=2)  # [B,1,9,H,W]
    up = (nb.view(b, 1, 9, 1, 1, h, w) * weight).sum(dim=2)           # [B,1,4,4,H,W]
    return up.permute(0, 1, 4, 2, 5, 3).reshape(b, 1, scale * h, scale * w)


# --------------------------------------------------------------------------------------------
# IterMVS.forward                                                        itermvs.py:253-329
# --------------------------------------------------------------------------------------------
def itermvs_forward(wts: Weights, ref_feas: Dict[str, Tensor], src_feas: Dict[str, Sequence[Tensor]],
                    ref_projs: Dict[str, Tensor], src_projs: Dict[str, Sequence[Tensor]],
                    depth_min: Tensor, depth_max: Tensor, iteration: int, num_sample: int = 32,
                    trace: Optional[dict] = None):
    """Test-mode forward. Returns (depth, depth_upsampled, confidence, confidence_upsampled).
    If `trace` is a dict it receives every intermediate (used for stage-wise parity)."""
    b, _, h2, w2 = ref_feas["level2"].shape
    up_w = upsample_weights(wts, ref_feas["level2"])
    inv_min = (1.0 / depth_min).view(b, 1, 1, 1)
    inv_max = (1.0 / depth_max).view(b, 1, 1, 1)
    ds0 = initial_depth_samples(inv_min, inv_max, num_sample, h2 // 2, w2 // 2)
    ev = evaluation_init(wts, ref_feas["level3"], src_feas["level3"], ref_projs["level3"], src_projs["level3"],
                         ds0, inv_min, inv_max)
    view_weights = ev["view_weights"]
    hidden = hidden_init(wts, ev["corr"])
    nd, prob0 = depth_init(wts, hidden)
    if trace is not None:
        trace.update(upsample_weight=up_w, view_weights=view_weights, corr_init=ev["corr"],
                     aggregated_init=ev["aggregated"], per_view_corr=ev["per_view_corr"],
                     per_view_weight=ev["per_view_weight"], depth_initial=ev["depth"],
                     hidden0=hidden, nd0=nd, prob0=prob0, corr_iter=[], agg_iter=[], hidden_iter=[],
                     nd_iter=[], prob_iter=[])
    depth = depth_up = conf = conf_up = None
    for it in range(iteration):
        samples = {f"level{l}": iteration_depth_samples(nd, l, inv_min, inv_max) for l in (1, 2, 3)}
        corr, aggs = evaluation_iter(wts, ref_feas, src_feas, ref_projs, src_projs, samples, view_weights,
                                     return_aggregated=True)
        last = it == iteration - 1
        if last:
            depth = depth_unnormalization(nd, inv_min, inv_max)      # itermvs.py:319 (pre-update value)
        hidden, nd, prob, conf_new, _ = update_step(wts, hidden, nd, corr, confidence_flag=last)
        if trace is not None:
            trace["corr_iter"].append(corr)
            trace["agg_iter"].append(aggs)
            trace["hidden_iter"].append(hidden)
            trace["nd_iter"].append(nd)
            trace["prob_iter"].append(prob)
        if last:
            conf = conf_new
            depth_up = depth_unnormalization(convex_upsample(nd, up_w), inv_min, inv_max)
            conf_up = bilinear_up(conf, 4)
    return depth, depth_up, conf, conf_up


# --------------------------------------------------------------------------------------------
# FeatureNet + Pipeline (test mode, BN in eval mode)                         net.py:7-128
# --------------------------------------------------------------------------------------------
def _conv_bn(wts: Weights, x: Tensor, prefix: str, stride: int = 1, relu: bool = True) -> Tensor:
    """module.py:6-29 (ConvBnReLU / ConvBn), BatchNorm with running statistics, eps 1e-5."""
    x = F.conv2d(x, wts[prefix + "conv.weight"], stride=stride, padding=1)
    x = F.batch_norm(x, wts[prefix + "bn.running_mean"], wts[prefix + "bn.running_var"],
                     wts[prefix + "bn.weight"], wts[prefix + "bn.bias"], training=False, eps=1e-5)
    return F.relu(x) if relu else x


def _res_block(wts: Weights, x: Tensor, prefix: str, stride: int) -> Tensor:
    """module.py:32-50."""
    y = _conv_bn(wts, _conv_bn(wts, x, prefix + "conv1.", stride), prefix + "conv2.", 1, relu=False)
    if stride != 1:
        x = _conv_bn(wts, x, prefix + "downsample.", stride, relu=False)
    return F.relu(x + y)


def feature_net(wts: Weights, img: Tensor, prefix="feature_net.") -> Dict[str, Tensor]:
    """net.py:56-65 for one view batch [B,3,H,W] -> {'level1','level2','level3'} NCHW."""
    f0 = _conv_bn(wts, img, prefix + "conv1.")
    f1 = _res_block(wts, _res_block(wts, f0, prefix + "layer1.0.", 2), prefix + "layer1.1.", 1)
    f2 = _res_block(wts, _res_block(wts, f1, prefix + "layer2.0.", 2), prefix + "layer2.1.", 1)
    f3 = _res_block(wts, _res_block(wts, f2, prefix + "layer3.0.", 2), prefix + "layer3.1.", 1)
    out = {"level3": F.conv2d(f3, wts[prefix + "output3.weight"], wts[prefix + "output3.bias"], padding=1)}
    intra = bilinear_up(f3, 2) + F.conv2d(f2, wts[prefix + "inner2.weight"], wts[prefix + "inner2.bias"])
    out["level2"] = F.conv2d(intra, wts[prefix + "output2.weight"], wts[prefix + "output2.bias"], padding=1)
    intra = bilinear_up(intra, 2) + F.conv2d(f1, wts[prefix + "inner1.weight"], wts[prefix + "inner1.bias"])
    out["level1"] = F.conv2d(intra, wts[prefix + "output1.weight"], wts[prefix + "output1.bias"], padding=1)
    return out


def pipeline_forward(wts: Weights, imgs: Dict[str, Tensor], proj_matrices: Dict[str, Tensor], depth_min: Tensor,
                     depth_max: Tensor, iteration: int = 4, num_sample: int = 32, trace: Optional[dict] = None):
    """net.py:78-128, test mode. Returns {'depths_upsampled','confidence_upsampled'} (+ 'depth',
    'confidence' at quarter resolution for convenience)."""
    with torch.no_grad():
        views = torch.unbind(imgs["level_0"], 1)
        feats = [feature_net(wts, v) for v in views]
        ref = {k: feats[0][k] for k in ("level1", "level2", "level3")}
        src = {k: [f[k] for f in feats[1:]] for k in ("level1", "level2", "level3")}
        rp, sp = {}, {}
        for l in (1, 2, 3):
            pm = torch.unbind(proj_matrices[f"level_{l}"].float(), 1)
            rp[f"level{l}"] = pm[0]
            sp[f"level{l}"] = list(pm[1:])
        if trace is not None:
            trace["ref_feature"] = ref
            trace["src_features"] = src
        depth, depth_up, conf, conf_up = itermvs_forward(wts, ref, src, rp, sp, depth_min.float(), depth_max.float(),
                                                         iteration, num_sample, trace)
    return {"depths_upsampled": depth_up, "confidence_upsampled": conf_up, "depth": depth, "confidence": conf}
