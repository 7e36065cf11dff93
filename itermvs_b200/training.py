"""Training-mode forward + backward of `Pipeline` (reference models/net.py:78-120 and the `self.training` / `not
self.test` branches of models/itermvs.py:253-329 as train.py:194-243 drives them).

Split of work (SURVEY 8b "Autograd"):
  * the plane sweep -- hypotheses -> homography warp -> bilinear sampling -> group-wise correlation (-> view-weighted
    aggregation in the iterations) -- runs on the fused sm_100a kernels in BOTH directions: `FusedCorrInit` /
    `FusedCorrIter` are torch.autograd.Functions over imvs_warpcorr_init / _iter and their *_backward twins.  Nothing
    is saved for the backward except the inputs: the reference's autograd graph keeps a [C,R,H,W] warped volume and
    the same-size product per differentiable_warping call (52 calls at 4 views / 4 iterations: 574 MB of the 1 721 MB
    a 640x512 training forward saves for backward, tools/saved_for_backward.py); here the backward recomputes the
    sampling positions.  Gradients go to the feature pyramids only --
    the grid is built under no_grad in the reference (module.py:77), the iteration's view weights and hypotheses are
    detached (itermvs.py:295, 282-283).
  * the convolution stacks (FeatureNet with BatchNorm batch statistics, PixelViewWeight, CorrNet, ConvGRU, heads,
    upsampling weights) run as ATen / cuDNN modules with torch autograd -- the very nn.Conv2d / BatchNorm2d members
    that hold the parameters for the inference kernels, so state_dict, optimizer and checkpoints are shared.  The
    tensor-core inference kernels have no backward; they stay the eval()/test path.

There is no CPU path: the fused operators need the CUDA library (the CPU test-suite substitutes an emulation of the kernel sources for it; the product
never does).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

from . import _lib
from . import ops

Tensor = torch.Tensor


# ---- backend hooks (tests/test_training_cpu.py points them at the kernel-source simulation) -------------------
def _L():
    return _lib.lib()


def _st():
    return ops._stream()


def _chk(t: Tensor, name: str) -> Tensor:
    return ops._chk(t, name)


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"itermvs_b200 {what}: {_L().imvs_last_error().decode('utf-8', 'replace')}")


def _compose(proj: Tensor, nan_flag=None) -> Tensor:
    """[B,V,4,4] (view 0 = reference) -> [B,V-1,12] rot|trans of src @ inverse(ref); constants for autograd.
    `nan_flag` (ops.NanFlag): the deferred form of the reference's `assert not isnan(proj)` (module.py:83, 87)."""
    proj = _chk(proj.detach().float(), "proj")
    b, v = proj.shape[:2]
    out = torch.empty(b, v - 1, 12, device=proj.device, dtype=torch.float32)
    _check(_L().imvs_compose_projections(proj.data_ptr(), b, v, out.data_ptr(), nan_flag.ptr() if nan_flag is not None else None,
                                         _st()), "compose_projections")
    return out


# ---- fused plane-sweep operators with CUDA backward ----------------------------------------------------------
class FusedCorrInit(torch.autograd.Function):
    """fea3 [B,V,H3,W3,48] (view 0 = reference), rt3 [B,S,12], depth_sample [B,D,H3,W3] ->
    per-view group correlation [B,S,D,P3,8] (itermvs.py:45-51)."""

    @staticmethod
    def forward(ctx, fea3: Tensor, rt3: Tensor, depth_sample: Tensor) -> Tensor:
        fea3, ds = _chk(fea3, "fea3"), _chk(depth_sample, "depth_sample")
        b, v, h3, w3, _ = fea3.shape
        d = ds.shape[1]
        corr = torch.empty(b, v - 1, d, h3 * w3, 8, device=fea3.device, dtype=torch.float32)
        _check(_L().imvs_warpcorr_init(fea3.data_ptr(), rt3.data_ptr(), None, None, ds.data_ptr(), corr.data_ptr(),
                                       b, v, h3, w3, d, _st()), "warpcorr_init")
        ctx.save_for_backward(fea3, rt3, ds)
        return corr

    @staticmethod
    def backward(ctx, grad_corr: Tensor):
        fea3, rt3, ds = ctx.saved_tensors
        b, v, h3, w3, _ = fea3.shape
        d = ds.shape[1]
        g = _chk(grad_corr.float(), "grad_corr")
        gfea = torch.empty_like(fea3)
        _check(_L().imvs_warpcorr_init_backward(fea3.data_ptr(), rt3.data_ptr(), None, None, ds.data_ptr(), g.data_ptr(),
                                                gfea.data_ptr(), b, v, h3, w3, d, _st()), "warpcorr_init_backward")
        return gfea, None, None


class FusedCorrIter(torch.autograd.Function):
    """Three pyramids [B,V,H_l,W_l,C_l], composed projections, explicit hypotheses [B,R_l,H2,W2] per level and the
    (detached) view weights [B,S,H2,W2] -> aggregated correlation [B,10,P2,8] (itermvs.py:86-120)."""

    @staticmethod
    def forward(ctx, fea1, fea2, fea3, rt1, rt2, rt3, smp1, smp2, smp3, vw) -> Tensor:
        feas = [_chk(f, "fea") for f in (fea1, fea2, fea3)]
        smps = [_chk(s, "depth_sample") for s in (smp1, smp2, smp3)]
        vw = _chk(vw, "view_weights")
        b, v, h2, w2, _ = feas[1].shape
        agg = torch.empty(b, 10, h2 * w2, 8, device=vw.device, dtype=torch.float32)
        _check(_L().imvs_warpcorr_iter(feas[0].data_ptr(), feas[1].data_ptr(), feas[2].data_ptr(), rt1.data_ptr(), rt2.data_ptr(),
                                       rt3.data_ptr(), None, 0, 1, vw.data_ptr(), None, None, smps[0].data_ptr(),
                                       smps[1].data_ptr(), smps[2].data_ptr(), agg.data_ptr(), b, v, h2, w2, _st()), "warpcorr_iter")
        ctx.save_for_backward(*feas, rt1, rt2, rt3, *smps, vw)
        return agg

    @staticmethod
    def backward(ctx, grad_agg: Tensor):
        f1, f2, f3, rt1, rt2, rt3, s1, s2, s3, vw = ctx.saved_tensors
        b, v, h2, w2, _ = f2.shape
        g = _chk(grad_agg.float(), "grad_agg")
        g1, g2, g3 = torch.empty_like(f1), torch.empty_like(f2), torch.empty_like(f3)
        _check(_L().imvs_warpcorr_iter_backward(f1.data_ptr(), f2.data_ptr(), f3.data_ptr(), rt1.data_ptr(), rt2.data_ptr(),
                                                rt3.data_ptr(), None, 0, 1, vw.data_ptr(), None, None, s1.data_ptr(), s2.data_ptr(),
                                                s3.data_ptr(), g.data_ptr(), g1.data_ptr(), g2.data_ptr(), g3.data_ptr(),
                                                b, v, h2, w2, _st()), "warpcorr_iter_backward")
        return g1, g2, g3, None, None, None, None, None, None, None


# ---- convolution stacks on the parameter-holding modules (ATen / cuDNN, torch autograd) -----------------------
def _conv_bn(m, x: Tensor, relu: bool) -> Tensor:
    y = m.bn(m.conv(x))                        # BatchNorm2d follows m.training: batch statistics when training
    return F.relu(y) if relu else y


def _res_block(m, x: Tensor) -> Tensor:        # module.py:32-50
    y = _conv_bn(m.conv2, _conv_bn(m.conv1, x, True), False)
    if m.downsample is not None:
        x = _conv_bn(m.downsample, x, False)
    return F.relu(x + y)


def featurenet_pyramids(fnet, imgs: Tensor):
    """net.py:36-50 (the `not self.test` branch: all B*V views in one batch) -> channels-last pyramids
    [B,V,H_l,W_l,C_l] for levels 1, 2, 3, differentiable."""
    b, v, _, h, w = imgs.shape
    x = imgs.reshape(b * v, -1, h, w).float()
    fea0 = _conv_bn(fnet.conv1, x, True)
    fea1 = _res_block(fnet.layer1[1], _res_block(fnet.layer1[0], fea0))
    fea2 = _res_block(fnet.layer2[1], _res_block(fnet.layer2[0], fea1))
    fea3 = _res_block(fnet.layer3[1], _res_block(fnet.layer3[0], fea2))
    out3 = fnet.output3(fea3)
    intra = F.interpolate(fea3, scale_factor=2, mode="bilinear") + fnet.inner2(fea2)
    out2 = fnet.output2(intra)
    intra = F.interpolate(intra, scale_factor=2, mode="bilinear") + fnet.inner1(fea1)
    out1 = fnet.output1(intra)

    def cl(t):
        return t.reshape(b, v, t.shape[1], t.shape[2], t.shape[3]).permute(0, 1, 3, 4, 2).contiguous()
    return cl(out1), cl(out2), cl(out3)


def _slices(vol: Tensor) -> Tensor:
    """[B,G,N,H,W] -> [B*N,G,H,W] (itermvs.py:344-346, 369-371)."""
    b, g, n, h, w = vol.shape
    return vol.permute(0, 2, 1, 3, 4).reshape(b * n, g, h, w)


def pixel_view_weight(m, corr: Tensor) -> Tensor:
    """itermvs.py:341-350.  corr [B,G,N,H,W] -> [B,1,H,W]."""
    b, _, n, h, w = corr.shape
    x = m.conv[1](F.relu(m.conv[0].conv(_slices(corr)))).reshape(b, n, h, w)     # reshape: cuDNN may answer channels-last
    return torch.softmax(x, dim=1).max(dim=1)[0].unsqueeze(1)


def corr_net(m, corr: Tensor) -> Tensor:
    """itermvs.py:367-381.  corr [B,G,N,H,W] -> [B,N,H,W]."""
    b, _, n, h, w = corr.shape
    c0 = F.relu(m.conv0.conv(_slices(corr)))
    c1 = F.relu(m.conv1.conv(c0))
    x = F.relu(m.conv2.conv(c1))
    x = c1 + m.conv3(x)
    x = c0 + m.conv4(x)
    return m.conv5(x).reshape(b, n, h, w)


def conv_gru(m, h: Tensor, x: Tensor) -> Tensor:
    """module.py:59-66."""
    hx = torch.cat([h, x], dim=1)
    z = torch.sigmoid(m.convz(hx))
    r = torch.sigmoid(m.convr(hx))
    q = torch.tanh(m.convq(torch.cat([r * h, x], dim=1)))
    return (1 - z) * h + z * q


def window_regression(probability: Tensor, radius: int = 4) -> Tensor:
    """itermvs.py:173-189 / 203-219: arg-max bin, clamped +-radius window (edge bins counted twice when clamped),
    expectation over the window; the indices carry no gradient, the probabilities do."""
    bins = probability.shape[1]
    with torch.no_grad():
        top = torch.argmax(probability, dim=1, keepdim=True)
        offs = torch.arange(-radius, radius + 1, device=probability.device).view(1, -1, 1, 1)
        idx = torch.clamp(top + offs, min=0, max=bins - 1)
    p = torch.gather(probability, 1, idx)
    num = torch.zeros_like(p[:, :1])
    den = torch.full_like(p[:, :1], 1e-6)
    for i in range(2 * radius + 1):                 # same accumulation order as the reference's loop
        num = num + idx[:, i:i + 1] * p[:, i:i + 1]
        den = den + p[:, i:i + 1]
    return (num / den) / (bins - 1.0)


def _to_planar(vol: Tensor, h: int, w: int) -> Tensor:
    """kernel layout [..., N, P, 8] -> reference layout [..., 8, N, H, W]."""
    lead = vol.shape[:-3]
    n = vol.shape[-3]
    k = len(lead)
    return vol.reshape(*lead, n, h, w, 8).permute(*range(k), k + 3, k, k + 1, k + 2)


# ---- Evaluation (itermvs.py:33-126) on channels-last pyramids --------------------------------------------------
def evaluation_init(ev, fea3: Tensor, rt3: Tensor, depth_sample: Tensor, inv_min: Tensor, inv_max: Tensor):
    """The `view_weights == None` branch (itermvs.py:36-82): per-view correlation on the fused kernel, PixelViewWeight,
    aggregation (training form: no in-place adds), CorrNet, the soft-argmax initial depth.
    Returns (view_weights [B,S,H2,W2] -- NOT detached, as in the reference --, corr [B,D,H3,W3], depth [B,1,H2,W2])."""
    b, v, h3, w3, _ = fea3.shape
    d = depth_sample.shape[1]
    corr_views = _to_planar(FusedCorrInit.apply(fea3, rt3, depth_sample), h3, w3)      # [B,S,8,D,H3,W3]
    corr_sum, vw_sum, view_weights = 0, 1e-5, []
    for i in range(v - 1):
        c = corr_views[:, i]
        vw_i = pixel_view_weight(ev.pixel_view_weight, c)                              # [B,1,H3,W3]
        view_weights.append(F.interpolate(vw_i, scale_factor=2, mode="bilinear"))
        corr_sum = corr_sum + c * vw_i.unsqueeze(1)
        vw_sum = vw_sum + vw_i.unsqueeze(1)
    corr = corr_net(ev.corr_conv1[2], corr_sum / vw_sum)                                # [B,D,H3,W3]
    prob0 = torch.softmax(corr, dim=1)
    index = torch.arange(0, d, 1, device=corr.device, dtype=torch.float32).view(1, d, 1, 1)
    nd0 = torch.sum(index * prob0, dim=1, keepdim=True) / (d - 1.0)
    depth0 = F.interpolate(ops.depth_unnormalization(nd0, inv_min, inv_max), scale_factor=2, mode="bilinear")
    return torch.cat(view_weights, dim=1), corr, depth0


def evaluation_iter(ev, fea1: Tensor, fea2: Tensor, fea3: Tensor, rts: Sequence[Tensor], samples: Sequence[Tensor],
                    view_weights: Tensor) -> Tensor:
    """The iteration branch (itermvs.py:84-126): fused warp + correlation + view-weighted aggregation of the three
    levels, one CorrNet per level.  `view_weights` is used as a constant (the caller passes .detach(), itermvs.py:295).
    Returns corr [B,10,H2,W2]."""
    _, _, h2, w2, _ = fea2.shape
    agg = _to_planar(FusedCorrIter.apply(fea1, fea2, fea3, rts[0], rts[1], rts[2], samples[0].contiguous(), samples[1].contiguous(),
                                         samples[2].contiguous(), view_weights.detach().contiguous()), h2, w2)     # [B,8,10,H2,W2]
    return torch.cat([corr_net(ev.corr_conv1[0], agg[:, :, 0:4]), corr_net(ev.corr_conv1[1], agg[:, :, 4:8]),
                      corr_net(ev.corr_conv1[2], agg[:, :, 8:10])], dim=1)


def evaluation_forward(ev, ref_feature, src_features, ref_proj, src_projs, depth_sample, inverse_depth_min=None,
                       inverse_depth_max=None, view_weights=None):
    """Evaluation.forward with the reference's own arguments (dicts of NCHW maps, itermvs.py:33), differentiable."""
    def stack(level):
        maps = [ref_feature[level], *src_features[level]]
        fea = torch.stack(maps, dim=1).permute(0, 1, 3, 4, 2).contiguous()
        proj = torch.stack([ref_proj[level], *src_projs[level]], dim=1).float().to(fea.device)
        flag = ops.NanFlag(fea.device)
        rt = _compose(proj, flag)
        flag.raise_if_set()
        return fea, rt
    if view_weights is None:
        fea3, rt3 = stack("level3")
        return evaluation_init(ev, fea3, rt3, depth_sample, inverse_depth_min, inverse_depth_max)
    (f1, r1), (f2, r2), (f3, r3) = stack("level1"), stack("level2"), stack("level3")
    return evaluation_iter(ev, f1, f2, f3, (r1, r2, r3), [depth_sample[f"level{l}"] for l in (1, 2, 3)], view_weights)


# ---- Update (itermvs.py:159-220) ---------------------------------------------------------------------------------
def update_hidden_init(upd, corr: Tensor) -> Tensor:
    return torch.tanh(F.interpolate(upd.hidden_init_head(corr), scale_factor=2, mode="bilinear"))


def update_depth(upd, hidden: Tensor):
    probability = torch.softmax(upd.depth_head(hidden), dim=1)
    return window_regression(probability, upd.radius), probability


def update_forward(upd, hidden: Tensor, normalized_depth: Tensor, corr: Tensor, confidence_flag: bool):
    """Update.forward (itermvs.py:192-220): (hidden, normalized_depth, probability, confidence, confidence_logit)."""
    hidden = conv_gru(upd.gru, hidden, torch.cat([normalized_depth, corr], dim=1))
    conf0 = upd.confidence_head(hidden) if confidence_flag else None
    nd, probability = update_depth(upd, hidden)
    return hidden, nd, probability, (torch.sigmoid(conf0) if confidence_flag else None), conf0


# ---- the estimator, training structure (itermvs.py:253-329 with test=False) ------------------------------------
def itermvs_train_forward(net, fea1: Tensor, fea2: Tensor, fea3: Tensor, projs: Sequence[Tensor], depth_min: Tensor,
                          depth_max: Tensor):
    """net: itermvs_b200.estimator.IterMVS.  fea_l: channels-last pyramids [B,V,H_l,W_l,C_l]; projs: level 1..3
    projection stacks [B,V,4,4].  Returns (depths, depths_upsampled, confidences, confidence_upsampled) exactly as
    the reference's training forward does."""
    ev, upd = net.evaluation, net.update
    b, v, h2, w2, _ = fea2.shape
    s = v - 1
    h3, w3 = h2 // 2, w2 // 2
    dev = fea2.device
    depths: Dict[str, List[Tensor]] = {"combine": [], "probability": [], "initial": []}
    confidences: List[Tensor] = []
    depths_upsampled: List[Tensor] = []
    confidence_upsampled = None
    flag = ops.NanFlag(dev)
    rts = [_compose(p, flag) for p in projs]
    flag.raise_if_set()                      # "nan in proj" (module.py:87); the reference synchronises on every warp call

    ref2 = fea2[:, 0].permute(0, 3, 1, 2)                                   # NCHW view of the reference feature
    up_w = torch.softmax(net.upsample(ref2).reshape(b, 1, 9, 4, 4, h2, w2), dim=2)        # itermvs.py:262-264
    inv_min = (1.0 / depth_min).reshape(b, 1, 1, 1)
    inv_max = (1.0 / depth_max).reshape(b, 1, 1, 1)

    # ---- initialisation (itermvs.py:270-283)
    samples0 = net.depth_initialization(inv_min, inv_max, h3, w3, dev)
    view_weights, corr, depth0 = evaluation_init(ev, fea3, rts[2], samples0, inv_min, inv_max)
    depths["initial"].append(depth0)

    hidden = update_hidden_init(upd, corr)                                              # itermvs.py:275-279
    nd, probability = update_depth(upd, hidden)
    conf0 = upd.confidence_head(hidden)
    depths["combine"].append(ops.depth_unnormalization(nd, inv_min, inv_max))
    depths["probability"].append(probability)
    confidences.append(conf0)
    nd = nd.detach()

    # ---- iterations (itermvs.py:285-315)
    vw_const = view_weights.detach().contiguous()
    for it in range(net.iteration):
        smp = []
        for lvl in ("level1", "level2", "level3"):
            ns = torch.clamp(nd + net.corr_interval[lvl].to(dev) * net.interval_scale, min=0, max=1)
            smp.append(ops.depth_unnormalization(ns, inv_min, inv_max).contiguous())
        corr = evaluation_iter(ev, fea1, fea2, fea3, rts, smp, vw_const)
        hidden, nd, probability, conf, conf0 = update_forward(upd, hidden, nd, corr, True)
        depths["combine"].append(ops.depth_unnormalization(nd, inv_min, inv_max))
        depths["probability"].append(probability)
        confidences.append(conf0)
        if it == net.iteration - 1:
            depths_upsampled.append(ops.depth_unnormalization(ops.upsample(nd, up_w), inv_min, inv_max))
            confidence_upsampled = F.interpolate(conf, scale_factor=4, mode="bilinear")
        nd = nd.detach()
    return depths, depths_upsampled, confidences, confidence_upsampled


def pipeline_train_forward(model, imgs, proj_matrices, depth_min, depth_max):
    """model: itermvs_b200.Pipeline in train() mode.  Same output dict as net.py:115-120."""
    fea1, fea2, fea3 = featurenet_pyramids(model.feature_net, imgs["level_0"])
    projs = [proj_matrices[f"level_{l}"].float() for l in (1, 2, 3)]
    depths, depths_upsampled, confidences, confidence_upsampled = itermvs_train_forward(
        model.iter_mvs, fea1, fea2, fea3, projs, depth_min.float(), depth_max.float())
    return {"depths": depths, "depths_upsampled": depths_upsampled, "confidences": confidences,
            "confidence_upsampled": confidence_upsampled}
