"""GPU parity tests: every CUDA entry point (through the Python mirror of the reference interface,
which calls the C ABI) against the CPU oracle and the reference-generated golden fixtures.

Tolerances: the path is fp32 end to end; single operators are compared at 1e-5..1e-4 absolute
(fp32 re-association only), the end-to-end depth at the north star's 1e-3 relative.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import itermvs_oracle as O
from itermvs_b200.synthetic import make_sample, random_feature_pyramids

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


def maxerr(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


@pytest.fixture(scope="module")
def dev():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model(dev, dtu_weights):
    import itermvs_b200
    m = itermvs_b200.Pipeline(iteration=4, test=True)
    m.load_state_dict(dtu_weights, strict=True)
    return m.to(dev).eval()


def test_library_loaded_is_in_tree():
    from itermvs_b200 import _lib
    _lib.lib()
    assert _lib.library_path().endswith("itermvs_b200/csrc/libitermvs_b200.so")
    maps = open("/proc/self/maps").read()
    assert "libitermvs_b200.so" in maps


@pytest.mark.parametrize("tag", ["same", "fea2x", "fea_half", "b2"])
def test_differentiable_warping_golden(dev, stage_kats, tag):
    from itermvs_b200 import differentiable_warping
    k = stage_kats
    out = differentiable_warping(T(k[f"warp_{tag}_fea"]).to(dev), T(k[f"warp_{tag}_src_proj"]).to(dev),
                                 T(k[f"warp_{tag}_ref_proj"]).to(dev), T(k[f"warp_{tag}_depth"]).to(dev))
    assert out.shape == k[f"warp_{tag}_out"].shape
    # white-noise features: |d fea / d px| ~ 3, fp32 position error ~2e-5 px  =>  ~1e-4 worst case
    assert maxerr(out, T(k[f"warp_{tag}_out"])) < 2e-4


def test_differentiable_warping_backward_vs_oracle_autograd(dev, stage_kats):
    """CUDA backward (grad w.r.t. src_fea; the grid carries none, module.py:77) against torch autograd through the
    oracle's explicit 4-tap restatement, on the golden inputs incl. the z <= 0.01 pixels and at the three
    resolution ratios; plus linearity in grad_out."""
    from itermvs_b200 import differentiable_warping
    k = stage_kats
    for tag in ("same", "fea2x", "fea_half", "b2"):
        fea = T(k[f"warp_{tag}_fea"])
        sp, rp, dep = T(k[f"warp_{tag}_src_proj"]), T(k[f"warp_{tag}_ref_proj"]), T(k[f"warp_{tag}_depth"])
        g = torch.Generator().manual_seed(5)
        gout = torch.randn(k[f"warp_{tag}_out"].shape, generator=g)
        f_cpu = fea.clone().requires_grad_(True)
        (O.differentiable_warping(f_cpu, sp, rp, dep) * gout).sum().backward()
        f_gpu = fea.to(dev).requires_grad_(True)
        out = differentiable_warping(f_gpu, sp.to(dev), rp.to(dev), dep.to(dev))
        assert out.requires_grad
        (out * gout.to(dev)).sum().backward()
        scale = float(f_cpu.grad.abs().max())
        assert maxerr(f_gpu.grad, f_cpu.grad) < 2e-4 * max(scale, 1.0), tag
        # taps that fall outside the map receive nothing: total mass matches the oracle's
        assert abs(float(f_gpu.grad.sum()) - float(f_cpu.grad.sum())) < 1e-2 * max(1.0, float(f_cpu.grad.abs().sum()) * 1e-3)
    # no gradient is requested for / delivered to the projections and the depth samples
    dep_g = dep.to(dev).requires_grad_(True)
    f_gpu = fea.to(dev).requires_grad_(True)
    differentiable_warping(f_gpu, sp.to(dev), rp.to(dev), dep_g).sum().backward()
    assert dep_g.grad is None


def test_differentiable_warping_nan_assert(dev):
    from itermvs_b200 import differentiable_warping
    fea = torch.randn(1, 16, 8, 8, device=dev)
    proj = torch.eye(4, device=dev)[None].clone()
    bad = proj.clone()
    bad[0, 0, 0] = float("nan")
    with pytest.raises(AssertionError):
        differentiable_warping(fea, proj, bad, torch.full((1, 2, 8, 8), 500.0, device=dev))


def test_compose_and_layout(dev):
    from itermvs_b200 import compose_projections, nchw_to_nhwc, nhwc_to_nchw
    s = make_sample(160, 128, n_src=3, batch=2, seed=5, scene="noise")
    proj = s["proj_matrices"]["level_2"].float()
    rt = compose_projections(proj.to(dev)).cpu()
    for b in range(2):
        for v in range(3):
            m = proj[b, v + 1].double() @ torch.inverse(proj[b, 0].double())
            ref = torch.cat([m[:3, :3].reshape(-1), m[:3, 3]]).float()
            assert torch.allclose(rt[b, v], ref, rtol=2e-6, atol=1e-6)
    x = torch.randn(3, 48, 17, 23, device=dev)
    y = nchw_to_nhwc(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(nhwc_to_nchw(y), x)


def test_corrnet_pvw_gru_hinit_golden(dev, stage_kats, model):
    k = stage_kats
    ev, upd = model.iter_mvs.evaluation, model.iter_mvs.update
    for i in range(3):
        out = ev.corr_conv1[i](T(k["corrnet_in"]).to(dev))
        assert maxerr(out, T(k[f"corrnet{i}_out"])) < 2e-5
    assert maxerr(ev.pixel_view_weight(T(k["pvw_in"]).to(dev)), T(k["pvw_out"])) < 1e-5
    assert maxerr(upd.gru(T(k["gru_h"]).to(dev), T(k["gru_x"]).to(dev)), T(k["gru_out"])) < 1e-5
    assert maxerr(upd.hidden_init(T(k["hinit_in"]).to(dev)), T(k["hinit_out"])) < 2e-5


def test_heads_probability_and_window_regression(dev, stage_kats, model, dtu_weights):
    upd = model.iter_mvs.update
    upd.return_probability = True
    h = T(stage_kats["gru_h"]).to(dev)
    try:
        nd, prob = upd.depth_init(h)
        conf, conf0 = upd.conf_init(h)
    finally:
        upd.return_probability = None
    prob_ref = torch.softmax(T(stage_kats["head_logits"]), dim=1)
    # fp32 softmax: the 256-term sum is order dependent (ATen sums sequentially, the kernel as a tree): ~1e-5
    assert maxerr(prob, prob_ref) < 3e-5      # sequential fp32 sum of 254 tiny terms on the CPU side loses ~1e-5
    assert maxerr(conf0, T(stage_kats["conf_logit"])) < 2e-5
    assert maxerr(conf, torch.sigmoid(T(stage_kats["conf_logit"]))) < 1e-5
    # arg-max on a near-flat random distribution is chaotic: check the window regression
    # against the oracle's regression applied to the SAME probabilities
    assert maxerr(nd, O.window_regression(prob.cpu())) < 2e-6


def test_window_regression_edges(dev, model):
    """Clamped +-4 window with duplicate edge bins (itermvs.py:203-219) incl. arg-max at 0,1,254,255:
    drive the head through weights that make logit c = 20*[c == target(px)] exactly."""
    import copy
    upd = copy.deepcopy(model.iter_mvs.update)
    with torch.no_grad():
        for p in upd.parameters():
            p.zero_()
        # pixel j carries a one-hot hidden state on channel j; the centre taps of conv0 and fc1 copy it,
        # fc2 turns it into logits 12 at bin target_j and 9 at bin target_j + 1 (small integers: exact)
        targets = [0, 1, 2, 3, 4, 5, 100, 250, 251, 252, 253, 254, 255, 17, 64, 200]
        for j, tj in enumerate(targets):
            upd.depth_head[0].weight[j, j, 1, 1] = 1.0
            upd.depth_head[2].weight[j, j, 0, 0] = 1.0
            upd.depth_head[4].weight[tj, j, 0, 0] = 12.0
            upd.depth_head[4].weight[min(tj + 1, 255), j, 0, 0] += 9.0
    h = torch.zeros(1, 32, 2, 8)
    for j in range(16):
        h[0, j, j // 8, j % 8] = 1.0
    upd.return_probability = True
    nd, prob = upd.to(dev).depth_init(h.to(dev))
    w = {"iter_mvs.update." + k: v.detach().cpu() for k, v in upd.state_dict().items()}
    nd_ref, prob_ref = O.depth_init(w, h)
    assert maxerr(prob, prob_ref) < 3e-5      # sequential fp32 sum of 254 tiny terms on the CPU side loses ~1e-5
    assert torch.equal(prob.argmax(1).cpu(), prob_ref.argmax(1))
    assert maxerr(nd, nd_ref) < 1e-6


def test_fused_tcgen05_head(dev, model, e2e_d32):
    """The default head (csrc/headfused.cuh: fc1 + fc2 + softmax / arg-max / window regression in one tcgen05 kernel, logits in
    TMEM) against (a) the reference's own Update outputs on the reference's own hidden states (e2e_d32.npz) and (b) the unfused
    kernels (IMVS_TUNE_HEADFUSED=0) on the same input; incl. the confidence branch and a ragged last 128-pixel tile."""
    import os
    from itermvs_b200 import _lib
    assert _lib.get_conv_passes() == 4
    upd = model.iter_mvs.update
    fix = e2e_d32
    its = sorted(int(k[len("hidden_iter"):]) for k in fix if k.startswith("hidden_iter"))
    assert its
    for it in its:
        h = T(fix[f"hidden_iter{it}"]).to(dev)
        want = fix[f"nd_iter{it}"]
        nd_f, prob = upd.depth_init(h)
        assert prob is None                                     # no probability requested -> the fused kernel ran
        conf_f, conf0_f = upd.conf_init(h)
        os.environ["IMVS_TUNE_HEADFUSED"] = "0"
        try:
            nd_u, _ = upd.depth_init(h)
            conf_u, conf0_u = upd.conf_init(h)
        finally:
            del os.environ["IMVS_TUNE_HEADFUSED"]
        err_f = np.abs(nd_f.cpu().numpy() - want)
        err_u = np.abs(nd_u.cpu().numpy() - want)
        d_fu = (nd_f - nd_u).abs()
        print(f"head iter {it}: fused vs reference max {err_f.max():.2e} (>1e-4: {100 * (err_f > 1e-4).mean():.4f}%), unfused vs reference "
              f"max {err_u.max():.2e}, fused vs unfused max {float(d_fu.max()):.2e}, conf logit fused vs unfused {float((conf0_f - conf0_u).abs().max()):.2e}")
        assert (err_f > 1e-4).mean() < 1e-3 and np.median(err_f) < 1e-6
        assert float((d_fu > 1e-4).float().mean()) < 1e-3
        assert float((conf0_f - conf0_u).abs().max()) < 1e-4 and float((conf_f - conf_u).abs().max()) < 1e-5
    # ragged tile: 5 x 7 pixels (35 < 128)
    h = torch.randn(2, 32, 5, 7, device=dev).tanh()
    nd_f, _ = upd.depth_init(h)
    os.environ["IMVS_TUNE_HEADFUSED"] = "0"
    try:
        nd_u, _ = upd.depth_init(h)
    finally:
        del os.environ["IMVS_TUNE_HEADFUSED"]
    assert float(((nd_f - nd_u).abs() > 1e-4).float().mean()) < 0.05           # near-flat random distributions: rare arg-max flips
    assert model.iter_mvs.update is upd and _lib.lib().imvs_tcgen05_status() == 0


def _feature_inputs(dev, width, height, n_src, batch, seed):
    ref, srcs = random_feature_pyramids(width, height, n_src, batch, seed)
    s = make_sample(width, height, n_src=n_src, batch=batch, seed=seed, scene="noise")
    rp, sp = {}, {}
    for l in (1, 2, 3):
        pm = torch.unbind(s["proj_matrices"][f"level_{l}"].float(), 1)
        rp[f"level{l}"], sp[f"level{l}"] = pm[0], list(pm[1:])
    to = lambda d: {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in d.items()}
    return (ref, srcs, rp, sp, s), (to(ref), to(srcs), to(rp), to(sp))


@pytest.mark.parametrize("batch,n_src", [(1, 2), (2, 3), (1, 1)])
def test_evaluation_init_branch(dev, model, dtu_weights, batch, n_src):
    (ref, srcs, rp, sp, s), (gref, gsrcs, grp, gsp) = _feature_inputs(dev, 160, 128, n_src, batch, seed=11)
    inv_min = (1.0 / s["depth_min"]).view(batch, 1, 1, 1)
    inv_max = (1.0 / s["depth_max"]).view(batch, 1, 1, 1)
    ds = O.initial_depth_samples(inv_min, inv_max, 32, 16, 20)
    ds[:, 3, :2] = -10.0                                             # exercise the z <= 0.01 substitution
    want = O.evaluation_init(dtu_weights, ref["level3"], srcs["level3"], rp["level3"], sp["level3"], ds, inv_min, inv_max)
    vw, corr, depth = model.iter_mvs.evaluation(gref, gsrcs, grp, gsp, ds.to(dev), inv_min.to(dev), inv_max.to(dev))
    assert maxerr(vw, want["view_weights"]) < 2e-5
    assert maxerr(corr, want["corr"]) < 1e-4
    assert maxerr(depth, want["depth"]) / 600.0 < 1e-4


# n_src 1..8 run the kernels unrolled per view count, 9 the rolled-loop instantiation (up to 16 views)
@pytest.mark.parametrize("batch,n_src", [(1, 4), (2, 3), (1, 1), (1, 7), (2, 9)])
def test_evaluation_iter_branch(dev, model, dtu_weights, batch, n_src):
    (ref, srcs, rp, sp, s), (gref, gsrcs, grp, gsp) = _feature_inputs(dev, 160, 128, n_src, batch, seed=12)
    g = torch.Generator().manual_seed(3)
    inv_min = (1.0 / s["depth_min"]).view(batch, 1, 1, 1)
    inv_max = (1.0 / s["depth_max"]).view(batch, 1, 1, 1)
    nd = torch.rand(batch, 1, 32, 40, generator=g)
    nd[:, :, 0, :4] = torch.tensor([0.0, 1.0, 0.001, 0.999])      # clamp at both ends of the range
    samples = {f"level{l}": O.iteration_depth_samples(nd, l, inv_min, inv_max) for l in (1, 2, 3)}
    vw = torch.rand(batch, n_src, 32, 40, generator=g)
    want = O.evaluation_iter(dtu_weights, ref, srcs, rp, sp, samples, vw)
    got = model.iter_mvs.evaluation(gref, gsrcs, grp, gsp, {k: v.to(dev) for k, v in samples.items()},
                                    view_weights=vw.to(dev))
    assert maxerr(got, want) < 1e-4


def test_upsample_outputs(dev, model, dtu_weights):
    from itermvs_b200 import _lib, ops
    g = torch.Generator().manual_seed(9)
    b, h2, w2 = 2, 16, 24
    ref2 = torch.randn(b, 32, h2, w2, generator=g)
    nd = torch.rand(b, 1, h2, w2, generator=g)
    conf = torch.rand(b, 1, h2, w2, generator=g)
    dmin, dmax = torch.tensor([425.0, 300.0]), torch.tensor([935.0, 800.0])
    inv_min, inv_max = (1 / dmin).view(b, 1, 1, 1), (1 / dmax).view(b, 1, 1, 1)
    want_d = O.depth_unnormalization(O.convex_upsample(nd, O.upsample_weights(dtu_weights, ref2)), inv_min, inv_max)
    want_c = O.bilinear_up(conf, 4)
    wts = model.iter_mvs.packed(dev)
    d_up = torch.empty(b, 1, 4 * h2, 4 * w2, device=dev)
    c_up = torch.empty_like(d_up)
    scratch = torch.empty(b * 64 * h2 * w2, device=dev)
    gd = lambda t: t.to(dev).contiguous()
    ref2g, ndg, confg, dming, dmaxg = gd(ref2.permute(0, 2, 3, 1)), gd(nd), gd(conf), gd(dmin), gd(dmax)   # feature channels-last
    _lib.check(_lib.lib().imvs_upsample_outputs(wts.ref, ref2g.data_ptr(), h2 * w2 * 32, ndg.data_ptr(), h2 * w2, 1,
                                                confg.data_ptr(), dming.data_ptr(), dmaxg.data_ptr(), d_up.data_ptr(),
                                                c_up.data_ptr(), scratch.data_ptr(), b, h2, w2, ops._stream()))
    assert maxerr(d_up, want_d) / 500.0 < 2e-6
    assert maxerr(c_up, want_c) < 1e-6


def _frac_bad(a, b, tol):
    a, b = a.detach().cpu().double().numpy(), np.asarray(b, np.float64)
    return float((np.abs(a - b) > tol * np.maximum(np.abs(b), 1e-6)).mean())


@pytest.mark.parametrize("which", ["e2e_d8", "e2e_d32"])
def test_pipeline_matches_reference_fixture(dev, dtu_weights, request, which):
    """End to end against outputs of the reference itself (tests/golden/make_golden.py)."""
    import itermvs_b200
    fix = request.getfixturevalue(which)
    d = int(fix["num_sample"])
    m = itermvs_b200.Pipeline(iteration=int(fix["iteration"]), test=True)
    if d != 32:
        m.iter_mvs.update.hidden_init_head[0] = torch.nn.Conv2d(d, 64, 3, stride=1, padding=1, bias=False)
    sd = dict(dtu_weights)
    for k, v in fix.items():
        if k.startswith("extra:"):
            sd[k[6:]] = T(v)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    s = make_sample(int(fix["width"]), int(fix["height"]), n_src=int(fix["n_src"]), batch=1, seed=int(fix["seed"]), scene="plane")
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    with torch.no_grad():
        out = m(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    torch.cuda.synchronize()
    m._last_nan_flag.raise_if_set()
    du, cu_ = out["depths_upsampled"], out["confidence_upsampled"]
    assert du.shape == fix["depths_upsampled"].shape
    bad = _frac_bad(du, fix["depths_upsampled"], 1e-3)
    med = float(np.median(np.abs(du.cpu().numpy() - fix["depths_upsampled"]) / fix["depths_upsampled"]))
    print(f"{which}: depth rel err > 1e-3 on {100 * bad:.3f}% px, median rel err {med:.2e}")
    # measured on B200 (round 2): 0 % of pixels beyond 1e-3 in both cases, median 0 / 9.1e-8
    assert med < 1e-6
    assert bad < 1e-4
    assert float((np.abs(cu_.cpu().numpy() - fix["confidence_upsampled"]) > 1e-3).mean()) < 1e-3


def test_all_predictions_forward_matches_reference(dev, dtu_weights, e2e_allpred):
    """Pipeline(test=False).eval() under no_grad -- train.py's validation pass -- against the reference's own
    outputs: initial depth, every intermediate depth / probability / confidence logit, the upsampled outputs,
    and full_loss evaluated on them."""
    import itermvs_b200
    from itermvs_b200.synthetic import plane_depth_map
    fix = e2e_allpred
    w, h = int(fix["width"]), int(fix["height"])
    m = itermvs_b200.Pipeline(iteration=int(fix["iteration"]), test=False)
    m.load_state_dict(dtu_weights, strict=True)
    m = m.to(dev).eval()
    s = make_sample(w, h, n_src=int(fix["n_src"]), batch=1, seed=int(fix["seed"]), scene="plane")
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    with torch.no_grad():
        out = m(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    assert set(out) == {"depths", "depths_upsampled", "confidences", "confidence_upsampled"}
    assert set(out["depths"]) == {"combine", "probability", "initial"}
    n_pred = int(fix["iteration"]) + 1
    assert len(out["depths"]["combine"]) == n_pred and len(out["depths"]["probability"]) == n_pred
    assert len(out["confidences"]) == n_pred and len(out["depths_upsampled"]) == 1
    rel = lambda a, b: np.abs(a.cpu().numpy() - b) / np.abs(b)
    assert np.median(rel(out["depths"]["initial"][0], fix["depth_initial"])) < 1e-5
    assert (rel(out["depths"]["initial"][0], fix["depth_initial"]) > 1e-3).mean() < 1e-3
    for i in range(n_pred):
        r = rel(out["depths"]["combine"][i], fix[f"combine{i}"])
        print(f"all-predictions combine{i}: median rel {np.median(r):.2e}, px>1e-3 {100 * (r > 1e-3).mean():.4f}%")
        assert np.median(r) < 1e-6 and (r > 1e-3).mean() < 1e-4, (i, np.median(r), (r > 1e-3).mean())     # measured: 0, 0 %
        p = out["depths"]["probability"][i]
        assert p.shape == (1, 256, h // 4, w // 4)
        assert float((p.sum(1) - 1).abs().max()) < 1e-4
        same_bin = (p.argmax(1).cpu().numpy() == fix[f"probability{i}_argmax"]).mean()
        print(f"all-predictions probability{i}: same arg-max bin {100 * same_bin:.4f}%")
        assert same_bin > 0.9999, (i, same_bin)                                                          # measured: 100 %
        dp = np.abs(p[:, :, ::4, ::4].cpu().numpy() - fix[f"probability{i}_s4"])
        assert np.median(dp.max(axis=1)) < 1e-4
        dc = np.abs(out["confidences"][i].cpu().numpy() - fix[f"confidence_logit{i}"])
        print(f"all-predictions confidence logit{i}: median abs err {np.median(dc):.2e}, >5e-2: {100 * (dc > 5e-2).mean():.4f}%")
        assert np.median(dc) < 1e-4 and (dc > 5e-2).mean() < 1e-4, (i, np.median(dc))                    # measured: 2e-6, 0 %
    r = rel(out["depths_upsampled"][0], fix["depths_upsampled"])
    assert np.median(r) < 1e-6 and (r > 1e-3).mean() < 1e-4
    assert (np.abs(out["confidence_upsampled"].cpu().numpy() - fix["confidence_upsampled"]) > 1e-3).mean() < 1e-3
    # the loss of the reference on the reference's outputs vs our loss on our outputs (same gt / masks as the generator)
    d0 = torch.from_numpy(plane_depth_map(w, h).astype(np.float32))[None, None].to(dev)
    gt = {"level_0": d0, "level_2": torch.nn.functional.interpolate(d0, scale_factor=0.25, mode="nearest")}
    mask = {k: torch.ones_like(v) for k, v in gt.items()}
    mask["level_0"][..., :6, :] = 0
    mask["level_2"][..., :2, :] = 0
    loss = float(itermvs_b200.full_loss(out["depths"], out["depths_upsampled"], out["confidences"], gt, mask,
                                        s["depth_min"].to(dev), s["depth_max"].to(dev)))
    print(f"all-predictions forward: loss {loss:.5f} (reference {float(fix['loss']):.5f})")
    assert abs(loss - float(fix["loss"])) < 2e-3 * float(fix["loss"])
    # eval() with autograd on (fine-tuning with frozen BatchNorm statistics, as the reference allows) runs the
    # differentiable path: same predictions as the forward-only kernels, and gradients reach the parameters
    out_g = m(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    rel_g = ((out_g["depths_upsampled"][0] - out["depths_upsampled"][0]).abs() / out["depths_upsampled"][0]).median()
    assert float(rel_g) < 1e-3, float(rel_g)          # cuDNN convolutions (possibly TF32) vs the fp32-grade kernels
    itermvs_b200.full_loss(out_g["depths"], out_g["depths_upsampled"], out_g["confidences"], gt, mask,
                           s["depth_min"].to(dev), s["depth_max"].to(dev)).backward()
    gw = m.iter_mvs.update.gru.convz.weight.grad
    assert gw is not None and torch.isfinite(gw).all() and float(gw.abs().sum()) > 0
    m.zero_grad(set_to_none=True)


def test_streaming_two_in_flight_is_race_free(dev, model):
    """graph.StreamingPipeline with two reference views in flight (own workspace, static buffers and compute
    stream per slot): every result must equal the plain forward of the same inputs, bit for bit."""
    from itermvs_b200.graph import StreamingPipeline
    samples = [make_sample(320, 256, n_src=3, batch=1, seed=sd, scene="plane") for sd in (11, 12, 13)]
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    want = []
    with torch.no_grad():
        for s in samples:
            o = model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
            want.append((o["depths_upsampled"].cpu(), o["confidence_upsampled"].cpu()))
    s0 = samples[0]
    sp = StreamingPipeline(model, cu(s0["imgs"]), cu(s0["proj_matrices"]), s0["depth_min"].to(dev), s0["depth_max"].to(dev))
    host = [({"level_0": s["imgs"]["level_0"].pin_memory()}, {k: v.float().pin_memory() for k, v in s["proj_matrices"].items()
                                                             if k in ("level_1", "level_2", "level_3")},
             s["depth_min"].pin_memory(), s["depth_max"].pin_memory()) for s in samples]
    n = 12
    outs = [(torch.empty(1, 1, 256, 320).pin_memory(), torch.empty(1, 1, 256, 320).pin_memory()) for _ in range(n)]
    for k in range(n):
        sp.submit(*host[k % 3], *outs[k])
    sp.drain()
    torch.cuda.synchronize()
    for k in range(n):
        assert torch.equal(outs[k][0], want[k % 3][0]), f"depth of streamed sample {k} differs"
        assert torch.equal(outs[k][1], want[k % 3][1]), f"confidence of streamed sample {k} differs"


def test_full_size_pipeline_vs_oracle(dev, model, dtu_weights):
    """BASELINE config 2 (640x512, 4 src, D=32, 4 iterations) on the consistent plane scene:
    depth within 1e-3 relative of the oracle (north star tolerance), stage traces tighter."""
    s = make_sample(640, 512, n_src=4, batch=1, seed=0, scene="plane")
    trace = {}
    want = O.pipeline_forward(dtu_weights, s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"], iteration=4, trace=trace)
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    with torch.no_grad():
        out = model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    torch.cuda.synchronize()
    d, dref = out["depths_upsampled"].cpu(), want["depths_upsampled"]
    rel = ((d - dref).abs() / dref).numpy()
    c, cref = out["confidence_upsampled"].cpu(), want["confidence_upsampled"]
    print(f"config2: depth L1 {float((d - dref).abs().mean()):.3e} mm, rel err mean {rel.mean():.2e} max {rel.max():.2e}, "
          f"px>1e-3: {100 * (rel > 1e-3).mean():.4f}%  conf max err {float((c - cref).abs().max()):.2e}")
    # measured on B200 (round 2): mean 8.2e-8, max 2.9e-6, confidence max error 6.1e-5
    assert rel.mean() < 1e-6
    assert rel.max() < 1e-4
    assert float((c - cref).abs().max()) < 1e-3


def test_uint8_images_equal_loader_normalisation(dev, model):
    """f-4: raw 8-bit images straight into the pipeline (a quarter of the H2D bytes); the first FeatureNet layer normalises
    them as the reference's loaders do -- 2 * np.array(img, dtype=np.float32) / 255. - 1 (datasets/dtu_yao_eval.py:63-64) -- so
    the result is bit-identical to feeding that float image."""
    s = make_sample(320, 256, n_src=3, batch=1, seed=21, scene="plane")
    u8 = ((s["imgs"]["level_0"] + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
    as_float = torch.from_numpy(2 * u8.numpy().astype(np.float32) / 255. - 1)       # the loader's arithmetic, on the host
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    with torch.no_grad():
        a = model({"level_0": u8.to(dev)}, cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
        a = {k: v.clone() for k, v in a.items()}
        b = model({"level_0": as_float.to(dev)}, cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    assert torch.equal(a["depths_upsampled"], b["depths_upsampled"]) and torch.equal(a["confidence_upsampled"], b["confidence_upsampled"])
    # and through the serving loop from pinned uint8 host buffers
    from itermvs_b200.graph import StreamingPipeline
    proj = {k: s["proj_matrices"][k].float() for k in ("level_1", "level_2", "level_3")}
    sp = StreamingPipeline(model, {"level_0": u8.to(dev)}, cu(proj), s["depth_min"].to(dev), s["depth_max"].to(dev), n_slots=2)
    outs = [(torch.empty(1, 1, 256, 320).pin_memory(), torch.empty(1, 1, 256, 320).pin_memory()) for _ in range(2)]
    host = {"level_0": u8.pin_memory()}
    for k in range(4):
        sp.submit(host, {k_: v.pin_memory() for k_, v in proj.items()}, s["depth_min"].pin_memory(), s["depth_max"].pin_memory(), *outs[k % 2])
    sp.drain(check_nan=True)
    torch.cuda.synchronize()
    assert torch.equal(outs[1][0].to(dev), b["depths_upsampled"])


def test_fp16_range_guard(dev, model):
    """Mode 4 splits activations into fp16 hi + lo; beyond +-65504 the split saturates and the result would be a finite wrong
    number.  Every convolution checks its accumulators and raises bit 1 of the device status word -- the silent failure
    mode is loud: images scaled by 1e6 trip it, normal inputs do not."""
    from itermvs_b200 import _lib
    s = make_sample(320, 256, n_src=2, batch=1, seed=31, scene="plane")
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    _lib.device_status(clear=True)
    with torch.no_grad():
        model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    assert _lib.device_status(clear=True) == 0
    _lib.check_device_status()
    big = {"level_0": (s["imgs"]["level_0"] * 1e6).to(dev)}
    with torch.no_grad():
        model(big, cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    assert _lib.device_status(clear=False) & 2
    with pytest.raises(OverflowError):
        _lib.check_device_status()
    assert _lib.device_status(clear=True) == 0


def _golden(name):
    import os
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)) as z:
        return {k: z[k] for k in z.files}


def test_cfg2_matches_reference_fixture(dev, model):
    """BASELINE configs[1] (640x512, 4 src, D=32, 4 iterations) against outputs of THE REFERENCE ITSELF
    (tests/golden/make_golden_cfg.py -> e2e_cfg2.npz): the north star's bar -- every pixel of depth within 1e-3
    relative -- asserted as a maximum, not a fraction."""
    fix = _golden("e2e_cfg2.npz")
    s = make_sample(640, 512, n_src=4, batch=1, seed=0, scene="plane")
    chk = np.array([float(s["imgs"]["level_0"].double().sum()), float(s["imgs"]["level_0"].double().abs().sum())])
    assert np.allclose(chk, fix["img_checksum"], rtol=1e-9), "synthetic generator drifted from the fixture's inputs"
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    with torch.no_grad():
        out = model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    torch.cuda.synchronize()
    d, c = out["depths_upsampled"].cpu().numpy(), out["confidence_upsampled"].cpu().numpy()
    rel = np.abs(d - fix["depths_upsampled"]) / fix["depths_upsampled"]
    cerr = np.abs(c - fix["confidence_upsampled"])
    print(f"cfg2 vs reference: depth rel err max {rel.max():.2e} mean {rel.mean():.2e}; confidence abs err max {cerr.max():.2e}")
    assert rel.max() < 1e-3, rel.max()
    assert rel.mean() < 1e-6
    assert cerr.max() < 1e-3, cerr.max()


def test_cfg5_full_size_matches_reference_fixture(dev, dtu_weights):
    """BASELINE configs[4] at FULL size (1920x1056, 7 source views, 4 iterations) against the reference itself
    (e2e_cfg5.npz, every 4th pixel of every 4th row).  D=32 is the checkpoint's configuration: every sampled pixel
    within 1e-3.  D=48 needs a re-created (random, stored) hidden_init_head[0]: the estimator then runs
    outside its trained regime (the reference's own mean confidence is 0.06), arg-max bins of flat distributions flip
    at fp32-reassociation level, so that case is median-tight with a bounded fraction of moved pixels."""
    import itermvs_b200
    fix = _golden("e2e_cfg5.npz")
    s = make_sample(1920, 1056, n_src=7, batch=1, seed=3, scene="plane")
    chk = np.array([float(s["imgs"]["level_0"].double().sum()), float(s["imgs"]["level_0"].double().abs().sum())])
    assert np.allclose(chk, fix["img_checksum"], rtol=1e-9)
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    for D in (32, 48):
        m = itermvs_b200.Pipeline(iteration=4, test=True)
        sd = dict(dtu_weights)
        if D != 32:
            m.iter_mvs.update.hidden_init_head[0] = torch.nn.Conv2d(D, 64, 3, stride=1, padding=1, bias=False)
            sd["iter_mvs.update.hidden_init_head.0.weight"] = T(fix[f"hidden_init_head0_d{D}"])
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).eval()
        with torch.no_grad():
            out = m(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
        torch.cuda.synchronize()
        m._last_nan_flag.raise_if_set()
        d = out["depths_upsampled"].cpu().numpy()[..., ::4, ::4]
        c = out["confidence_upsampled"].cpu().numpy()[..., ::4, ::4]
        want_d, want_c = fix[f"depths_upsampled_d{D}"], fix[f"confidence_upsampled_d{D}"]
        rel = np.abs(d - want_d) / want_d
        cerr = np.abs(c - want_c)
        print(f"cfg5 D={D} vs reference: depth rel err max {rel.max():.2e} median {np.median(rel):.2e}, px>1e-3 {100 * (rel > 1e-3).mean():.4f}%; "
              f"confidence abs err max {cerr.max():.2e}")
        if D == 32:
            assert rel.max() < 1e-3, rel.max()           # measured on B200: max 8.3e-5
            # the confidence is a sigmoid of a head on the hidden state: a handful of pixels with an ambiguous distribution
            # amplify fp32-reassociation noise beyond 1e-3 absolute (measured: max 1.6e-3, the reference's own CPU vs GPU
            # runs differ alike); bound the maximum and the fraction
            assert cerr.max() < 5e-3 and (cerr > 1e-3).mean() < 5e-4, (cerr.max(), (cerr > 1e-3).mean())   # measured 1.7e-4
        else:
            assert np.median(rel) < 1e-5 and (rel > 1e-3).mean() < 0.2
        del m
        torch.cuda.empty_cache()


def test_reference_itself_on_this_gpu(dev, model):
    """The unmodified reference (baseline/_ref, tools/install_ref.py) executed on this GPU through its stock ATen/cuDNN
    path (TF32 off, so that it is the fp32 computation its CPU path does) against this path, same inputs, at the
    benchmark configuration: every pixel within 1e-3."""
    from oracle import reference_arm as RA
    if not RA.available():
        pytest.skip(RA.why_unavailable())
    ref = RA.load_pipeline(iteration=4).to(dev)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        s = make_sample(640, 512, n_src=4, batch=1, seed=5, scene="plane")
        cu = lambda x: {k: v.to(dev) for k, v in x.items()}
        with torch.no_grad():
            want = ref(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
            out = model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
        torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    rel = ((out["depths_upsampled"] - want["depths_upsampled"]).abs() / want["depths_upsampled"])
    cerr = (out["confidence_upsampled"] - want["confidence_upsampled"]).abs()
    print(f"reference on this GPU (fp32 cuDNN) vs this path: depth rel err max {float(rel.max()):.2e} mean {float(rel.mean()):.2e}; "
          f"confidence abs err max {float(cerr.max()):.2e}")
    assert float(rel.max()) < 1e-3 and float(cerr.max()) < 1e-3


def test_size_independent_properties(dev, model):
    """Properties that hold at any size: (1) batch elements are independent (replicas) -- a batch of two
    different scenes equals the two run alone, bit for bit; (2) all outputs are finite and inside
    [depth_min, depth_max]; (3) the call is deterministic."""
    s2 = make_sample(320, 256, n_src=3, batch=2, seed=4, scene="plane")
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    with torch.no_grad():
        both = model(cu(s2["imgs"]), cu(s2["proj_matrices"]), s2["depth_min"].to(dev), s2["depth_max"].to(dev))
        both = {k: v.clone() for k, v in both.items()}
        again = model(cu(s2["imgs"]), cu(s2["proj_matrices"]), s2["depth_min"].to(dev), s2["depth_max"].to(dev))
        assert torch.equal(both["depths_upsampled"], again["depths_upsampled"])
        for b in range(2):
            one = model({k: v[b:b + 1].to(dev) for k, v in s2["imgs"].items()},
                        {k: v[b:b + 1].to(dev) for k, v in s2["proj_matrices"].items()},
                        s2["depth_min"][b:b + 1].to(dev), s2["depth_max"][b:b + 1].to(dev))
            # every kernel tiles per image and reduces in a fixed order: a batch element is bit-identical to itself alone
            nd_ = int((one["depths_upsampled"] != both["depths_upsampled"][b:b + 1]).sum())
            print(f"batch element {b}: {nd_} of {one['depths_upsampled'].numel()} depth values differ from the batch-of-one run")
            assert torch.equal(one["depths_upsampled"], both["depths_upsampled"][b:b + 1])
            assert torch.equal(one["confidence_upsampled"], both["confidence_upsampled"][b:b + 1])
    d = both["depths_upsampled"]
    assert torch.isfinite(d).all() and float(d.min()) >= 425.0 - 1e-2 and float(d.max()) <= 935.0 + 1e-2
    c = both["confidence_upsampled"]
    assert torch.isfinite(c).all() and float(c.min()) >= 0 and float(c.max()) <= 1


def test_featurenet_vs_oracle(dev, model, dtu_weights):
    """FeatureNet on the tensor-core conv kernels (BN folded, 3xTF32 split) against the fp32 CPU oracle."""
    s = make_sample(320, 256, n_src=2, batch=1, seed=6, scene="plane")
    x = s["imgs"]["level_0"]
    got = model.feature_net(x.to(dev))                     # reference return format: level -> list of NCHW views
    for v in range(3):
        want = O.feature_net(dtu_weights, x[:, v])
        for lvl in ("level1", "level2", "level3"):
            err = maxerr(got[lvl][v], want[lvl])
            scale = float(want[lvl].abs().max())
            assert err < 2e-5 * max(scale, 1.0), (lvl, v, err, scale)


def test_single_pass_tf32_mode(dev, model, dtu_weights):
    """conv_passes=1 (plain TF32 tensor-core convolutions, the precision of the reference's own cuDNN
    path on Ampere+): depth must still meet the 1e-3 relative tolerance on the consistent scene."""
    from itermvs_b200 import _lib
    s = make_sample(640, 512, n_src=4, batch=1, seed=0, scene="plane")
    want = O.pipeline_forward(dtu_weights, s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"], iteration=4)
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    _lib.set_conv_passes(1)
    try:
        with torch.no_grad():
            out = model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
        torch.cuda.synchronize()
    finally:
        _lib.set_conv_passes(4)
    d, dref = out["depths_upsampled"].cpu(), want["depths_upsampled"]
    rel = ((d - dref).abs() / dref).numpy()
    print(f"TF32 single pass: depth rel err mean {rel.mean():.2e} median {np.median(rel):.2e} max {rel.max():.2e}, "
          f"px>1e-3: {100 * (rel > 1e-3).mean():.3f}%")
    assert np.median(rel) < 2e-4
    assert (rel > 1e-3).mean() < 0.05


def test_3xtf32_mode_matches_default(dev, model, dtu_weights):
    """mode 3 (3-product TF32 split) and the default mode 4 (3-product FP16 split) are both fp32-grade:
    each within 1e-3 of the oracle everywhere, and within 1e-4 of each other."""
    from itermvs_b200 import _lib
    s = make_sample(640, 512, n_src=4, batch=1, seed=0, scene="plane")
    want = O.pipeline_forward(dtu_weights, s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"], iteration=4)
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    outs = {}
    assert _lib.get_conv_passes() == 4
    try:
        for mode in (4, 3):
            _lib.set_conv_passes(mode)
            with torch.no_grad():
                outs[mode] = model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))["depths_upsampled"].cpu()
    finally:
        _lib.set_conv_passes(4)
    dref = want["depths_upsampled"]
    for mode, d in outs.items():
        rel = ((d - dref).abs() / dref).numpy()
        print(f"mode {mode}: depth rel err mean {rel.mean():.2e} max {rel.max():.2e}")
        assert rel.max() < 1e-3 and rel.mean() < 2e-6
    assert float(((outs[3] - outs[4]).abs() / dref).max()) < 1e-4


def test_error_behaviour(dev, model):
    """Bad arguments surface as Python exceptions carrying the library's message (no crash, no sync)."""
    from itermvs_b200 import _lib
    import ctypes as C
    pb = _lib.Problem(1, 5, 500, 640, 32, 4)                 # H not a multiple of 32
    assert _lib.lib().imvs_forward_workspace_bytes(C.byref(pb)) == 0
    assert b"multiples of 32" in _lib.lib().imvs_last_error()
    with pytest.raises(RuntimeError):
        _lib.set_conv_passes(2)
    with pytest.raises(RuntimeError):                        # CPU tensors are rejected: there is no CPU path
        model({"level_0": torch.zeros(1, 3, 3, 64, 64)}, {}, torch.ones(1), torch.ones(1))


def _d48_model(dev, dtu_weights):
    import itermvs_b200
    torch.manual_seed(0)
    m = itermvs_b200.Pipeline(iteration=4, test=True)
    m.iter_mvs.update.hidden_init_head[0] = torch.nn.Conv2d(48, 64, 3, stride=1, padding=1, bias=False)
    sd = {k: v for k, v in dtu_weights.items() if k != "iter_mvs.update.hidden_init_head.0.weight"}
    m.load_state_dict(sd, strict=False)
    return m.to(dev).eval()


def test_config5_stress_shape(dev, dtu_weights):
    """BASELINE config 5 (Tanks&Temples shape): 1920x1056, 7 source views, D=48 (hidden_init_head.0
    re-created for 48 hypotheses as in SURVEY 8c), 4 iterations: size-independent properties at full
    size (runs, deterministic, finite, inside the depth range); oracle parity for D=48 / 7 views at a
    size the oracle finishes in seconds."""
    m = _d48_model(dev, dtu_weights)
    cu = lambda x: {k: v.to(dev) for k, v in x.items()}
    s = make_sample(1920, 1056, n_src=7, batch=1, seed=2, scene="plane")
    with torch.no_grad():
        out = m(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
        d1 = out["depths_upsampled"].clone()
        out2 = m(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    assert torch.equal(d1, out2["depths_upsampled"])
    assert d1.shape == (1, 1, 1056, 1920) and torch.isfinite(d1).all()
    assert float(d1.min()) >= 425 - 1e-2 and float(d1.max()) <= 935 + 1e-2
    c = out["confidence_upsampled"]
    assert torch.isfinite(c).all() and float(c.min()) >= 0 and float(c.max()) <= 1
    # D = 48, 7 source views against the oracle at 320x256
    s = make_sample(320, 256, n_src=7, batch=1, seed=2, scene="plane")
    with torch.no_grad():
        out = m(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    w = dict(dtu_weights)
    w["iter_mvs.update.hidden_init_head.0.weight"] = m.iter_mvs.update.hidden_init_head[0].weight.detach().cpu()
    want = O.pipeline_forward(w, s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"], iteration=4, num_sample=48)
    rel = ((out["depths_upsampled"].cpu() - want["depths_upsampled"]).abs() / want["depths_upsampled"]).numpy()
    print(f"D=48/7src @320x256: depth rel err median {np.median(rel):.2e} mean {rel.mean():.2e}, px>1e-3: {100 * (rel > 1e-3).mean():.3f}%")
    # hidden_init_head.0 is random for D != 32 (no checkpoint exists): the estimator runs outside its trained
    # regime, distributions are flat and isolated arg-max bins flip at fp32-reassociation level:
    # median-tight, a bounded fraction of pixels may move by a bin
    # measured on B200 (round 2): median 0, mean 6.0e-8, 0 % of pixels beyond 1e-3
    assert np.median(rel) < 1e-6, (np.median(rel), rel.mean())
    assert (rel > 1e-3).mean() < 5e-3, (rel > 1e-3).mean()


def test_tcgen05_path_matches_mma_sync(dev, stage_kats, model):
    """TF32 1-pass mode: the tcgen05/TMEM convolution (GRU, head conv0) against the mma.sync kernels on the
    same TF32-rounded operands -- they differ only by fp32 accumulation order."""
    from itermvs_b200 import _lib
    L = _lib.lib()
    upd = model.iter_mvs.update
    g = torch.Generator().manual_seed(5)
    h = torch.tanh(torch.randn(2, 32, 40, 72, generator=g)).to(dev)        # W = 72: two 60-wide tiles, ragged
    x = (torch.randn(2, 11, 40, 72, generator=g) * 0.5).to(dev)
    _lib.set_conv_passes(1)
    try:
        L.imvs_set_tcgen05(0)
        ref_h = upd.gru(h, x)
        upd.return_probability = True
        ref_nd, ref_p = upd.depth_init(h)
        L.imvs_set_tcgen05(1)
        got_h = upd.gru(h, x)
        got_nd, got_p = upd.depth_init(h)
        assert L.imvs_tcgen05_status() == 0, "a tcgen05 kernel timed out on its mbarrier"
    finally:
        upd.return_probability = None
        L.imvs_set_tcgen05(1)
        _lib.set_conv_passes(4)
    print("tcgen05 vs mma.sync: gru max diff", maxerr(got_h, ref_h), "prob max diff", maxerr(got_p, ref_p))
    # identical TF32 products and accumulation order; the tcgen05 epilogue uses ex2.approx-based gates
    # (|err| ~1e-6 relative on the exponential), far below the TF32 operand rounding (5e-4)
    assert maxerr(got_h, ref_h) < 3e-4
    assert maxerr(got_p, ref_p) < 3e-4
    # and against the fp32 golden GRU within TF32 accuracy
    k = stage_kats
    _lib.set_conv_passes(1)
    try:
        out = upd.gru(T(k["gru_h"]).to(dev), T(k["gru_x"]).to(dev))
    finally:
        _lib.set_conv_passes(4)
    assert maxerr(out, T(k["gru_out"])) < 5e-3


@pytest.mark.parametrize("cin,cout,dil", [(16, 16, 1), (32, 32, 1), (48, 48, 1), (48, 32, 1), (48, 16, 1), (32, 32, 2)])
@pytest.mark.parametrize("shape", [(2, 40, 72), (1, 9, 31), (5, 128, 160)])
def test_conv3x3_tma_tcgen05_vs_fp64(dev, cin, cout, dil, shape):
    """The persistent TMA + tcgen05 convolution (csrc/tc5pconv.cuh; split-plane operands, three fp16 products, fp32
    accumulation in TMEM) against an fp64 convolution of the same fp32 operands.  Measured on B200: max error 0.7 / 1.4 /
    2.0e-6 of the output scale for 27 / 54 / 81 accumulation steps (the tensor core's fp32 accumulate truncates), asserted 5e-6.  Shapes: ragged tiles (W = 72: two full 30-column tiles + 12; 9 x 31: one partial tile, single M-block
    path), and 5 x 128 x 160 (the two-M-block persistent path with several tiles per CTA).  Both store paths (fp32 NHWC,
    split planes) and the residual / ReLU / bias epilogue are exercised."""
    from itermvs_b200 import _lib, _pack
    L = _lib.lib()
    n, h, w = shape
    g = torch.Generator().manual_seed(cin * 1000 + cout + h)
    x = torch.randn(n, h, w, cin, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (3.0 * cin ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(n, h, w, cout, generator=g)
    packs = _pack.pack_mma_conv(wt.to(dev))
    wumma = packs[5]
    assert wumma is not None
    ws_bytes = L.imvs_conv3x3_tcgen05_workspace_bytes(n, h, w, cin, cout)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    xd, bd, rd = x.to(dev).contiguous(), bias.to(dev), res.to(dev).contiguous()
    want = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), bias.double(), padding=dil, dilation=dil)
    want_res = torch.relu(want + res.permute(0, 3, 1, 2).double())
    st = torch.cuda.current_stream().cuda_stream
    for use_res, relu, via_split in ((False, 0, 0), (True, 1, 0), (True, 1, 1), (False, 0, 1)):
        out = torch.full((n, h, w, cout), float("nan"), device=dev)
        _lib.check(L.imvs_conv3x3_tcgen05(xd.data_ptr(), wumma.data_ptr(), bd.data_ptr(), rd.data_ptr() if use_res else None, out.data_ptr(),
                                          ws.data_ptr(), ws_bytes, n, h, w, cin, cout, dil, relu, via_split, st), "conv3x3_tcgen05")
        torch.cuda.synchronize()
        assert _lib.device_status(clear=True) == 0
        ref = (want_res if use_res else want).permute(0, 2, 3, 1)
        err = float((out.cpu().double() - ref).abs().max())
        scale = float(ref.abs().max())
        print(f"tc5p {cin}->{cout} dil {dil} {shape} res={use_res} split_out={via_split}: max err {err:.2e} (scale {scale:.2f})")
        assert torch.isfinite(out).all()
        assert err < 5e-6 * max(scale, 1.0), err


@pytest.mark.parametrize("switch", ["TC5P_CORR", "CORR_FUSED", "CORR_TILE"])
def test_corrnet_on_tma_tcgen05_kernel(dev, stage_kats, model, monkeypatch, switch):
    """CorrNet with all six layers on the persistent TMA + tcgen05 kernel: 8-channel layers through an aliased K chunk, the two
    stride-2 layers on parity planes, the two transposed layers with four parity accumulators, three weight sets chosen per
    slice.  IMVS_TUNE_TC5P_CORR=1: one launch per layer; IMVS_TUNE_CORR_FUSED=1: the whole pass as ONE cooperative launch with
    grid-wide barriers between the layers; IMVS_TUNE_CORR_TILE=1: one launch that keeps a 32 x 32 tile in shared memory through
    all six layers (exact fp32 FFMA, corrnet_tile.cuh).  Against the reference's KAT and, in the three-set batched form the iterations use
    (graph-captured as the serving loop does), against the default mma.sync path."""
    from itermvs_b200 import _lib
    k = stage_kats
    ev = model.iter_mvs.evaluation
    x = T(k["corrnet_in"]).to(dev)
    monkeypatch.setenv("IMVS_TUNE_" + switch, "1")
    got = ev.corr_conv1[0](x)
    torch.cuda.synchronize()
    assert _lib.device_status(clear=True) == 0
    monkeypatch.setenv("IMVS_TUNE_" + switch, "0")
    ref = ev.corr_conv1[0](x)
    print(switch, "corrnet tcgen05 vs KAT", maxerr(got, T(k["corrnet0_out"])), "vs mma.sync", maxerr(got, ref))
    assert maxerr(got, T(k["corrnet0_out"])) < 2e-5
    assert maxerr(got, ref) < 2e-5
    # the batched three-set form (10 slices: 4 + 4 + 2) through the iteration branch of Evaluation
    s = make_sample(320, 256, n_src=2, batch=1, seed=7, scene="plane")
    cu = lambda d: {kk: v.to(dev) for kk, v in d.items()}
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("IMVS_TUNE_" + switch, flag)
        with torch.no_grad():
            out = model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
        torch.cuda.synchronize()
        outs.append(out["depths_upsampled"].clone())
    assert _lib.device_status(clear=True) == 0
    rel = ((outs[0] - outs[1]).abs() / outs[1]).max()
    print(switch, "pipeline with CorrNet on tcgen05 vs default: max rel depth difference", float(rel))
    assert float(rel) < 1e-4
    # captured in a CUDA graph and replayed (cooperative launch + memset node inside the capture)
    from itermvs_b200.graph import GraphedPipeline
    monkeypatch.setenv("IMVS_TUNE_" + switch, "1")
    g = GraphedPipeline(model, cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev), workspace_slot=5)
    for _ in range(3):
        o = g(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
    torch.cuda.synchronize()
    assert _lib.device_status(clear=True) == 0
    assert torch.equal(o["depths_upsampled"], outs[0])


@pytest.mark.parametrize("switch,value", [("CORR_TILE05", 1), ("CORR_TILE05", 2), ("CORR_TILE05", 3), ("CORR_TILEMID", 1),
                                          ("WCI_VPER", 1), ("WCI_VPER", 2), ("WC_WARPS", 26), ("WC_WARPS", 28)])
def test_launch_granularity_switches_are_bit_identical(dev, model, monkeypatch, switch, value):
    """The launch-granularity switches measured in round 2 (profiles/ps_experiments_r02.md section 6, profiles/README.md: tile
    shapes of CorrNet's mma.sync layers, source views of the init plane sweep split over blocks, 26 / 28 warps in the iteration
    kernel) only change which CTA / warp computes an output, never its summation order: the pipeline's outputs are bit-identical.
    Exception: CORR_TILEMID (one row-tile per warp) changes the number of accumulator sets of the fp16 3-product MMAs
    (mmaconv.cuh: NACC), i.e. the order of three fp32 additions per output: equal to the last few ulp."""
    from itermvs_b200 import _lib
    s = make_sample(320, 256, n_src=3, batch=2, seed=11, scene="plane")
    cu = lambda d: {kk: v.to(dev) for kk, v in d.items()}
    outs = []
    for flag in (str(value), None):
        if flag is None:
            monkeypatch.delenv("IMVS_TUNE_" + switch, raising=False)
        else:
            monkeypatch.setenv("IMVS_TUNE_" + switch, flag)
        with torch.no_grad():
            out = model(cu(s["imgs"]), cu(s["proj_matrices"]), s["depth_min"].to(dev), s["depth_max"].to(dev))
        torch.cuda.synchronize()
        outs.append({k: out[k].clone() for k in ("depths_upsampled", "confidence_upsampled")})
    assert _lib.device_status(clear=True) == 0
    for k in outs[0]:
        if switch == "CORR_TILEMID":
            rel = (outs[0][k] - outs[1][k]).abs() / outs[1][k].abs().clamp_min(1e-3)
            print(switch, k, "max rel", float(rel.max()), "px > 1e-5:", float((rel > 1e-5).float().mean()))
            assert float((rel > 1e-5).float().mean()) <= 1e-4 and float(rel.max()) < 1e-3, (switch, value, k)   # (an arg-max tie may flip)
        else:
            assert torch.equal(outs[0][k], outs[1][k]), (switch, value, k)


@pytest.mark.parametrize("width,height,n_src,d", [(160, 128, 2, 8), (320, 256, 3, 48), (96, 64, 9, 32)])
def test_padded_level3_entry_points(dev, width, height, n_src, d):
    """imvs_pad_level3 + imvs_warpcorr_init_padded / imvs_warpcorr_iter_padded (the estimator's default since round 2: level 3 read
    from a copy with 256 bytes per texel, include/itermvs_b200.h) against the 48-channel entry points on the same inputs: the padded
    copy holds exactly the documented floats, the outputs agree to fp32 summation order."""
    from itermvs_b200 import _lib, ops
    L = _lib.lib()
    g = torch.Generator().manual_seed(5)
    V = n_src + 1
    fea = [torch.randn(2, V, height // s_, width // s_, c, generator=g).to(dev) for s_, c in ((2, 16), (4, 32), (8, 48))]
    smp = make_sample(width, height, n_src=n_src, batch=2, seed=3, scene="noise")
    rts = [ops.compose_projections(smp["proj_matrices"][f"level_{l}"].float().to(dev)) for l in (1, 2, 3)]
    H2, W2, H3, W3 = height // 4, width // 4, height // 8, width // 8
    nd = torch.rand(2, H2 * W2, generator=g).to(dev)
    vw = torch.rand(2, n_src, H2 * W2, generator=g).to(dev)
    dmin, dmax = smp["depth_min"].float().to(dev), smp["depth_max"].float().to(dev)
    st = torch.cuda.current_stream().cuda_stream
    f3p = torch.full((2, V, H3, W3, 64), float("nan"), device=dev)
    _lib.check(L.imvs_pad_level3(fea[2].data_ptr(), f3p.data_ptr(), 2, V, H3, W3, st), "pad_level3")
    for grp in range(8):
        assert torch.equal(f3p[..., 4 * grp:4 * grp + 4], fea[2][..., 6 * grp:6 * grp + 4])
        assert torch.equal(f3p[..., 32 + 4 * grp:34 + 4 * grp], fea[2][..., 6 * grp + 4:6 * grp + 6])
        assert float(f3p[..., 34 + 4 * grp:36 + 4 * grp].abs().max()) == 0.0
    outs = []
    for padded in (False, True):
        f3 = f3p if padded else fea[2]
        init_fn = L.imvs_warpcorr_init_padded if padded else L.imvs_warpcorr_init
        iter_fn = L.imvs_warpcorr_iter_padded if padded else L.imvs_warpcorr_iter
        corr = torch.full((2, n_src, d, H3 * W3, 8), float("nan"), device=dev)
        agg = torch.full((2, 10, H2 * W2, 8), float("nan"), device=dev)
        _lib.check(init_fn(f3.data_ptr(), rts[2].data_ptr(), dmin.data_ptr(), dmax.data_ptr(), None, corr.data_ptr(), 2, V, H3, W3, d, st),
                   "warpcorr_init")
        _lib.check(iter_fn(fea[0].data_ptr(), fea[1].data_ptr(), f3.data_ptr(), rts[0].data_ptr(), rts[1].data_ptr(), rts[2].data_ptr(),
                           nd.data_ptr(), H2 * W2, 1, vw.data_ptr(), dmin.data_ptr(), dmax.data_ptr(), None, None, None, agg.data_ptr(),
                           2, V, H2, W2, st), "warpcorr_iter")
        torch.cuda.synchronize()
        outs.append((corr, agg))
    for a, b, name in ((outs[0][0], outs[1][0], "corr"), (outs[0][1], outs[1][1], "agg")):
        assert torch.isfinite(b).all()
        print(name, "padded vs 48-channel: max abs diff", maxerr(a, b), "of", float(a.abs().max()))
        assert maxerr(a, b) < 2e-6 * max(1.0, float(a.abs().max()))
    assert torch.equal(outs[0][1][:, :8], outs[1][1][:, :8])          # levels 1 and 2 run the same code: bit-identical


def test_image_pyramid_and_prefetch_loader_on_device(dev, model, tmp_path):
    """f-4 on the GPU: (1) `imvs_image_pyramid_u8` -- raw 8-bit image -> 2 x / 255 - 1 -> cv2.resize(INTER_LINEAR) -> levels 1..3 --
    against OpenCV (1 ulp of the value: its SIMD / IPP paths may fuse a product); (2) `io.PrefetchLoader`: PNG files decoded on
    a background thread into pinned uint8 buffers, uploaded and prepared on a copy stream, yields the loader's sample dict on the
    device; its level_0 equals the host loader's (`io.load_views` = the reference's `__getitem__`) to 1 ulp and the pipeline's
    depth map from it equals the one from the host-loaded sample."""
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    from itermvs_b200 import io as mio
    rng = np.random.default_rng(11)
    for (h0, w0), (w, h) in (((1200, 1600), (1152, 864)), ((300, 400), (320, 256)), ((256, 320), (320, 256))):
        img = rng.integers(0, 256, size=(h0, w0, 3), dtype=np.uint8)
        want0 = cv2.resize(2 * img.astype(np.float32) / 255. - 1, (w, h), interpolation=cv2.INTER_LINEAR)
        got = mio.image_pyramid_device(torch.from_numpy(img).to(dev), (w, h))
        torch.cuda.synchronize()
        e0 = float(np.abs(got["level_0"].cpu().numpy() - want0.transpose(2, 0, 1)).max())
        assert e0 <= 2.5e-7, e0
        for k in (1, 2, 3):
            wantk = cv2.resize(want0, (w >> k, h >> k), interpolation=cv2.INTER_LINEAR)
            assert float(np.abs(got[f"level_{k}"].cpu().numpy() - wantk.transpose(2, 0, 1)).max()) <= 5e-7
    with pytest.raises(RuntimeError):
        mio.image_pyramid_device(torch.zeros(4, 4, 3, dtype=torch.uint8), (4, 4))
    # a scan on disk in the loader's directory layout
    s = make_sample(320, 256, n_src=2, batch=1, seed=13, scene="plane")
    scan = tmp_path / "scan1"
    (scan / "images").mkdir(parents=True)
    (scan / "cams_1").mkdir()
    u8 = ((s["imgs"]["level_0"][0] + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)          # [V, 3, H, W]
    proj0 = s["proj_matrices"]["level_0"][0].double().numpy()
    for v in range(3):
        Image.fromarray(u8[v].permute(1, 2, 0).numpy()).save(str(scan / "images" / f"{v:08d}.png"))
        # cam file: extrinsics = identity-free form is not recoverable from a projection; store K = P[:3,:3], E = [I | K^-1 t]
        K = proj0[v][:3, :3]
        E = np.eye(4)
        E[:3, 3] = np.linalg.solve(K, proj0[v][:3, 3])
        with open(scan / "cams_1" / f"{v:08d}_cam.txt", "w") as f:
            f.write("extrinsic\n" + "\n".join(" ".join(f"{x:.9g}" for x in row) for row in E) + "\n\nintrinsic\n" +
                    "\n".join(" ".join(f"{x:.9g}" for x in row) for row in K) + f"\n\n{float(s['depth_min'][0])} 1.0 192 {float(s['depth_max'][0])}\n")
    items = [([str(scan / "images" / f"{v:08d}.png") for v in order], [str(scan / "cams_1" / f"{v:08d}_cam.txt") for v in order])
             for order in ((0, 1, 2), (1, 0, 2), (2, 0, 1))]
    outs = []
    with torch.no_grad():
        for sample in mio.PrefetchLoader(items, img_wh=(320, 256), orig_wh=(320, 256), device=dev):
            assert sample["imgs"]["level_0"].shape == (1, 3, 3, 256, 320) and sample["imgs"]["level_3"].shape == (1, 3, 3, 32, 40)
            out = model(sample["imgs"], sample["proj_matrices"], sample["depth_min"], sample["depth_max"])
            outs.append((sample, out["depths_upsampled"].clone()))
    assert len(outs) == 3
    # the host loader on the same files (jpg in the reference; the decode is not what is under test)
    def host_sample(order):
        imgs, proj = [], []
        for v in order:
            k_, e_, dmin, dmax = mio.read_cam_file(str(scan / "cams_1" / f"{v:08d}_cam.txt"))
            imgs.append(2 * np.array(Image.open(str(scan / "images" / f"{v:08d}.png")), dtype=np.float32) / 255. - 1)
            proj.append(mio.projection_pyramid(k_, e_, (320, 256), (320, 256)))
        return np.stack(imgs).transpose(0, 3, 1, 2), {lv: np.stack([p[lv] for p in proj]) for lv in proj[0]}
    himg, hproj = host_sample((0, 1, 2))
    assert float((outs[0][0]["imgs"]["level_0"][0].cpu() - torch.from_numpy(himg)).abs().max()) <= 2.5e-7
    for lv in hproj:
        assert torch.equal(outs[0][0]["proj_matrices"][lv][0].cpu(), torch.from_numpy(hproj[lv]))
    with torch.no_grad():
        ref = model({"level_0": torch.from_numpy(himg)[None].to(dev)}, {lv: torch.from_numpy(p)[None].to(dev) for lv, p in hproj.items()},
                    outs[0][0]["depth_min"], outs[0][0]["depth_max"])
    rel = ((outs[0][1] - ref["depths_upsampled"]).abs() / ref["depths_upsampled"]).max()
    print("prefetch loader vs host loader: max rel depth difference", float(rel))
    assert float(rel) < 1e-4
