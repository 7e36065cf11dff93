"""CPU-side execution of the REAL kernel sources of the plane-sweep path (warp.cu, warpcorr.cu) through
tests/cusim (a fiber-per-thread emulation of the CUDA execution model -- test infrastructure, never loaded by the
product), compared with the oracle on small inputs.  What this covers without a GPU: index arithmetic, the
shared-memory record protocol, the warp-shuffle regrouping of the 48-channel level, the persistent block's item
queue with ragged tiles / several blocks, the hypothesis generation from nd, the aggregation, the backward
scatter.  What it cannot cover: anything about speed, inter-warp memory-model races, the tensor-core files.

The GPU parity tests (tests/test_gpu_parity.py) remain the parity tests proper; these run under -m "not gpu".
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from oracle import itermvs_oracle as O
from itermvs_b200.synthetic import make_sample, random_feature_pyramids

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cusim"))
import cusim_build  # noqa: E402

vp, ci, sz = C.c_void_p, C.c_int, C.c_size_t


@pytest.fixture(scope="module")
def sim():
    lib = C.CDLL(cusim_build.build())
    lib.imvs_last_error.restype = C.c_char_p
    sigs = {
        "imvs_compose_projections": [vp, ci, ci, vp, vp, vp],
        "imvs_differentiable_warping": [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp],
        "imvs_differentiable_warping_backward": [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp],
        "imvs_nchw_to_nhwc": [vp, vp, ci, ci, ci, ci, vp],
        "imvs_warpcorr_init": [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp],
        "imvs_aggregate_init": [vp, vp, vp, ci, ci, ci, ci, vp],
        "imvs_warpcorr_iter": [vp, vp, vp, vp, vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp],
        "imvs_pad_level3": [vp, vp, ci, ci, ci, ci, vp],
        "imvs_warpcorr_init_padded": [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp],
        "imvs_warpcorr_iter_padded": [vp, vp, vp, vp, vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = ci, args
    return lib


def ok(lib, rc):
    assert rc == 0, lib.imvs_last_error().decode()


def P(t):
    return None if t is None else t.data_ptr()


def f32(t):
    return t.detach().float().contiguous()


def maxerr(a, b):
    return float((a.double() - b.double()).abs().max())


def stack_views(ref, srcs):
    """NCHW per-view maps -> [B][V][H][W][C] (view 0 = reference), the layout of the fused kernels."""
    return f32(torch.stack([ref] + list(srcs), dim=1).permute(0, 1, 3, 4, 2))


def pad_level3(lib, fea3):
    """imvs_pad_level3: [B][V][H3][W3][48] -> 64 floats per texel; also checks the documented positions."""
    b, v, h3, w3, _ = fea3.shape
    out = torch.full((b, v, h3, w3, 64), float("nan"))
    ok(lib, lib.imvs_pad_level3(P(fea3), P(out), b, v, h3, w3, None))
    for g in range(8):
        assert torch.equal(out[..., 4 * g:4 * g + 4], fea3[..., 6 * g:6 * g + 4])
        assert torch.equal(out[..., 32 + 4 * g:32 + 4 * g + 2], fea3[..., 6 * g + 4:6 * g + 6])
        assert float(out[..., 32 + 4 * g + 2:32 + 4 * g + 4].abs().max()) == 0.0
    return out


def compose(lib, ref_proj, src_projs):
    proj = f32(torch.stack([ref_proj] + list(src_projs), dim=1))
    b, v = proj.shape[:2]
    out = torch.empty(b, v - 1, 12)
    ok(lib, lib.imvs_compose_projections(P(proj), b, v, P(out), None, None))
    return out


def feature_inputs(width, height, n_src, batch, seed):
    ref, srcs = random_feature_pyramids(width, height, n_src, batch, seed)
    s = make_sample(width, height, n_src=n_src, batch=batch, seed=seed, scene="noise")
    rp, sp = {}, {}
    for l in (1, 2, 3):
        pm = torch.unbind(s["proj_matrices"][f"level_{l}"].float(), 1)
        rp[f"level{l}"], sp[f"level{l}"] = pm[0], list(pm[1:])
    return ref, srcs, rp, sp, s


def test_sim_is_not_the_product_library():
    from itermvs_b200 import _build
    assert "tests/cusim/_build" in cusim_build.LIB and cusim_build.LIB != _build.LIB
    import itermvs_b200
    src = "".join(open(os.path.join(os.path.dirname(itermvs_b200.__file__), f)).read()
                  for f in os.listdir(os.path.dirname(itermvs_b200.__file__)) if f.endswith(".py"))
    assert "cusim" not in src and "libitermvs_sim" not in src


@pytest.mark.parametrize("tag", ["same", "fea2x", "fea_half", "b2"])
def test_differentiable_warping_source_on_cpu(sim, stage_kats, tag):
    """warp_nchw_kernel against the reference-generated fixture, its backward against oracle autograd."""
    k = stage_kats
    fea, sp, rp, dep = (f32(torch.from_numpy(k[f"warp_{tag}_{n}"])) for n in ("fea", "src_proj", "ref_proj", "depth"))
    b, c, h1, w1 = fea.shape
    _, d, h, w = dep.shape
    out = torch.empty(b, c, d, h, w)
    rt = torch.empty(b, 12)
    ok(sim, sim.imvs_differentiable_warping(P(fea), P(sp), P(rp), P(dep), P(out), b, c, h1, w1, d, h, w, P(rt), None, None))
    assert maxerr(out, torch.from_numpy(k[f"warp_{tag}_out"])) < 2e-4
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    f_cpu = fea.clone().requires_grad_(True)
    (O.differentiable_warping(f_cpu, sp, rp, dep) * gout).sum().backward()
    gfea = torch.full(fea.shape, float("nan"))
    ok(sim, sim.imvs_differentiable_warping_backward(P(gout), P(sp), P(rp), P(dep), P(gfea), b, c, h1, w1, d, h, w, P(rt), None, None))
    assert maxerr(gfea, f_cpu.grad) < 2e-4 * max(1.0, float(f_cpu.grad.abs().max()))


@pytest.mark.parametrize("padded", [False, True])
@pytest.mark.parametrize("batch,n_src,d", [(1, 2, 32), (2, 3, 8), (1, 1, 48)])
def test_warpcorr_init_source_on_cpu(sim, batch, n_src, d, padded):
    """warpcorr_init_kernel + aggregate_init_kernel against the oracle's warp -> group correlation -> aggregation,
    with explicit samples (incl. z <= 0.01 substitutions) and with in-kernel hypothesis generation."""
    ref, srcs, rp, sp, s = feature_inputs(96, 64, n_src, batch, seed=11)       # level 3: 12 x 8
    h3, w3 = ref["level3"].shape[2:]
    inv_min = (1.0 / s["depth_min"]).view(batch, 1, 1, 1)
    inv_max = (1.0 / s["depth_max"]).view(batch, 1, 1, 1)
    ds = O.initial_depth_samples(inv_min, inv_max, d, h3, w3)
    fea3 = stack_views(ref["level3"], srcs["level3"])
    init_fn = sim.imvs_warpcorr_init
    if padded:                    # the same kernel on the 256-byte-per-texel copy of the pyramid
        fea3, init_fn = pad_level3(sim, fea3), sim.imvs_warpcorr_init_padded
    rt3 = compose(sim, rp["level3"], sp["level3"])
    dmin, dmax = f32(s["depth_min"]), f32(s["depth_max"])

    def want_for(samples):
        return torch.stack([O.group_correlation(O.differentiable_warping(src, p, rp["level3"], samples), ref["level3"])
                            for src, p in zip(srcs["level3"], sp["level3"])], dim=1)      # [B,S,G,D,H,W]

    corr = torch.full((batch, n_src, d, h3 * w3, 8), float("nan"))
    ok(sim, init_fn(P(fea3), P(rt3), P(dmin), P(dmax), None, P(corr), batch, n_src + 1, h3, w3, d, None))
    got = corr.view(batch, n_src, d, h3, w3, 8).permute(0, 1, 5, 2, 3, 4)
    want = want_for(ds)
    assert maxerr(got, want) < 1e-4
    ds2 = ds.clone()
    ds2[:, 3, :2] = -10.0
    ok(sim, init_fn(P(fea3), P(rt3), None, None, P(f32(ds2)), P(corr), batch, n_src + 1, h3, w3, d, None))
    want2 = want_for(ds2)
    assert maxerr(corr.view(batch, n_src, d, h3, w3, 8).permute(0, 1, 5, 2, 3, 4), want2) < 1e-4
    # aggregation (itermvs.py:59-69)
    vw3 = torch.rand(batch, n_src, h3, w3, generator=torch.Generator().manual_seed(2))
    agg = torch.full((batch, d, h3 * w3, 8), float("nan"))
    ok(sim, sim.imvs_aggregate_init(P(corr), P(vw3), P(agg), batch, n_src, d, h3 * w3, None))
    w = vw3.view(batch, n_src, 1, 1, h3, w3)
    want_agg = (want2 * w).sum(1) / (1e-5 + w.sum(1))
    assert maxerr(agg.view(batch, d, h3, w3, 8).permute(0, 4, 1, 2, 3), want_agg) < 1e-4


# 1..8 source views run the unrolled instantiations, 9 the rolled one; CUSIM_SMS picks how many persistent blocks
# share the tile range (1 block; 3 blocks with a ragged split; more blocks than tiles)
@pytest.mark.parametrize("batch,n_src,width,height,sms", [(1, 4, 96, 64, 3), (2, 3, 64, 64, 1), (1, 1, 64, 32, 64),
                                                          (1, 7, 64, 32, 2), (1, 9, 64, 32, 3), (1, 2, 96, 96, 5),
                                                          (1, 16, 64, 32, 3)])       # 16 = IMVS_MAX_VIEWS
@pytest.mark.parametrize("padded", [False, True])
def test_warpcorr_iter_source_on_cpu(sim, batch, n_src, width, height, sms, monkeypatch, padded):
    monkeypatch.setenv("CUSIM_SMS", str(sms))
    ref, srcs, rp, sp, s = feature_inputs(width, height, n_src, batch, seed=12)
    h2, w2 = ref["level2"].shape[2:]
    g = torch.Generator().manual_seed(3)
    inv_min = (1.0 / s["depth_min"]).view(batch, 1, 1, 1)
    inv_max = (1.0 / s["depth_max"]).view(batch, 1, 1, 1)
    nd = torch.rand(batch, 1, h2, w2, generator=g)
    nd[:, :, 0, :4] = torch.tensor([0.0, 1.0, 0.001, 0.999])          # clamp at both ends of the range
    samples = {f"level{l}": O.iteration_depth_samples(nd, l, inv_min, inv_max) for l in (1, 2, 3)}
    vw = torch.rand(batch, n_src, h2, w2, generator=g)
    aggs = _aggregated_only(ref, srcs, rp, sp, samples, vw)
    want = torch.cat(aggs, dim=2)                                       # [B,8,10,H2,W2]
    feas = [stack_views(ref[f"level{l}"], srcs[f"level{l}"]) for l in (1, 2, 3)]
    iter_fn = sim.imvs_warpcorr_iter
    if padded:
        feas[2], iter_fn = pad_level3(sim, feas[2]), sim.imvs_warpcorr_iter_padded
    rts = [compose(sim, rp[f"level{l}"], sp[f"level{l}"]) for l in (1, 2, 3)]
    dmin, dmax = f32(s["depth_min"]), f32(s["depth_max"])
    ndc = f32(nd)

    def run(explicit):
        agg = torch.full((batch, 10, h2 * w2, 8), float("nan"))
        smp = [f32(samples[f"level{l}"]) for l in (1, 2, 3)] if explicit else [None] * 3
        ok(sim, iter_fn(P(feas[0]), P(feas[1]), P(feas[2]), P(rts[0]), P(rts[1]), P(rts[2]),
                        None if explicit else P(ndc), h2 * w2, 1, P(f32(vw)),
                        None if explicit else P(dmin), None if explicit else P(dmax),
                        P(smp[0]), P(smp[1]), P(smp[2]), P(agg), batch, n_src + 1, h2, w2, None))
        return agg.view(batch, 10, h2, w2, 8).permute(0, 4, 1, 2, 3)

    first = run(True)
    assert maxerr(first, want) < 1e-4
    # hypotheses generated in the kernel from nd (itermvs.py:289-293): same result up to the rounding of 1/depth
    assert maxerr(run(False), want) < 2e-4
    # a different interleaving of the block's warps (who draws which item from the shared-memory queue, when) must
    # not change a single bit
    monkeypatch.setenv("CUSIM_SHUFFLE", str(7 + sms))
    assert torch.equal(run(True), first)
    monkeypatch.delenv("CUSIM_SHUFFLE")


def _aggregated_only(ref, srcs, rp, sp, samples, vw):
    """The oracle's evaluation_iter up to (not including) CorrNet: itermvs.py:86-120."""
    aggs = []
    for l in (1, 2, 3):
        key = f"level{l}"
        ref_l = O.resample_ref_feature(ref[key], l)
        ds = samples[key]
        b, r, h, w = ds.shape
        corr_sum, vw_sum = 0, 1e-5
        for i, (src, p) in enumerate(zip(srcs[key], sp[key])):
            corr = O.group_correlation(O.differentiable_warping(src, p, rp[key], ds), ref_l)
            v = vw[:, i].reshape(b, 1, 1, h, w)
            corr_sum = corr_sum + corr * v
            vw_sum = vw_sum + v
        aggs.append(corr_sum / vw_sum)
    return aggs


def test_fusion_kernels_source_on_cpu(sim, fusion_kat):
    """geo_consistency_kernel / fuse_finalize_kernel (fusion.cu) against the outputs of the reference's own
    reproject_with_depth / check_geometric_consistency / filter_depth loop (tests/golden/make_golden_fusion.py)."""
    from itermvs_b200.fusion import pair_cameras          # host-side camera algebra (numpy only)
    cf = C.c_float
    sim.imvs_check_geometric_consistency.restype = ci
    sim.imvs_check_geometric_consistency.argtypes = [vp, vp, vp, cf, cf, vp, vp, vp, vp, vp, vp, ci, ci, vp]
    sim.imvs_filter_depth_view.restype = ci
    sim.imvs_filter_depth_view.argtypes = [vp, vp, vp, vp, ci, cf, cf, cf, ci, vp, vp, vp, vp, vp, vp, ci, ci, vp]
    z = fusion_kat
    n = int(z["n_src"])
    depths = [np.ascontiguousarray(z[f"depth{v}"], dtype=np.float32) for v in range(1, n + 1)]
    d0 = np.ascontiguousarray(z["depth0"], dtype=np.float32)
    h, w = d0.shape
    flips = 0
    for v in range(1, n + 1):
        cams = np.ascontiguousarray(pair_cameras(z["K0"], z["E0"], z[f"K{v}"], z[f"E{v}"]))
        mask = np.empty((h, w), np.uint8)
        rep, xs, ys = (np.empty((h, w), np.float32) for _ in range(3))
        ok(sim, sim.imvs_check_geometric_consistency(d0.ctypes.data, depths[v - 1].ctypes.data, cams.ctypes.data, 1.0, 0.01,
                                                     mask.ctypes.data, rep.ctypes.data, xs.ctypes.data, ys.ctypes.data,
                                                     None, None, h, w, None))
        good = np.isfinite(z[f"x_src{v}"])
        assert np.allclose(xs[good], z[f"x_src{v}"][good], rtol=1e-6, atol=1e-4)
        assert np.allclose(ys[good], z[f"y_src{v}"][good], rtol=1e-6, atol=1e-4)
        diff = mask.astype(bool) != z[f"mask{v}"]
        flips += int(diff.sum())
        assert np.allclose(rep[~diff], z[f"reprojected{v}"][~diff], rtol=2e-6, atol=1e-4)
    assert flips <= 3, flips
    cams = np.ascontiguousarray(np.stack([pair_cameras(z["K0"], z["E0"], z[f"K{v}"], z[f"E{v}"]) for v in range(1, n + 1)]))
    srcs = np.ascontiguousarray(np.stack(depths))
    conf = np.ascontiguousarray(z["confidence"], dtype=np.float32)
    acc = np.empty((h, w), np.float32)
    cnt = np.empty((h, w), np.int32)
    avg = np.empty((h, w), np.float64)
    pm, gm, fm = (np.empty((h, w), np.uint8) for _ in range(3))
    ok(sim, sim.imvs_filter_depth_view(d0.ctypes.data, conf.ctypes.data, srcs.ctypes.data, cams.ctypes.data, n, 1.0, 0.01, 0.3, 3,
                                       acc.ctypes.data, cnt.ctypes.data, avg.ctypes.data, pm.ctypes.data, gm.ctypes.data,
                                       fm.ctypes.data, h, w, None))
    assert np.array_equal(pm.astype(bool), z["photo_mask"])
    assert int((gm.astype(bool) != z["geo_mask"]).sum()) <= 2 and int((fm.astype(bool) != z["final_mask"]).sum()) <= 2
    same = gm.astype(bool) == z["geo_mask"]
    with np.errstate(invalid="ignore"):
        close = np.isclose(avg, z["depth_est_averaged"], rtol=2e-6, atol=1e-4, equal_nan=True)
    assert (close | ~same).mean() > 0.9995


def _unstack_grad(g, n_src):
    """[B][V][H][W][C] gradient -> (reference NCHW, [source NCHW])."""
    g = g.permute(0, 1, 4, 2, 3)
    return g[:, 0], [g[:, v + 1] for v in range(n_src)]


def _sig():
    return [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp], \
           [vp, vp, vp, vp, vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]


@pytest.mark.parametrize("batch,n_src,d,explicit", [(1, 2, 8, True), (2, 3, 5, False)])
def test_warpcorr_init_backward_source_on_cpu(sim, batch, n_src, d, explicit):
    """warpcorr_init_bwd_kernel against torch autograd through the oracle's warp -> group correlation chain."""
    sim.imvs_warpcorr_init_backward.restype, sim.imvs_warpcorr_init_backward.argtypes = ci, _sig()[0]
    ref, srcs, rp, sp, s = feature_inputs(96, 64, n_src, batch, seed=21)
    h3, w3 = ref["level3"].shape[2:]
    inv_min = (1.0 / s["depth_min"]).view(batch, 1, 1, 1)
    inv_max = (1.0 / s["depth_max"]).view(batch, 1, 1, 1)
    ds = O.initial_depth_samples(inv_min, inv_max, d, h3, w3)
    if explicit:
        ds[:, 1, :2] = -10.0                      # z <= 0.01 substitution: those samples land outside the map
    r3 = ref["level3"].clone().requires_grad_(True)
    s3 = [t.clone().requires_grad_(True) for t in srcs["level3"]]
    corr = torch.stack([O.group_correlation(O.differentiable_warping(src, p, rp["level3"], ds), r3)
                        for src, p in zip(s3, sp["level3"])], dim=1)              # [B,S,G,D,H,W]
    gcorr = torch.randn(corr.shape, generator=torch.Generator().manual_seed(4))
    (corr * gcorr).sum().backward()
    fea3 = stack_views(ref["level3"], srcs["level3"])
    rt3 = compose(sim, rp["level3"], sp["level3"])
    g_k = f32(gcorr.permute(0, 1, 3, 4, 5, 2).reshape(batch, n_src, d, h3 * w3, 8))
    gfea = torch.full(fea3.shape, float("nan"))
    dmin, dmax = f32(s["depth_min"]), f32(s["depth_max"])
    ok(sim, sim.imvs_warpcorr_init_backward(P(fea3), P(rt3), None if explicit else P(dmin), None if explicit else P(dmax),
                                            P(f32(ds)) if explicit else None, P(g_k), P(gfea), batch, n_src + 1, h3, w3, d, None))
    gref, gsrcs = _unstack_grad(gfea, n_src)
    assert maxerr(gref, r3.grad) < 2e-4 * max(1.0, float(r3.grad.abs().max()))
    for got, t in zip(gsrcs, s3):
        assert maxerr(got, t.grad) < 2e-4 * max(1.0, float(t.grad.abs().max()))


@pytest.mark.parametrize("batch,n_src,explicit", [(1, 3, True), (2, 2, False)])
def test_warpcorr_iter_backward_source_on_cpu(sim, batch, n_src, explicit):
    """warpcorr_iter_bwd_kernel<level> against torch autograd through the oracle's iteration branch up to (not
    including) CorrNet: source features, reference features through the 2x2 mean / bilinear x2 resampling."""
    sim.imvs_warpcorr_iter_backward.restype, sim.imvs_warpcorr_iter_backward.argtypes = ci, _sig()[1]
    ref, srcs, rp, sp, s = feature_inputs(64, 64, n_src, batch, seed=22)
    h2, w2 = ref["level2"].shape[2:]
    g = torch.Generator().manual_seed(6)
    inv_min = (1.0 / s["depth_min"]).view(batch, 1, 1, 1)
    inv_max = (1.0 / s["depth_max"]).view(batch, 1, 1, 1)
    nd = torch.rand(batch, 1, h2, w2, generator=g)
    nd[:, :, 0, :4] = torch.tensor([0.0, 1.0, 0.001, 0.999])
    samples = {f"level{l}": O.iteration_depth_samples(nd, l, inv_min, inv_max) for l in (1, 2, 3)}
    vw = torch.rand(batch, n_src, h2, w2, generator=g)
    rg = {k: v.clone().requires_grad_(True) for k, v in ref.items()}
    sg = {k: [t.clone().requires_grad_(True) for t in v] for k, v in srcs.items()}
    want = torch.cat(_aggregated_only(rg, sg, rp, sp, samples, vw), dim=2)            # [B,8,10,H2,W2]
    gagg = torch.randn(want.shape, generator=g)
    (want * gagg).sum().backward()
    feas = [stack_views(ref[f"level{l}"], srcs[f"level{l}"]) for l in (1, 2, 3)]
    rts = [compose(sim, rp[f"level{l}"], sp[f"level{l}"]) for l in (1, 2, 3)]
    smp = [f32(samples[f"level{l}"]) for l in (1, 2, 3)] if explicit else [None] * 3
    gk = f32(gagg.permute(0, 2, 3, 4, 1).reshape(batch, 10, h2 * w2, 8))
    gf = [torch.full(f.shape, float("nan")) for f in feas]
    dmin, dmax, ndc = f32(s["depth_min"]), f32(s["depth_max"]), f32(nd)
    ok(sim, sim.imvs_warpcorr_iter_backward(P(feas[0]), P(feas[1]), P(feas[2]), P(rts[0]), P(rts[1]), P(rts[2]),
                                            None if explicit else P(ndc), h2 * w2, 1, P(f32(vw)),
                                            None if explicit else P(dmin), None if explicit else P(dmax),
                                            P(smp[0]), P(smp[1]), P(smp[2]), P(gk), P(gf[0]), P(gf[1]), P(gf[2]),
                                            batch, n_src + 1, h2, w2, None))
    for l in (1, 2, 3):
        gref, gsrcs = _unstack_grad(gf[l - 1], n_src)
        want_ref = rg[f"level{l}"].grad
        assert maxerr(gref, want_ref) < 2e-4 * max(1.0, float(want_ref.abs().max())), l
        for got, t in zip(gsrcs, sg[f"level{l}"]):
            assert maxerr(got, t.grad) < 2e-4 * max(1.0, float(t.grad.abs().max())), l


def test_kernels_stay_inside_their_buffers():
    """Memcheck without a GPU: the plane-sweep / warping kernels, forward and backward, on buffers that end (and, in a
    second pass, start) flush against an inaccessible page (tests/cusim/guarded.py).  An out-of-bounds read or write by
    one element would end the child with SIGSEGV."""
    import subprocess
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cusim", "memcheck_run.py")
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MEMCHECK-OK" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    # every inference kernel (FeatureNet + estimator, tensor-core engine included) at 32x32 with images, pyramids,
    # projections, outputs and both workspaces on guarded buffers
    r = subprocess.run([sys.executable, script, "--forward"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MEMCHECK-FORWARD-OK" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-2000:])


def test_load_tracer_counts(sim, tmp_path, monkeypatch):
    """The access-pattern tracer behind tools/wavefront_model.py: warp-level grouping of the __ldg calls.  For
    aggregate_init (one float + one float4 per thread and view, fully coalesced) the counts are known in closed form."""
    import json
    trace = tmp_path / "trace.jsonl"
    b, s_, d, p3 = 1, 3, 4, 256
    def aligned128(t):                                            # line counts depend on the 128-byte phase of the base
        buf = torch.empty(t.numel() + 32)
        off = ((-buf.data_ptr()) % 128) // 4
        out = buf[off:off + t.numel()].view(t.shape)
        out.copy_(t)
        return out
    corr, vw3, agg = aligned128(torch.rand(b, s_, d, p3, 8)), aligned128(torch.rand(b, s_, p3)), torch.empty(b, d, p3, 8)
    monkeypatch.setenv("CUSIM_TRACE", str(trace))
    ok(sim, sim.imvs_aggregate_init(P(corr), P(vw3), P(agg), b, s_, d, p3, None))
    monkeypatch.setenv("CUSIM_TRACE", "")
    rec = json.loads(trace.read_text().splitlines()[-1])
    tot = {k: sum(site[k] for site in rec["sites"]) for k in ("requests", "lanes", "bytes", "sectors", "lines", "wavefronts")}
    threads = b * d * p3 * 2
    assert tot["lanes"] == threads * s_ * 2                         # per thread and view: one weight, one float4 of corr
    assert tot["requests"] == threads // 32 * s_ * 2
    assert tot["bytes"] == threads * s_ * (4 + 16)
    # a warp's 32 float4 are 512 contiguous bytes = 16 sectors / 4 lines; its 32 weights (16 pixels x 2 halves) 64 bytes
    assert tot["sectors"] == threads // 32 * s_ * (16 + 2) and tot["lines"] == threads // 32 * s_ * (4 + 1)
    assert tot["wavefronts"] == tot["lines"]


def test_wavefront_model_tool_runs():
    """tools/wavefront_model.py (the offline access-pattern model used for DESIGN section 8) at a toy size: both plane-sweep
    launches are traced, the counters are consistent with each other."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "wavefront_model.py"), "--width", "64", "--height", "64",
                        "--views", "2", "--sms", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout)
    for name in ("init", "iter"):
        m = out[name]["model"]
        assert m["requests"] > 0 and m["lines"] <= m["sectors"] <= 4 * m["lines"]
        assert m["wavefronts"] >= max(m["lines"], out[name]["delivered_floor_wavefronts"] - 1)
        assert "ncu" not in out[name]                     # the measured counters belong to the benchmark configuration only


def test_image_pyramid_kernels_vs_cv2_on_cpu():
    """f-4 on the device (csrc/imageprep.cu) executed through the CPU emulation: raw 8-bit image -> 2 x / 255 - 1 ->
    cv2.resize(INTER_LINEAR) -> three coarser levels, against OpenCV itself (1 ulp: its SIMD paths may fuse one product)."""
    cv2 = pytest.importorskip("cv2")
    lib = C.CDLL(cusim_build.build())
    vp, ci = C.c_void_p, C.c_int
    lib.imvs_image_pyramid_u8.restype = ci
    lib.imvs_image_pyramid_u8.argtypes = [vp, ci, ci, vp, vp, vp, vp, ci, ci, vp]
    rng = np.random.default_rng(3)
    for (h0, w0), (w, h) in (((60, 80), (64, 32)), ((48, 64), (64, 48)), ((33, 47), (96, 64))):
        img = rng.integers(0, 256, size=(h0, w0, 3), dtype=np.uint8)
        f = 2 * img.astype(np.float32) / 255. - 1
        want0 = cv2.resize(f, (w, h), interpolation=cv2.INTER_LINEAR)
        outs = [np.full((3, h >> k, w >> k), np.nan, np.float32) for k in range(4)]
        rc = lib.imvs_image_pyramid_u8(img.ctypes.data, h0, w0, *[o.ctypes.data for o in outs], h, w, None)
        assert rc == 0
        assert np.abs(outs[0] - want0.transpose(2, 0, 1)).max() <= 2.5e-7, (h0, w0, w, h)
        for k in (1, 2, 3):
            wantk = cv2.resize(want0, (w >> k, h >> k), interpolation=cv2.INTER_LINEAR)
            assert np.abs(outs[k] - wantk.transpose(2, 0, 1)).max() <= 5e-7, k
