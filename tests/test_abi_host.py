"""CPU-side checks: the C-ABI library loads and exports every symbol include/itermvs_b200.h declares
(no compute calls without a GPU), host-side logic (weight packing, TF32 split, BN folding, synthetic
generator), and the product package never touches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from itermvs_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "itermvs_b200.h")).read()
    declared = set(re.findall(r"\b(imvs_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"imvs_wpair", "imvs_weights", "imvs_problem", "imvs_corrnet_weights", "imvs_featurenet_weights"}
    handle = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, missing
    assert set(_lib.EXPORTED_SYMBOLS) <= declared | {"imvs_launches_total"}
    assert handle.imvs_abi_version() == _lib.ABI_VERSION


def test_host_side_validation_without_gpu():
    from itermvs_b200 import _lib
    L = _lib.lib()
    pb = _lib.Problem(1, 5, 512, 640, 32, 4)
    nbytes = L.imvs_forward_workspace_bytes(C.byref(pb))
    assert 50e6 < nbytes < 200e6
    # fused tcgen05 head: conv0 + one kernel per head call; ConvGRU: operand split + z|r + q on the TMA / tcgen05 kernel
    assert L.imvs_forward_launch_count(C.byref(pb)) == 20 + 12 * 4      # default: one prologue launch (compose x3 + padded level 3)
    for bad in (_lib.Problem(1, 1, 512, 640, 32, 4), _lib.Problem(1, 5, 512, 650, 32, 4), _lib.Problem(1, 5, 512, 640, 30, 4),
                _lib.Problem(0, 5, 512, 640, 32, 4), _lib.Problem(1, 40, 512, 640, 32, 4)):
        assert L.imvs_forward_workspace_bytes(C.byref(bad)) == 0
        assert len(L.imvs_last_error()) > 0
    assert L.imvs_featurenet_workspace_bytes(5, 512, 640) > 0
    assert L.imvs_get_conv_passes() == 4
    # argument errors of the plane-sweep entry points (forward and backward) are reported before anything is launched
    a = 0x1000                                    # a 16-byte aligned non-null address; never dereferenced on these paths
    cases = [
        (L.imvs_warpcorr_init, (None, a, a, a, None, a, 1, 5, 64, 80, 32, None), "null pointer"),
        (L.imvs_warpcorr_init, (a, a, a, a, None, a, 1, 18, 64, 80, 32, None), "source views"),
        (L.imvs_warpcorr_init, (a + 4, a, a, a, None, a, 1, 5, 64, 80, 32, None), "16-byte aligned"),
        (L.imvs_warpcorr_init_backward, (a, a, a, a, None, None, a, 1, 5, 64, 80, 32, None), "null pointer"),
        (L.imvs_warpcorr_init_backward, (a, a, None, None, None, a, a, 1, 5, 64, 80, 32, None), "null pointer"),
        (L.imvs_warpcorr_init_backward, (a, a, a, a, None, a, a, 1, 1, 64, 80, 32, None), "source views"),
        (L.imvs_warpcorr_init_backward, (a, a, a, a, None, a, a, 1, 5, 64, 80, 1, None), "bad shape"),
        (L.imvs_warpcorr_iter_backward, (a, a, a, a, a, a, a, 1, 1, a, a, a, a, None, None, a, a, a, a, 1, 5, 128, 160, None), "either all three"),
        (L.imvs_warpcorr_iter_backward, (a, a, a, a, a, a, None, 0, 1, a, None, None, None, None, None, a, a, a, a, 1, 5, 128, 160, None),
         "either all three"),
        (L.imvs_warpcorr_iter_backward, (a, a, a, a, a, a, a, 1, 1, a, a, a, None, None, None, a, a, a, a, 1, 5, 127, 160, None), "must be even"),
        (L.imvs_warpcorr_iter_backward, (a, a, a, a, a, a, a, 1, 1, a, a, a, None, None, None, a, a, a + 8, a, 1, 5, 128, 160, None),
         "16-byte aligned"),
    ]
    for fn, args, expect in cases:
        assert fn(*args) != 0 and expect in L.imvs_last_error().decode(), (fn.__name__, expect, L.imvs_last_error())


def test_tf32_split_and_packing(dtu_weights):
    from itermvs_b200 import _pack
    w = torch.randn(10000) * torch.logspace(-6, 3, 10000)
    hi, full = _pack.split_tf32(w)
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0          # 10-bit mantissa
    assert float(((w - hi).abs() / w.abs()).max()) <= 2 ** -11 + 1e-9       # round to nearest
    assert torch.equal(full, w)                                             # the 3-pass mode splits fp32 in registers
    # round-half-away on an exact tie
    t = torch.tensor([1.0 + 2 ** -11, -(1.0 + 2 ** -11)])
    assert torch.equal(_pack.round_tf32(t), torch.tensor([1.0 + 2 ** -10, -(1.0 + 2 ** -10)]))
    # conv packing: [Cout,Cin,3,3] -> [9][CinP][CoutP]
    w = dtu_weights["iter_mvs.update.gru.convq.weight"]
    hi, full, um, f16, f16u, f16i = _pack.pack_mma_conv(w, cinp=48)
    # tcgen05 fp16 order [tap][hi|lo][CinK/8][CoutP][8 halves]: hi + lo reproduces the weight to fp32 grade
    assert f16u.shape == (9, 2, 6, 32, 8) and f16u.dtype == torch.float16
    rec16 = (f16u[:, 0].float() + f16u[:, 1].float()).permute(0, 1, 3, 2).reshape(9, 48, 32)
    assert bool(((rec16 - full).abs() <= full.abs() * 2.0 ** -21 + 2.0 ** -24).all())
    assert torch.equal(f16u[4, 0, 1, 7, 3], full[4, 11, 7].half())
    # the TMA + tcgen05 kernel's order [tap][CinK/8][hi|lo][CoutP][8]: rows [0, CoutP) / [CoutP, 2 CoutP) of a chunk are B_hi / B_lo
    assert f16i.shape == (9, 6, 2, 32, 8) and torch.equal(f16i.permute(0, 2, 1, 3, 4), f16u) and f16i.is_contiguous()
    assert um.shape == (1, 9, 12, 32, 4) and torch.equal(um[0, 4, 3, 7], hi[4, 12:16, 7])
    assert hi.shape == (9, 48, 32)
    rec = full[:, :43, :].reshape(3, 3, 43, 32).permute(3, 2, 0, 1)
    assert torch.equal(rec, w)
    assert float(hi[:, 43:, :].abs().max()) == 0.0
    wt = dtu_weights["iter_mvs.evaluation.corr_conv1.0.conv3.weight"]          # ConvTranspose [Cin,Cout,3,3]
    hi, full, _, _, _, _ = _pack.pack_mma_tconv(wt)
    assert hi.shape == (9, 32, 16)
    assert torch.equal(full[4], wt[:, :, 1, 1])


def test_fp16_split_packing(dtu_weights):
    """mode 4 weights: [tap][CinK/2][CoutP][{hi,lo}] int32, each a half2 of channels (k, k+1), k in the low half;
    hi + lo reproduces the fp32 weight to 2^-22 relative (2^-25 absolute floor), saturating beyond fp16 range."""
    from itermvs_b200 import _pack
    w = torch.randn(9, 24, 16) * torch.logspace(-5, 2, 16)
    w[0, 0, 0] = 1e6
    p = _pack.pack_f16x3(w)
    assert p.shape == (9, 16, 16, 2) and p.dtype == torch.int32            # CinK = 32
    halves = p.view(torch.float16).reshape(9, 16, 16, 2, 2)                # [...][hi|lo][k even|k odd]
    rec = (halves[..., 0, :].float() + halves[..., 1, :].float()).permute(0, 1, 3, 2).reshape(9, 32, 16)
    assert float(rec[:, 24:].abs().max()) == 0.0
    assert float(rec[0, 0, 0]) == 65504.0 + 65504.0                        # saturated, finite
    ww, rr = w.clone(), rec[:, :24].clone()
    ww[0, 0, 0] = rr[0, 0, 0] = 0.0
    err = (rr - ww).abs()
    assert bool((err <= ww.abs() * 2.0 ** -21 + 2.0 ** -24).all())
    # a real layer: the packed pair order matches the fp32 packing
    _, full, _, f16, _, _ = _pack.pack_mma_conv(dtu_weights["iter_mvs.update.gru.convq.weight"], cinp=48)
    h = f16.view(torch.float16).reshape(9, 24, 32, 2, 2)
    assert torch.equal(h[4, 5, 7, 0, 1], full[4, 11, 7].half()) and torch.equal(h[4, 5, 7, 0, 0], full[4, 10, 7].half())


def test_fused_head_blob_layout(dtu_weights):
    """csrc/headfused.cuh operand blob: W1 hi | W1 lo | W2 hi | W2 lo in the UMMA K-major canonical order [K/8][N][8 halves],
    then the fp32 bias."""
    from itermvs_b200 import _pack
    w1 = dtu_weights["iter_mvs.update.depth_head.2.weight"]
    w2 = dtu_weights["iter_mvs.update.depth_head.4.weight"]
    b2 = dtu_weights["iter_mvs.update.depth_head.4.bias"]
    blob = _pack.pack_head_fused(w1, w2, b2)
    assert blob.dtype == torch.uint8 and blob.numel() == 2 * 4096 + 2 * 32768 + 1024
    h = blob[:2 * 4096 + 2 * 32768].view(torch.float16)
    w1hi, w1lo = h[:2048].reshape(4, 64, 8), h[2048:4096].reshape(4, 64, 8)
    w2hi, w2lo = h[4096:4096 + 16384].reshape(8, 256, 8), h[4096 + 16384:].reshape(8, 256, 8)
    n, k = 37, 21
    assert torch.equal(w1hi[k // 8, n, k % 8], w1[n, k, 0, 0].half())
    assert abs(float(w1hi[k // 8, n, k % 8]) + float(w1lo[k // 8, n, k % 8]) - float(w1[n, k, 0, 0])) <= abs(float(w1[n, k, 0, 0])) * 2.0 ** -21 + 2.0 ** -24   # fp16-subnormal remainder floor
    n, k = 201, 60
    assert torch.equal(w2hi[k // 8, n, k % 8], w2[n, k, 0, 0].half())
    assert abs(float(w2hi[k // 8, n, k % 8]) + float(w2lo[k // 8, n, k % 8]) - float(w2[n, k, 0, 0])) <= abs(float(w2[n, k, 0, 0])) * 2.0 ** -21 + 2.0 ** -24
    assert torch.equal(blob[-1024:].view(torch.float32), b2)


def test_fp16_split_product_accuracy():
    """The arithmetic behind mode 4, emulated in torch: sum_k (a_hi b_hi + a_hi b_lo + a_lo b_hi) with fp16
    hi/lo parts (exact products, wide accumulation) against the exact dot product, over activation
    scales from fp16-subnormal remainders to near the top of the fp16 range."""
    g = torch.Generator().manual_seed(3)
    K = 432
    b = torch.randn(K, 64, generator=g) * 0.1
    bh = b.half(); bl = (b - bh.float()).half()
    for scale in (2.0 ** -16, 2.0 ** -12, 2.0 ** -6, 1.0, 2.0 ** 10, 2.0 ** 13):
        a = torch.randn(256, K, generator=g) * scale
        ah = a.clamp(-65504, 65504).half(); al = (a - ah.float()).half()     # cvt.rn.satfinite
        d = lambda x: x.double()
        got = d(ah) @ d(bh) + d(ah) @ d(bl) + d(al) @ d(bh)
        ref = d(a) @ d(b)
        denom = (d(a).abs() @ d(b).abs())               # forward-error scale of an fp32 dot product
        rel = float(((got - ref).abs() / denom).max())
        floor = 2.0 ** -25 * float(b.abs().sum(0).max()) / float(denom.min())   # absolute floor of the subnormal remainders
        assert rel < 2.0 ** -20 + 2 * floor, (scale, rel, floor)
        if scale >= 2.0 ** -6:
            assert rel < 2.0 ** -20, (scale, rel)


def test_bn_folding_matches_batchnorm(dtu_weights):
    import torch.nn.functional as F
    from itermvs_b200 import _pack
    sd = {k[len("feature_net."):]: v for k, v in dtu_weights.items() if k.startswith("feature_net.")}
    x = torch.randn(2, 8, 12, 14)
    p = "layer1.0.conv1."
    w, b = _pack.fold_bn(sd, p)
    got = F.conv2d(x, w, b, stride=2, padding=1)
    ref = F.batch_norm(F.conv2d(x, sd[p + "conv.weight"], stride=2, padding=1), sd[p + "bn.running_mean"], sd[p + "bn.running_var"],
                       sd[p + "bn.weight"], sd[p + "bn.bias"], training=False, eps=1e-5)
    assert float((got - ref).abs().max()) < 1e-5
    assert len(_pack.FNET_LAYERS) == 21


def test_state_dict_is_reference_compatible(dtu_weights):
    import itermvs_b200
    m = itermvs_b200.Pipeline(iteration=4, test=True)
    res = m.load_state_dict({"module." + k: v for k, v in dtu_weights.items()}, strict=True)   # DataParallel prefix accepted
    assert not res.missing_keys and not res.unexpected_keys
    ours = {k for k in m.state_dict() if not k.endswith("num_batches_tracked")}
    assert ours == set(dtu_weights)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "itermvs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_cpu_inputs_fail_loudly(dtu_weights):
    import itermvs_b200
    with pytest.raises(RuntimeError, match="CUDA"):
        itermvs_b200.differentiable_warping(torch.zeros(1, 8, 4, 4), torch.eye(4)[None], torch.eye(4)[None], torch.ones(1, 2, 4, 4))


def test_synthetic_generator_is_deterministic():
    from itermvs_b200.synthetic import make_sample, plane_depth_map
    a = make_sample(160, 128, n_src=2, seed=3)
    b = make_sample(160, 128, n_src=2, seed=3)
    assert torch.equal(a["imgs"]["level_0"], b["imgs"]["level_0"])
    assert a["imgs"]["level_0"].shape == (1, 3, 3, 128, 160) and a["proj_matrices"]["level_2"].shape == (1, 3, 4, 4)
    K3 = a["proj_matrices"]["level_3"][0, 0, :3, :3].numpy()
    K0 = a["proj_matrices"]["level_0"][0, 0, :3, :3].numpy()
    np.testing.assert_allclose(K3[:2] * 8, K0[:2], rtol=1e-12)
    gt = plane_depth_map(160, 128)
    assert 600 < gt.min() < gt.max() < 700


def _loss_inputs(z, n_pred=3):
    t = lambda k: torch.from_numpy(z[k])
    depths = {"initial": [t("initial")], "combine": [t(f"combine{i}") for i in range(n_pred)],
              "probability": [t(f"probability{i}") for i in range(n_pred)]}
    conf = [t(f"confidence{i}") for i in range(n_pred)]
    gt = {"level_0": t("gt0"), "level_2": t("gt2")}
    mask = {"level_0": t("mask0"), "level_2": t("mask2")}
    return depths, [t("upsampled")], conf, gt, mask, t("depth_min"), t("depth_max")


def test_full_loss_matches_reference_value(loss_kat):
    """full_loss (net.py:131-190) against the value the reference's own function returned for the same seeded
    predictions (batch 2, masks, gt outside the depth range, arg-max near / far from the gt bin)."""
    from itermvs_b200 import full_loss
    args = _loss_inputs(loss_kat)
    assert abs(float(full_loss(*args)) - float(loss_kat["loss"])) < 1e-5 * float(loss_kat["loss"])
    assert abs(float(full_loss(*args, regress=False)) - float(loss_kat["loss_noregress"])) < 1e-5 * float(loss_kat["loss_noregress"])
    # gradient flows to the predictions the reference trains on (host-side autograd; the CUDA backward is not built)
    depths, ups, conf, gt, mask, dmin, dmax = args
    ups[0].requires_grad_(True)
    conf[-1].requires_grad_(True)
    full_loss(depths, ups, conf, gt, mask, dmin, dmax).backward()
    assert ups[0].grad.abs().sum() > 0 and conf[-1].grad.abs().sum() > 0


def test_differentiable_entry_insists_on_cuda(dtu_weights):
    """test=False with trainable parameters and autograd on -- train() mode, or eval() as in fine-tuning with frozen
    BatchNorm statistics (works in the reference) -- is the differentiable path (itermvs_b200/training.py) and, like
    every entry, insists on CUDA tensors: there is no CPU fallback to fall into silently."""
    import itermvs_b200
    m = itermvs_b200.IterMVS(2, 32, 32, test=False).eval()
    x = {f"level{l}": torch.zeros(1, c, 8, 8) for l, c in ((1, 16), (2, 32), (3, 48))}
    srcs0 = {k: [v.clone()] for k, v in x.items()}
    proj0 = {k: torch.eye(4)[None] for k in x}
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, srcs0, proj0, {k: [v] for k, v in proj0.items()}, torch.ones(1), torch.ones(1))
    m.train()
    srcs = {k: [v.clone()] for k, v in x.items()}
    proj = {k: torch.eye(4)[None] for k in x}
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, srcs, proj, {k: [v] for k, v in proj.items()}, torch.ones(1), torch.ones(1))
