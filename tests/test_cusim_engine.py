"""The tensor-core convolution engine and the whole inference forward on the CPU suite: every *.cu of the library
except the tcgen05 variant is compiled by tests/cusim (mma.sync / cp.async / f16 conversions have C++ stand-ins with
the PTX fragment layouts, `#ifdef CUSIM` in csrc/mmaconv.cuh) and executed thread by thread, through the SAME Python
mirror of the reference interface the GPU uses, against the reference-generated fixtures.

How: for the duration of a test, `itermvs_b200._lib.lib()` hands out the emulation and the two tensor checks of
`itermvs_b200.ops` accept CPU tensors.  That substitution exists only in this file -- the product has no CPU path
(tests/test_abi_host.py::test_cpu_inputs_fail_loudly, tests/test_cusim_kernels.py::test_sim_is_not_the_product_library).
The stage tests below CALL THE BODIES of the GPU parity tests (tests/test_gpu_parity.py) with device = cpu, so the
two suites cannot drift apart.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cusim"))
import cusim_build  # noqa: E402

import test_gpu_parity as G  # noqa: E402   (function bodies only; its `gpu` mark applies to its own collection)
from itermvs_b200.synthetic import make_sample  # noqa: E402

CPU = torch.device("cpu")


@pytest.fixture()
def sim_product(monkeypatch):
    from itermvs_b200 import _lib, ops
    lib = C.CDLL(cusim_build.build())
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(lib, name)                    # every symbol of the C ABI exists in the emulated build as well
        fn.restype, fn.argtypes = res, args
    assert lib.imvs_abi_version() == _lib.ABI_VERSION

    def chk(t, name):
        if t.dtype != torch.float32:
            raise TypeError(f"itermvs_b200: {name} must be float32, got {t.dtype}")
        return t.contiguous()
    monkeypatch.setattr(_lib, "_lib", lib)
    monkeypatch.setattr(ops, "_chk", chk)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    return lib


@pytest.fixture()
def model(sim_product, dtu_weights):
    import itermvs_b200
    m = itermvs_b200.Pipeline(iteration=4, test=True)
    m.load_state_dict(dtu_weights, strict=True)
    return m.eval()


def _aligned_bytes(n):
    """The C ABI asks for 256-byte aligned workspaces (cudaMalloc granularity); torch's CPU allocator gives 64."""
    t = torch.empty(n + 256, dtype=torch.uint8)
    off = (-t.data_ptr()) % 256
    return t[off:off + n]


def test_conv_stages_on_cpu(model, stage_kats, dtu_weights):
    """CorrNet (strided + transposed convolutions, U-Net skips), PixelViewWeight, ConvGRU (gate epilogues), hidden_init,
    the depth / confidence heads with the softmax + arg-max + window regression, the clamped-window edge cases."""
    G.test_corrnet_pvw_gru_hinit_golden(CPU, stage_kats, model)
    G.test_heads_probability_and_window_regression(CPU, stage_kats, model, dtu_weights)
    G.test_window_regression_edges(CPU, model)
    # the tile-resident single-launch CorrNet (csrc/corrnet_tile.cuh, IMVS_TUNE_CORR_TILE=1) against the same KATs
    os.environ["IMVS_TUNE_CORR_TILE"] = "1"
    try:
        ev = model.iter_mvs.evaluation
        for i in range(3):
            out = ev.corr_conv1[i](G.T(stage_kats["corrnet_in"]))
            assert G.maxerr(out, G.T(stage_kats[f"corrnet{i}_out"])) < 2e-5
    finally:
        del os.environ["IMVS_TUNE_CORR_TILE"]


def test_upsample_outputs_on_cpu(model, dtu_weights):
    G.test_upsample_outputs(CPU, model, dtu_weights)


def test_end_to_end_fixture_on_cpu(sim_product, dtu_weights):
    """FeatureNet + the whole estimator (imvs_featurenet_forward + imvs_itermvs_forward: 18 + 50 launches) against the
    output of the reference itself (tests/golden/e2e_tiny.npz: 64x64, 1 source view, D=32, 2 iterations)."""
    import itermvs_b200
    from itermvs_b200 import _lib
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "e2e_tiny.npz")) as z:
        fix = {k: z[k] for k in z.files}
    w, h, n_src, iters, d = (int(fix[k]) for k in ("width", "height", "n_src", "iteration", "num_sample"))
    m = itermvs_b200.Pipeline(iteration=iters, test=True)
    m.load_state_dict(dtu_weights, strict=True)
    m.eval()
    s = make_sample(w, h, n_src=n_src, batch=1, seed=int(fix["seed"]), scene="plane")
    x = s["imgs"]["level_0"]
    assert abs(float(x.double().sum()) - float(fix["img_checksum"][0])) < 1e-6 * max(1.0, abs(float(fix["img_checksum"][0])))
    n = x.shape[0] * x.shape[1]
    m.feature_net._ws[0] = ((n, h, w, "cpu"), _aligned_bytes(sim_product.imvs_featurenet_workspace_bytes(n, h, w)))
    pb = _lib.Problem(1, n_src + 1, h, w, d, iters)
    m.iter_mvs._workspaces[0] = ((1, n_src + 1, h, w, d, iters, "cpu"),
                                 _aligned_bytes(sim_product.imvs_forward_workspace_bytes(C.byref(pb))))
    before = sim_product.imvs_launches_total()
    with torch.no_grad():
        f1, f2, f3 = m.feature_net.forward_nhwc(x)
        # FeatureNet on its own against the reference's feature maps
        for got, key in ((f2[:, 0], "ref_level2"), (f3[:, 0], "ref_level3"), (f3[:, 1], "src0_level3")):
            want = torch.from_numpy(fix[key])
            assert float((got.permute(0, 3, 1, 2) - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max())), key
        projs = [s["proj_matrices"][f"level_{l}"].float().contiguous() for l in (1, 2, 3)]
        depth, depth_up, conf, conf_up = m.iter_mvs.forward_packed(f1, f2, f3, projs[0], projs[1], projs[2],
                                                                    s["depth_min"].float(), s["depth_max"].float())
    launches = sim_product.imvs_launches_total() - before
    assert launches == sim_product.imvs_featurenet_launch_count() + sim_product.imvs_forward_launch_count(C.byref(pb))
    rel = np.abs(depth_up.numpy() - fix["depths_upsampled"]) / fix["depths_upsampled"]
    assert float(np.median(rel)) < 2e-5 and float((rel > 1e-3).mean()) < 0.03            # the GPU test's bounds
    assert float((np.abs(conf_up.numpy() - fix["confidence_upsampled"]) > 1e-3).mean()) < 0.03


@pytest.mark.parametrize("passes", [3, 1])
def test_tf32_modes_on_cpu(sim_product, model, stage_kats, passes):
    """The other two precision modes of the convolution engine (mma.sync.m16n8k8 TF32): 3 = 3xTF32 error-compensated
    (fp32-grade: the GPU test's bounds), 1 = single-pass TF32 (the reference's stock cuDNN precision; 10-bit mantissa
    bounds).  tcgen05 is switched off: that variant is the one source the emulation does not cover."""
    from itermvs_b200 import _lib
    sim_product.imvs_set_tcgen05(0)
    _lib.set_conv_passes(passes)
    try:
        if passes == 3:
            G.test_corrnet_pvw_gru_hinit_golden(CPU, stage_kats, model)
        else:
            k = stage_kats
            h = model.iter_mvs.update.gru(G.T(k["gru_h"]), G.T(k["gru_x"]))
            assert G.maxerr(h, G.T(k["gru_out"])) < 2e-2          # |h| <= 1, K = 387 products of 10-bit operands
            assert G.maxerr(h, G.T(k["gru_out"])) > 1e-6          # ... and it really ran in reduced precision
    finally:
        _lib.set_conv_passes(4)
        sim_product.imvs_set_tcgen05(1)


def test_results_do_not_depend_on_the_thread_schedule(sim_product, model, stage_kats, monkeypatch):
    """CUSIM_SHUFFLE: the fibers of a block run in a random order that changes every scheduling round.  A missing barrier
    between the warps that stage a tile / weights and the warps that consume them would make the output depend on the
    seed; it must be bit-identical."""
    k = stage_kats
    upd, ev = model.iter_mvs.update, model.iter_mvs.evaluation

    def run():
        with torch.no_grad():
            return (upd.gru(G.T(k["gru_h"]), G.T(k["gru_x"])), ev.corr_conv1[0](G.T(k["corrnet_in"])),
                    upd.hidden_init(G.T(k["hinit_in"])))
    base = run()
    for seed in ("1", "2"):
        monkeypatch.setenv("CUSIM_SHUFFLE", seed)
        for a, b in zip(base, run()):
            assert torch.equal(a, b), seed
    monkeypatch.delenv("CUSIM_SHUFFLE")
