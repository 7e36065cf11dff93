"""Pin the CPU oracle (oracle/itermvs_oracle.py) against vectors produced by the reference
itself (tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest
import torch

from oracle import itermvs_oracle as O
from itermvs_b200.synthetic import make_sample


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, atol, rtol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    lim = atol + rtol * np.abs(b)
    assert (err <= lim).all(), f"max abs err {err.max():.3e} (limit {lim.min():.3e}), at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.mark.parametrize("tag", ["same", "fea2x", "fea_half", "b2"])
def test_warp_matches_reference(stage_kats, tag):
    k = stage_kats
    out = O.differentiable_warping(T(k[f"warp_{tag}_fea"]), T(k[f"warp_{tag}_src_proj"]),
                                   T(k[f"warp_{tag}_ref_proj"]), T(k[f"warp_{tag}_depth"]))
    # explicit 4-tap restatement vs ATen grid_sample: fp32 reassociation only
    close(out.numpy(), k[f"warp_{tag}_out"], atol=2e-5)


def test_resamplers_match_interpolate(stage_kats):
    x = T(stage_kats["interp_in"])
    close(O.bilinear_up(x, 2).numpy(), stage_kats["interp_up2"], atol=1e-6)
    close(O.bilinear_up(x, 4).numpy(), stage_kats["interp_up4"], atol=1e-6)
    close(O.mean_pool2(x).numpy(), stage_kats["interp_half"], atol=1e-6)


def test_gru_corrnet_pvw_heads(stage_kats, dtu_weights):
    k, w = stage_kats, dtu_weights
    close(O.conv_gru(w, T(k["gru_h"]), T(k["gru_x"])).numpy(), k["gru_out"], atol=2e-6)
    for i in range(3):
        out = O.corr_net(w, T(k["corrnet_in"]), f"iter_mvs.evaluation.corr_conv1.{i}.")
        close(out.numpy(), k[f"corrnet{i}_out"], atol=2e-6)
    close(O.pixel_view_weight(w, T(k["pvw_in"])).numpy(), k["pvw_out"], atol=1e-6)
    close(O.hidden_init(w, T(k["hinit_in"])).numpy(), k["hinit_out"], atol=2e-6)
    close(O.depth_head_logits(w, T(k["gru_h"])).numpy(), k["head_logits"], atol=5e-6)
    close(O.confidence_logit(w, T(k["gru_h"])).numpy(), k["conf_logit"], atol=5e-6)


def test_window_regression_and_upsample(stage_kats):
    k = stage_kats
    prob = torch.softmax(T(k["regress_logits"]), dim=1)
    close(prob.numpy(), k["regress_prob"], atol=1e-7)
    close(O.window_regression(prob).numpy(), k["regress_nd"], atol=1e-6)
    close(O.convex_upsample(T(k["upsample_x"]), T(k["upsample_w"])).numpy(), k["upsample_out"], atol=1e-6)


def _run_e2e(fix, weights):
    w = dict(weights)
    for key, v in fix.items():
        if key.startswith("extra:"):
            w[key[6:]] = T(v)
    s = make_sample(int(fix["width"]), int(fix["height"]), n_src=int(fix["n_src"]), batch=1,
                    seed=int(fix["seed"]), scene="plane")
    chk = np.array([float(s["imgs"]["level_0"].double().sum()), float(s["imgs"]["level_0"].double().abs().sum())])
    np.testing.assert_allclose(chk, fix["img_checksum"], rtol=1e-9, err_msg="synthetic generator drifted")
    trace = {}
    out = O.pipeline_forward(w, s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"],
                             iteration=int(fix["iteration"]), num_sample=int(fix["num_sample"]), trace=trace)
    return out, trace


@pytest.mark.parametrize("which", ["e2e_d8", "e2e_d32"])
def test_pipeline_matches_reference(which, request, dtu_weights):
    fix = request.getfixturevalue(which)
    out, tr = _run_e2e(fix, dtu_weights)
    # features and the continuous stages: tight
    close(tr["ref_feature"]["level3"].numpy(), fix["ref_level3"], atol=2e-5)
    close(tr["ref_feature"]["level2"].numpy(), fix["ref_level2"], atol=2e-5)
    close(tr["src_features"]["level3"][0].numpy(), fix["src0_level3"], atol=2e-5)
    close(tr["view_weights"].numpy(), fix["view_weights"], atol=2e-5)
    close(tr["corr_init"].numpy(), fix["corr_init"], atol=5e-5)
    close(tr["hidden0"].numpy(), fix["hidden0"], atol=5e-5)
    # stages behind the arg-max: compare in relative depth, allow isolated bin flips
    def frac_bad(a, b, tol):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float((np.abs(a - b) > tol * np.maximum(np.abs(b), 1e-6)).mean())
    assert frac_bad(tr["nd0"].numpy(), fix["nd0"], 1e-3) < 0.005
    for it in range(int(fix["iteration"])):
        assert frac_bad(tr["nd_iter"][it].numpy(), fix[f"nd_iter{it}"], 1e-3) < 0.01
    d, dref = out["depths_upsampled"].numpy(), fix["depths_upsampled"]
    assert d.shape == dref.shape
    assert frac_bad(d, dref, 1e-3) < 0.01, frac_bad(d, dref, 1e-3)
    assert np.median(np.abs(d - dref) / dref) < 1e-5
    c, cref = out["confidence_upsampled"].numpy(), fix["confidence_upsampled"]
    assert float((np.abs(c - cref) > 1e-3).mean()) < 0.01
