"""Training path (itermvs_b200/training.py) against a training step of the REFERENCE ITSELF
(tests/golden/train_step_kat.npz, made by tests/golden/make_golden_train_step.py: Pipeline.train() forward,
full_loss, backward with the DTU checkpoint).

On this CPU-only suite the fused plane-sweep operators (forward and backward CUDA kernels) execute through
tests/cusim -- the real kernel sources, emulated -- by pointing training.py's three backend hooks at the simulation;
the convolution stacks are the same ATen modules the GPU path uses.  The GPU twin of this test is
tests/test_gpu_training.py.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cusim"))
import cusim_build  # noqa: E402

from itermvs_b200.synthetic import make_sample, plane_depth_map  # noqa: E402


def _ground_truth(width, height, batch):
    d0 = torch.from_numpy(plane_depth_map(width, height).astype(np.float32))[None, None].repeat(batch, 1, 1, 1)
    gt = {"level_0": d0, "level_2": F.interpolate(d0, scale_factor=0.25, mode="nearest")}
    mask = {k: torch.ones_like(v) for k, v in gt.items()}
    mask["level_0"][..., :6, :] = 0
    mask["level_2"][..., :2, :] = 0
    return gt, mask


@pytest.fixture()
def sim_backend(monkeypatch):
    from itermvs_b200 import _lib, training
    lib = C.CDLL(cusim_build.build())
    for name in ("imvs_compose_projections", "imvs_warpcorr_init", "imvs_warpcorr_iter", "imvs_warpcorr_init_backward",
                 "imvs_warpcorr_iter_backward", "imvs_last_error"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib._SIGNATURES[name]
    monkeypatch.setattr(training, "_L", lambda: lib)
    monkeypatch.setattr(training, "_st", lambda: None)
    monkeypatch.setattr(training, "_chk", lambda t, name: t.float().contiguous())
    return lib


@pytest.fixture(scope="module")
def kat():
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_step_kat.npz")) as z:
        return {k: z[k] for k in z.files}


def test_training_step_matches_reference(sim_backend, kat, dtu_weights):
    import itermvs_b200
    from itermvs_b200 import training
    w, h, n_src, iters, batch, seed = (int(kat[k]) for k in ("width", "height", "n_src", "iteration", "batch", "seed"))
    torch.manual_seed(0)
    m = itermvs_b200.Pipeline(iteration=iters, test=False)
    m.load_state_dict(dtu_weights, strict=True)
    m.train()
    s = make_sample(w, h, n_src=n_src, batch=batch, seed=seed, scene="plane")
    gt, mask = _ground_truth(w, h, batch)
    out = training.pipeline_train_forward(m, s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"])
    # forward: every prediction of the training structure (net.py:115-120)
    assert len(out["depths"]["combine"]) == iters + 1 and len(out["confidences"]) == iters + 1
    rel = lambda a, b: float(((a - b).abs() / b.abs().clamp_min(1e-6)).max())
    assert rel(out["depths"]["initial"][0].detach(), torch.from_numpy(kat["depth_initial"])) < 1e-3
    for i in range(iters + 1):
        same = (out["depths"]["probability"][i].argmax(1).numpy() == kat[f"probability{i}_argmax"])
        assert same.mean() > 0.995, (i, same.mean())
        d, dref = out["depths"]["combine"][i].detach(), torch.from_numpy(kat[f"combine{i}"])
        ok = torch.from_numpy(same).unsqueeze(1)
        assert rel(d[ok], dref[ok]) < 1e-3, i
        c, cref = out["confidences"][i].detach(), torch.from_numpy(kat[f"confidence_logit{i}"])
        assert float((c - cref).abs()[ok].max()) < 5e-3, i
    loss = itermvs_b200.full_loss(out["depths"], out["depths_upsampled"], out["confidences"], gt, mask, s["depth_min"], s["depth_max"])
    assert abs(loss.item() - float(kat["loss"])) < 1e-4 * float(kat["loss"])       # measured 1e-6
    # BatchNorm batch statistics were used and the running statistics updated, as in the reference's train() mode
    assert np.allclose(m.feature_net.conv1.bn.running_mean.numpy(), kat["bn_running_mean"], rtol=1e-4, atol=1e-6)
    assert np.allclose(m.feature_net.conv1.bn.running_var.numpy(), kat["bn_running_var"], rtol=1e-4, atol=1e-6)
    # backward: through the CUDA backward kernels of the plane sweep into FeatureNet
    loss.backward()
    params = dict(m.named_parameters())
    total_ref = float(np.sqrt((kat["grad_norms"] ** 2).sum()))
    worst = 0.0
    for name, ref_norm in zip(kat["grad_names"], kat["grad_norms"]):
        p = params[str(name)]
        got = 0.0 if p.grad is None else float(p.grad.double().norm())
        if ref_norm == 0.0:
            assert got == 0.0, name                                  # feature_net.inner3: unused in forward (net.py:25)
            continue
        worst = max(worst, abs(got - ref_norm) / max(ref_norm, 1e-3 * total_ref))
    assert worst < 1e-3, worst           # measured 2e-5
    for key in kat:
        if not key.startswith("grad:"):
            continue
        g, gref = params[key[5:]].grad, torch.from_numpy(kat[key])
        assert float((g - gref).abs().max()) < 1e-3 * max(float(gref.abs().max()), 1e-3 * total_ref), key


def test_train_mode_dispatch_and_errors(dtu_weights):
    """Pipeline.forward in train() mode goes to the training path (and, like every entry, refuses CPU tensors: the
    fused operators have no CPU implementation in the product)."""
    import itermvs_b200
    m = itermvs_b200.Pipeline(iteration=1, test=False)
    m.train()
    s = make_sample(64, 64, n_src=1, batch=1, seed=1, scene="plane")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"])


def test_train_step_with_flat_bucket(sim_backend, dtu_weights):
    """train_step (train.py:194-215) through FlatBucketDDP on the real model: gradients land in the bucket, the
    clipped Adam step moves the parameters, the unused FPN branch stays where it was, a second step reuses the views."""
    import itermvs_b200
    from itermvs_b200 import training
    from itermvs_b200.ddp import FlatBucketDDP, train_step

    class OnSim(torch.nn.Module):               # Pipeline.forward itself insists on CUDA tensors
        def __init__(self, pipe):
            super().__init__()
            self.pipe = pipe

        def forward(self, imgs, proj, dmin, dmax):
            return training.pipeline_train_forward(self.pipe, imgs, proj, dmin, dmax)

    m = itermvs_b200.Pipeline(iteration=1, test=False)
    m.load_state_dict(dtu_weights, strict=True)
    s = make_sample(64, 64, n_src=2, batch=1, seed=3, scene="plane")
    gt, mask = _ground_truth(64, 64, 1)
    sample = dict(s, depth=gt, mask=mask)
    ddp = FlatBucketDDP(OnSim(m))
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4)
    w0 = {k: p.detach().clone() for k, p in m.named_parameters()}
    l1, _ = train_step(ddp, opt, sample, itermvs_b200.full_loss)
    assert torch.isfinite(l1) and float(ddp.gradient_bucket.abs().sum()) > 0
    assert float(torch.sqrt((ddp.gradient_bucket.double() ** 2).sum())) <= 2.0 + 1e-4        # clip_grad_norm_(…, 2.0)
    l2, _ = train_step(ddp, opt, sample, itermvs_b200.full_loss)
    assert torch.isfinite(l2)
    moved = [k for k, p in m.named_parameters() if not torch.equal(p.detach(), w0[k])]
    assert "feature_net.conv1.conv.weight" in moved and "iter_mvs.update.gru.convq.weight" in moved
    assert not any(k.startswith("feature_net.inner3") for k in moved)


def test_modules_are_differentiable_in_train_mode(sim_backend, dtu_weights, monkeypatch):
    """Operator level: in train() mode with autograd on, the mirrored modules -- Evaluation (both branches, the
    reference's dict-of-NCHW arguments), Update, ConvGRU, CorrNet, PixelViewWeight -- return tensors that carry a graph,
    with the same values as the oracle; in eval() mode / under no_grad they stay on the inference kernels."""
    import itermvs_b200
    from itermvs_b200 import ops
    from itermvs_b200.synthetic import random_feature_pyramids
    from oracle import itermvs_oracle as O
    monkeypatch.setattr(ops, "_chk", lambda t, name: t.float().contiguous())
    m = itermvs_b200.Pipeline(iteration=1, test=False)
    m.load_state_dict(dtu_weights, strict=True)
    m.train()
    ev, upd = m.iter_mvs.evaluation, m.iter_mvs.update
    w, h, n_src = 64, 64, 2
    ref, srcs = random_feature_pyramids(w, h, n_src, 1, 5)
    s = make_sample(w, h, n_src=n_src, batch=1, seed=5, scene="noise")
    rp, sp = {}, {}
    for l in (1, 2, 3):
        pm = torch.unbind(s["proj_matrices"][f"level_{l}"].float(), 1)
        rp[f"level{l}"], sp[f"level{l}"] = pm[0], list(pm[1:])
    ref = {k: v.requires_grad_(True) for k, v in ref.items()}
    inv_min, inv_max = (1.0 / s["depth_min"]).view(1, 1, 1, 1), (1.0 / s["depth_max"]).view(1, 1, 1, 1)
    ds = O.initial_depth_samples(inv_min, inv_max, 32, h // 8, w // 8)
    vw, corr, depth = ev(ref, srcs, rp, sp, ds, inv_min, inv_max)
    want = O.evaluation_init(dtu_weights, ref["level3"].detach(), srcs["level3"], rp["level3"], sp["level3"], ds, inv_min, inv_max)
    assert float((vw.detach() - want["view_weights"]).abs().max()) < 2e-5 and float((corr.detach() - want["corr"]).abs().max()) < 1e-4
    assert corr.requires_grad and vw.requires_grad and depth.requires_grad
    nd = torch.rand(1, 1, h // 4, w // 4, generator=torch.Generator().manual_seed(1))
    samples = {f"level{l}": O.iteration_depth_samples(nd, l, inv_min, inv_max) for l in (1, 2, 3)}
    corr_it = ev(ref, srcs, rp, sp, samples, view_weights=vw.detach())
    want_it = O.evaluation_iter(dtu_weights, {k: v.detach() for k, v in ref.items()}, srcs, rp, sp, samples, vw.detach())
    assert float((corr_it.detach() - want_it).abs().max()) < 1e-4 and corr_it.requires_grad
    hidden = upd.hidden_init(corr)
    nd0, prob = upd.depth_init(hidden)
    conf, conf0 = upd.conf_init(hidden)
    hidden1, nd1, prob1, conf1, conf01 = upd(hidden, nd0.detach(), corr_it, confidence_flag=True)
    assert all(t.requires_grad for t in (hidden, nd0, prob, conf0, hidden1, nd1, prob1, conf01))
    (nd1.sum() + conf01.sum() + depth.sum()).backward()
    assert all(v.grad is not None and float(v.grad.abs().sum()) > 0 for v in ref.values())
    assert upd.gru.convq.weight.grad is not None and ev.corr_conv1[0].conv5.weight.grad is not None
    assert ev.pixel_view_weight.conv[0].conv.weight.grad is not None
    # the small modules on their own
    x = torch.randn(1, 8, 3, 8, 8)
    assert ev.corr_conv1[1](x).requires_grad and ev.pixel_view_weight(x).requires_grad
    assert upd.gru(torch.randn(1, 32, 8, 8), torch.randn(1, 11, 8, 8)).requires_grad


def test_nan_projection_raises_like_the_reference(sim_backend, dtu_weights):
    """module.py:83-87: `assert not torch.isnan(proj).any()` -- the training path keeps the AssertionError."""
    import itermvs_b200
    from itermvs_b200 import training
    m = itermvs_b200.Pipeline(iteration=1, test=False)
    m.load_state_dict(dtu_weights, strict=True)
    m.train()
    s = make_sample(64, 64, n_src=1, batch=1, seed=1, scene="plane")
    s["proj_matrices"]["level_2"][0, 1, 0, 0] = float("nan")
    with pytest.raises(AssertionError, match="nan in proj"):
        training.pipeline_train_forward(m, s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"])
