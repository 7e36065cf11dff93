"""GPU tests of the training path (SURVEY 8b autograd row): the backward kernels of the fused plane sweep against
torch autograd through the oracle, and one whole training step (Pipeline.train() forward, full_loss, backward)
against the step the REFERENCE ITSELF produced (tests/golden/train_step_kat.npz).

The same checks run on the CPU suite with the kernel sources executed through tests/cusim
(tests/test_cusim_kernels.py, tests/test_training_cpu.py); these are the on-device twins, through the C ABI.
Tolerances of the whole-step comparison are looser than on the CPU twin (measured there: loss 1e-6, gradients
2e-5): the convolutions run in cuDNN here, and an arg-max bin that flips on a pixel moves its regression window.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import itermvs_oracle as O
from itermvs_b200.synthetic import make_sample, plane_depth_map, random_feature_pyramids

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def maxerr(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def _inputs(width, height, n_src, batch, seed):
    ref, srcs = random_feature_pyramids(width, height, n_src, batch, seed)
    s = make_sample(width, height, n_src=n_src, batch=batch, seed=seed, scene="noise")
    rp, sp = {}, {}
    for l in (1, 2, 3):
        pm = torch.unbind(s["proj_matrices"][f"level_{l}"].float(), 1)
        rp[f"level{l}"], sp[f"level{l}"] = pm[0], list(pm[1:])
    return ref, srcs, rp, sp, s


def _cl(ref, srcs, dev):
    return torch.stack([ref] + list(srcs), dim=1).permute(0, 1, 3, 4, 2).contiguous().to(dev)


def _split(g, n_src):
    g = g.permute(0, 1, 4, 2, 3).cpu()
    return g[:, 0], [g[:, v + 1] for v in range(n_src)]


@pytest.mark.parametrize("batch,n_src,d", [(1, 2, 32), (2, 3, 8)])
def test_fused_corr_init_backward(dev, batch, n_src, d):
    from itermvs_b200 import training
    ref, srcs, rp, sp, s = _inputs(160, 128, n_src, batch, seed=21)
    h3, w3 = ref["level3"].shape[2:]
    inv_min = (1.0 / s["depth_min"]).view(batch, 1, 1, 1)
    inv_max = (1.0 / s["depth_max"]).view(batch, 1, 1, 1)
    ds = O.initial_depth_samples(inv_min, inv_max, d, h3, w3)
    ds[:, 1, :2] = -10.0
    r3 = ref["level3"].clone().requires_grad_(True)
    s3 = [t.clone().requires_grad_(True) for t in srcs["level3"]]
    corr = torch.stack([O.group_correlation(O.differentiable_warping(src, p, rp["level3"], ds), r3)
                        for src, p in zip(s3, sp["level3"])], dim=1)                  # [B,S,G,D,H,W]
    gcorr = torch.randn(corr.shape, generator=torch.Generator().manual_seed(4))
    (corr * gcorr).sum().backward()
    fea = _cl(ref["level3"], srcs["level3"], dev).requires_grad_(True)
    rt = training._compose(torch.stack([rp["level3"]] + sp["level3"], dim=1).to(dev))
    out = training.FusedCorrInit.apply(fea, rt, ds.to(dev))
    assert maxerr(out.view(batch, n_src, d, h3, w3, 8).permute(0, 1, 5, 2, 3, 4), corr) < 1e-4
    (out * gcorr.permute(0, 1, 3, 4, 5, 2).reshape(out.shape).to(dev)).sum().backward()
    gref, gsrcs = _split(fea.grad, n_src)
    assert maxerr(gref, r3.grad) < 2e-4 * max(1.0, float(r3.grad.abs().max()))
    for got, t in zip(gsrcs, s3):
        assert maxerr(got, t.grad) < 2e-4 * max(1.0, float(t.grad.abs().max()))


@pytest.mark.parametrize("batch,n_src", [(1, 4), (2, 2), (1, 9)])
def test_fused_corr_iter_backward(dev, batch, n_src):
    from itermvs_b200 import training
    ref, srcs, rp, sp, s = _inputs(160, 128, n_src, batch, seed=22)
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
    aggs = []
    for l in (1, 2, 3):                                   # itermvs.py:86-120 up to (not including) CorrNet
        key = f"level{l}"
        ref_l = O.resample_ref_feature(rg[key], l)
        corr_sum, vw_sum = 0, 1e-5
        for i, (src, p) in enumerate(zip(sg[key], sp[key])):
            c = O.group_correlation(O.differentiable_warping(src, p, rp[key], samples[key]), ref_l)
            v = vw[:, i].reshape(batch, 1, 1, h2, w2)
            corr_sum, vw_sum = corr_sum + c * v, vw_sum + v
        aggs.append(corr_sum / vw_sum)
    want = torch.cat(aggs, dim=2)                                                    # [B,8,10,H2,W2]
    gagg = torch.randn(want.shape, generator=g)
    (want * gagg).sum().backward()
    feas = [_cl(ref[f"level{l}"], srcs[f"level{l}"], dev).requires_grad_(True) for l in (1, 2, 3)]
    rts = [training._compose(torch.stack([rp[f"level{l}"]] + sp[f"level{l}"], dim=1).to(dev)) for l in (1, 2, 3)]
    out = training.FusedCorrIter.apply(feas[0], feas[1], feas[2], rts[0], rts[1], rts[2],
                                       *[samples[f"level{l}"].to(dev).contiguous() for l in (1, 2, 3)], vw.to(dev))
    assert maxerr(out.view(batch, 10, h2, w2, 8).permute(0, 4, 1, 2, 3), want) < 1e-4
    (out * gagg.permute(0, 2, 3, 4, 1).reshape(out.shape).to(dev)).sum().backward()
    for l in (1, 2, 3):
        gref, gsrcs = _split(feas[l - 1].grad, n_src)
        wr = rg[f"level{l}"].grad
        assert maxerr(gref, wr) < 2e-4 * max(1.0, float(wr.abs().max())), l
        for got, t in zip(gsrcs, sg[f"level{l}"]):
            assert maxerr(got, t.grad) < 2e-4 * max(1.0, float(t.grad.abs().max())), l


def _ground_truth(width, height, batch):
    d0 = torch.from_numpy(plane_depth_map(width, height).astype(np.float32))[None, None].repeat(batch, 1, 1, 1)
    gt = {"level_0": d0, "level_2": F.interpolate(d0, scale_factor=0.25, mode="nearest")}
    mask = {k: torch.ones_like(v) for k, v in gt.items()}
    mask["level_0"][..., :6, :] = 0
    mask["level_2"][..., :2, :] = 0
    return gt, mask


def test_training_step_matches_reference(dev, dtu_weights):
    import itermvs_b200
    from itermvs_b200.ddp import FlatBucketDDP, train_step
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_step_kat.npz")) as z:
        kat = {k: z[k] for k in z.files}
    w, h, n_src, iters, batch, seed = (int(kat[k]) for k in ("width", "height", "n_src", "iteration", "batch", "seed"))
    m = itermvs_b200.Pipeline(iteration=iters, test=False)
    m.load_state_dict(dtu_weights, strict=True)
    m = m.to(dev).train()
    s = make_sample(w, h, n_src=n_src, batch=batch, seed=seed, scene="plane")
    gt, mask = _ground_truth(w, h, batch)
    to = lambda d: {k: v.to(dev) for k, v in d.items()}
    sample = {"imgs": to(s["imgs"]), "proj_matrices": to(s["proj_matrices"]), "depth_min": s["depth_min"].to(dev),
              "depth_max": s["depth_max"].to(dev), "depth": to(gt), "mask": to(mask)}
    out = m(sample["imgs"], sample["proj_matrices"], sample["depth_min"], sample["depth_max"])       # Pipeline.forward, train()
    agree = [float((out["depths"]["probability"][i].argmax(1).cpu().numpy() == kat[f"probability{i}_argmax"]).mean())
             for i in range(iters + 1)]
    assert min(agree) > 0.98, agree
    loss = itermvs_b200.full_loss(out["depths"], out["depths_upsampled"], out["confidences"], sample["depth"], sample["mask"],
                                  sample["depth_min"], sample["depth_max"])
    assert abs(loss.item() - float(kat["loss"])) < 1e-2 * float(kat["loss"]), (loss.item(), float(kat["loss"]))
    loss.backward()
    params = dict(m.named_parameters())
    total_ref = float(np.sqrt((kat["grad_norms"] ** 2).sum()))
    total = float(np.sqrt(sum(float(p.grad.double().norm()) ** 2 for p in params.values() if p.grad is not None)))
    assert abs(total - total_ref) < 5e-2 * total_ref, (total, total_ref)
    assert params["feature_net.inner3.weight"].grad is None                      # unused in forward (net.py:25)
    g, gref = params["feature_net.conv1.conv.weight"].grad.cpu(), torch.from_numpy(kat["grad:feature_net.conv1.conv.weight"])
    assert float((g - gref).abs().max()) < 5e-2 * float(gref.abs().max())         # the end of the chain: through the plane sweep
    print(f"\n[training parity on device] argmax agreement {agree}, loss {loss.item():.6f} vs {float(kat['loss']):.6f}, "
          f"total grad norm {total:.5f} vs {total_ref:.5f}")
    # one optimizer step through the single-process DDP wrapper (train.py:194-215); parameters move, loss is finite
    ddp = FlatBucketDDP(m)
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.999))
    before = params["iter_mvs.update.gru.convq.weight"].detach().clone()
    step_loss, _ = train_step(ddp, opt, sample, itermvs_b200.full_loss)
    assert torch.isfinite(step_loss) and not torch.equal(before, params["iter_mvs.update.gru.convq.weight"].detach())


def test_modules_are_differentiable_in_train_mode(dev, dtu_weights):
    """Operator level: Evaluation (reference-style dict arguments, both branches) and Update in train() mode carry a graph
    and match the oracle; the same modules in eval() mode run the inference kernels and give the same numbers."""
    import itermvs_b200
    m = itermvs_b200.Pipeline(iteration=1, test=False)
    m.load_state_dict(dtu_weights, strict=True)
    m = m.to(dev).train()
    ev, upd = m.iter_mvs.evaluation, m.iter_mvs.update
    batch, n_src = 1, 2
    ref, srcs, rp, sp, s = _inputs(160, 128, n_src, batch, seed=5)
    to = lambda d: {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in d.items()}
    gref, gsrcs, grp, gsp = to(ref), to(srcs), to(rp), to(sp)
    gref = {k: v.requires_grad_(True) for k, v in gref.items()}
    inv_min = (1.0 / s["depth_min"]).view(batch, 1, 1, 1)
    inv_max = (1.0 / s["depth_max"]).view(batch, 1, 1, 1)
    ds = O.initial_depth_samples(inv_min, inv_max, 32, 16, 20)
    vw, corr, depth = ev(gref, gsrcs, grp, gsp, ds.to(dev), inv_min.to(dev), inv_max.to(dev))
    want = O.evaluation_init(dtu_weights, ref["level3"], srcs["level3"], rp["level3"], sp["level3"], ds, inv_min, inv_max)
    assert maxerr(vw, want["view_weights"]) < 1e-4 and maxerr(corr, want["corr"]) < 2e-4
    assert corr.requires_grad and vw.requires_grad
    hidden = upd.hidden_init(corr)
    nd0, _ = upd.depth_init(hidden)
    samples = {f"level{l}": O.iteration_depth_samples(nd0.detach().cpu(), l, inv_min, inv_max).to(dev) for l in (1, 2, 3)}
    corr_it = ev(gref, gsrcs, grp, gsp, samples, view_weights=vw.detach())
    hidden1, nd1, prob1, conf1, conf01 = upd(hidden, nd0.detach(), corr_it, confidence_flag=True)
    (nd1.sum() + conf01.sum() + depth.sum()).backward()
    assert all(v.grad is not None and float(v.grad.abs().sum()) > 0 for v in gref.values())
    assert upd.gru.convq.weight.grad is not None and ev.pixel_view_weight.conv[0].conv.weight.grad is not None
    m.eval()
    with torch.no_grad():
        vw_e, corr_e, _ = ev(gref, gsrcs, grp, gsp, ds.to(dev), inv_min.to(dev), inv_max.to(dev))
    assert maxerr(vw_e, vw) < 1e-4 and maxerr(corr_e, corr) < 2e-4           # inference kernels == differentiable path
