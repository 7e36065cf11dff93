#!/usr/bin/env python
"""Golden fixture for the test=False (all-predictions) forward and for full_loss, made by EXECUTING THE
REFERENCE ITSELF (build container only: needs /root/reference).

    python tests/golden/make_golden_train.py     ->  tests/golden/e2e_allpred_d32.npz

Reference objects used: models.net.Pipeline(iteration=2, test=False) with the DTU checkpoint in eval() mode
under torch.no_grad() -- exactly what train.py's validate_sample (train.py:250-262) runs -- and
models.net.full_loss on its outputs with the synthetic plane's ground-truth depth.  160x128, 2 source
views, plane scene, seed 5.  The 256-bin probability volumes are stored at every 4th pixel only (size).
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
warnings.filterwarnings("ignore")

from models.net import Pipeline, full_loss  # noqa: E402  (the reference)

from itermvs_b200.synthetic import make_sample, plane_depth_map  # noqa: E402

W, H, NSRC, ITERS, SEED = 160, 128, 2, 2, 5


def ground_truth(width, height):
    """depth_gt / mask dicts in the loader's format (datasets/dtu_yao.py: level_0 full, level_2 quarter)."""
    d0 = torch.from_numpy(plane_depth_map(width, height).astype(np.float32))[None, None]
    gt = {"level_0": d0, "level_2": F.interpolate(d0, scale_factor=0.25, mode="nearest")}
    mask = {k: torch.ones_like(v) for k, v in gt.items()}
    mask["level_0"][..., :6, :] = 0          # a masked-out band, so the mask indexing is exercised
    mask["level_2"][..., :2, :] = 0
    return gt, mask


def main():
    torch.set_num_threads(8)
    m = Pipeline(iteration=ITERS, test=False)
    sd = torch.load(os.path.join(REF, "checkpoints/dtu/model_000015.ckpt"), map_location="cpu")["model"]
    m.load_state_dict({k[7:]: v for k, v in sd.items()}, strict=True)
    m.eval()
    s = make_sample(W, H, n_src=NSRC, batch=1, seed=SEED, scene="plane")
    with torch.no_grad():
        out = m(s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"])
        gt, mask = ground_truth(W, H)
        loss = full_loss(out["depths"], out["depths_upsampled"], out["confidences"], gt, mask, s["depth_min"], s["depth_max"])
        loss_noreg = full_loss(out["depths"], out["depths_upsampled"], out["confidences"], gt, mask, s["depth_min"], s["depth_max"],
                               regress=False)
    g = {"width": W, "height": H, "n_src": NSRC, "iteration": ITERS, "seed": SEED,
         "loss": np.float64(loss.item()), "loss_noregress": np.float64(loss_noreg.item()),
         "depth_initial": out["depths"]["initial"][0].numpy(),
         "depths_upsampled": out["depths_upsampled"][0].numpy(),
         "confidence_upsampled": out["confidence_upsampled"].numpy()}
    for i, (d, p, c) in enumerate(zip(out["depths"]["combine"], out["depths"]["probability"], out["confidences"])):
        g[f"combine{i}"] = d.numpy()
        g[f"probability{i}_s4"] = p[:, :, ::4, ::4].contiguous().numpy()
        g[f"probability{i}_argmax"] = p.argmax(1).to(torch.int16).numpy()
        g[f"confidence_logit{i}"] = c.numpy()
    np.savez_compressed(os.path.join(HERE, "e2e_allpred_d32.npz"), **g)
    print("e2e_allpred_d32.npz loss", loss.item(), "noregress", loss_noreg.item(),
          "predictions", len(out["depths"]["combine"]), os.path.getsize(os.path.join(HERE, "e2e_allpred_d32.npz")) / 1e6, "MB")


def loss_kat():
    """full_loss (net.py:131-190) on small seeded random predictions: inputs + the reference's value."""
    torch.manual_seed(77)
    b, h, w, npred = 2, 6, 10, 3
    dmin, dmax = torch.tensor([425.0, 500.0]), torch.tensor([935.0, 900.0])
    inv_min, inv_max = (1.0 / dmin).view(b, 1, 1, 1), (1.0 / dmax).view(b, 1, 1, 1)
    nd_gt2 = torch.rand(b, 1, h, w) * 1.1 - 0.05                       # a few values outside [0,1]: clamp branch
    gt2 = 1.0 / (inv_max + nd_gt2 * (inv_min - inv_max))
    gt0 = F.interpolate(gt2, scale_factor=4, mode="bilinear")
    gt = {"level_0": gt0, "level_2": gt2}
    mask = {"level_0": (torch.rand(b, 1, 4 * h, 4 * w) > 0.2).float(), "level_2": (torch.rand(b, 1, h, w) > 0.2).float()}
    def near(x, sigma):
        return x * (1 + sigma * torch.randn_like(x))
    depths = {"initial": [near(gt2, 0.05)], "combine": [near(gt2, 0.02 / (i + 1)) for i in range(npred)],
              "probability": [torch.softmax(4 * torch.randn(b, 256, h, w), dim=1) for _ in range(npred)]}
    # make the arg-max land near the ground-truth bin for about half of the pixels (mask_2 both ways)
    idx = (nd_gt2.clamp(0, 1) * 255).floor().long()
    for p in depths["probability"]:
        hit = torch.rand(b, 1, h, w) > 0.5
        p.scatter_(1, idx, torch.where(hit, torch.full_like(nd_gt2, 0.9), p.gather(1, idx)))
    confidences = [torch.randn(b, 1, h, w) for _ in range(npred)]
    ups = [near(gt0, 0.01)]
    g = {"depth_min": dmin.numpy(), "depth_max": dmax.numpy(), "gt0": gt0.numpy(), "gt2": gt2.numpy(),
         "mask0": mask["level_0"].numpy(), "mask2": mask["level_2"].numpy(), "initial": depths["initial"][0].numpy(),
         "upsampled": ups[0].numpy()}
    for i in range(npred):
        g[f"combine{i}"] = depths["combine"][i].numpy()
        g[f"probability{i}"] = depths["probability"][i].numpy()
        g[f"confidence{i}"] = confidences[i].numpy()
    g["loss"] = np.float64(full_loss(depths, ups, confidences, gt, mask, dmin, dmax).item())
    g["loss_noregress"] = np.float64(full_loss(depths, ups, confidences, gt, mask, dmin, dmax, regress=False).item())
    np.savez_compressed(os.path.join(HERE, "loss_kat.npz"), **g)
    print("loss_kat.npz loss", g["loss"], "noregress", g["loss_noregress"], os.path.getsize(os.path.join(HERE, "loss_kat.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
    loss_kat()
