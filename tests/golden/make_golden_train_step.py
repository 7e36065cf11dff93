#!/usr/bin/env python
"""Golden fixture for one TRAINING step, made by EXECUTING THE REFERENCE ITSELF (build container only: needs
/root/reference).

    python tests/golden/make_golden_train_step.py     ->  tests/golden/train_step_kat.npz

What runs: models.net.Pipeline(iteration=2, test=False) with the DTU checkpoint in train() mode (BatchNorm batch
statistics), forward on a synthetic plane scene (batch 2, 96x64, 2 source views), models.net.full_loss against the
plane's ground truth, loss.backward() -- the body of train.py:194-215 without the optimizer.  Stored: the loss, every
prediction's depth (the 256-bin volumes only as arg-max), per-parameter gradient norms, the full gradients of a few
small tensors along the path (first FeatureNet conv, the three FPN output convs' biases, PixelViewWeight, one
CorrNet conv, GRU convq bias, head biases) and the updated BatchNorm running statistics of the first layer.
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

W, H, NSRC, ITERS, BATCH, SEED = 96, 64, 2, 2, 2, 9
FULL_GRADS = ["feature_net.conv1.conv.weight", "feature_net.output1.bias", "feature_net.output2.bias", "feature_net.output3.bias",
              "feature_net.inner1.bias", "iter_mvs.evaluation.pixel_view_weight.conv.0.conv.weight",
              "iter_mvs.evaluation.pixel_view_weight.conv.1.bias", "iter_mvs.evaluation.corr_conv1.0.conv5.weight",
              "iter_mvs.evaluation.corr_conv1.2.conv0.conv.weight", "iter_mvs.update.gru.convq.bias",
              "iter_mvs.update.depth_head.4.bias", "iter_mvs.update.confidence_head.2.bias",
              "iter_mvs.update.hidden_init_head.2.bias", "iter_mvs.upsample.2.weight"]


def ground_truth(width, height, batch):
    d0 = torch.from_numpy(plane_depth_map(width, height).astype(np.float32))[None, None].repeat(batch, 1, 1, 1)
    gt = {"level_0": d0, "level_2": F.interpolate(d0, scale_factor=0.25, mode="nearest")}
    mask = {k: torch.ones_like(v) for k, v in gt.items()}
    mask["level_0"][..., :6, :] = 0
    mask["level_2"][..., :2, :] = 0
    return gt, mask


def main():
    torch.set_num_threads(8)
    torch.manual_seed(0)
    m = Pipeline(iteration=ITERS, test=False)
    sd = torch.load(os.path.join(REF, "checkpoints/dtu/model_000015.ckpt"), map_location="cpu")["model"]
    m.load_state_dict({k[7:]: v for k, v in sd.items()}, strict=True)
    m.train()
    s = make_sample(W, H, n_src=NSRC, batch=BATCH, seed=SEED, scene="plane")
    gt, mask = ground_truth(W, H, BATCH)
    out = m(s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"])
    loss = full_loss(out["depths"], out["depths_upsampled"], out["confidences"], gt, mask, s["depth_min"], s["depth_max"])
    loss.backward()
    g = {"width": W, "height": H, "n_src": NSRC, "iteration": ITERS, "batch": BATCH, "seed": SEED,
         "loss": np.float64(loss.item()),
         "depth_initial": out["depths"]["initial"][0].detach().numpy(),
         "depths_upsampled": out["depths_upsampled"][0].detach().numpy(),
         "confidence_upsampled": out["confidence_upsampled"].detach().numpy(),
         "bn_running_mean": m.feature_net.conv1.bn.running_mean.numpy(),
         "bn_running_var": m.feature_net.conv1.bn.running_var.numpy()}
    for i, (d, p, c) in enumerate(zip(out["depths"]["combine"], out["depths"]["probability"], out["confidences"])):
        g[f"combine{i}"] = d.detach().numpy()
        g[f"probability{i}_argmax"] = p.argmax(1).to(torch.int16).numpy()
        g[f"confidence_logit{i}"] = c.detach().numpy()
    names, norms = [], []
    for k, p in m.named_parameters():
        names.append(k)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
        if p.grad is None:
            print("no grad:", k)
    g["grad_names"] = np.array(names)
    g["grad_norms"] = np.array(norms, dtype=np.float64)
    for k in FULL_GRADS:
        g["grad:" + k] = dict(m.named_parameters())[k].grad.numpy()
    path = os.path.join(HERE, "train_step_kat.npz")
    np.savez_compressed(path, **g)
    print("train_step_kat.npz loss", loss.item(), "params", len(names), "total grad norm", float(np.sqrt((g["grad_norms"] ** 2).sum())),
          os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
