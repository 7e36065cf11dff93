#!/usr/bin/env python
"""Golden outputs of THE REFERENCE ITSELF at the benchmark shapes (build container only; needs /root/reference).

    python tests/golden/make_golden_cfg.py

    e2e_cfg2.npz   BASELINE configs[1]: 640x512, 4 src views, D=32, 4 iterations, plane scene seed 0 -- the
                   reference's `depths_upsampled` / `confidence_upsampled` at full resolution.
    e2e_cfg5.npz   BASELINE configs[4]: 1920x1056, 7 src views, 4 iterations, plane scene seed 3, for D=32
                   (checkpoint-compatible) and D=48 (SURVEY 8c recipe: hidden_init_head[0] re-created with 48 input
                   channels under seed 0; that weight is stored).  Outputs stored at every 4th pixel of every 4th
                   row (the full maps are 8 MB each).
The inputs are regenerated from the seed by the tests (itermvs_b200.synthetic.make_sample); a checksum of the
images guards against generator drift.  Loaded through oracle/reference_arm.py, i.e. the copy in baseline/_ref
whose sha256 equals /root/reference's (tools/install_ref.py).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from itermvs_b200.synthetic import make_sample  # noqa: E402
from oracle import reference_arm as RA  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
import install_ref  # noqa: E402


def run(width, height, n_src, num_sample, iteration, seed):
    m = RA.load_pipeline(iteration=iteration, num_sample=num_sample)
    s = make_sample(width, height, n_src=n_src, batch=1, seed=seed, scene="plane")
    with torch.no_grad():
        out = m(s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"])
    chk = np.array([float(s["imgs"]["level_0"].double().sum()), float(s["imgs"]["level_0"].double().abs().sum())])
    extra = None
    if num_sample != 32:
        extra = m.iter_mvs.update.hidden_init_head[0].weight.detach().numpy().copy()
    return out["depths_upsampled"].numpy(), out["confidence_upsampled"].numpy(), chk, extra


if __name__ == "__main__":
    assert install_ref.install(verbose=False) == "installed", "needs /root/reference"
    torch.set_num_threads(8)
    d, c, chk, _ = run(640, 512, 4, 32, 4, seed=0)
    np.savez_compressed(os.path.join(HERE, "e2e_cfg2.npz"), depths_upsampled=d, confidence_upsampled=c, img_checksum=chk,
                        meta=np.array([640, 512, 4, 32, 4, 0]))
    print("e2e_cfg2.npz depth", float(d.min()), float(d.max()), "mean conf", float(c.mean()))
    g = {}
    for D in (32, 48):
        d, c, chk, extra = run(1920, 1056, 7, D, 4, seed=3)
        g[f"depths_upsampled_d{D}"] = d[..., ::4, ::4].copy()
        g[f"confidence_upsampled_d{D}"] = c[..., ::4, ::4].copy()
        g["img_checksum"] = chk
        if extra is not None:
            g[f"hidden_init_head0_d{D}"] = extra
        print(f"e2e_cfg5 D={D} depth", float(d.min()), float(d.max()), "mean conf", float(c.mean()))
    g["meta"] = np.array([1920, 1056, 7, 4, 3, 4])
    np.savez_compressed(os.path.join(HERE, "e2e_cfg5.npz"), **g)
