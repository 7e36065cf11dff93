"""Loader for the UNMODIFIED reference (`baseline/_ref`, installed by tools/install_ref.py).  TEST / BENCH
INFRASTRUCTURE ONLY: imported by bench.py (`--impl reference`, `gpu_stock_ref`) and by tests/; never by the
product package (tests/test_abi_host.py::test_product_never_imports_oracle covers oracle/ as a whole).

`load_pipeline()` returns the reference's own `models.net.Pipeline` (reference models/net.py:68-128) with the shipped
DTU checkpoint loaded exactly the way eval.py:122-125 does it (state_dict saved from the DataParallel wrapper, so the
`module.` prefix is stripped), in `eval()` mode.  For D != 32 the recipe of SURVEY.md 8c applies (D is hard-coded in
the reference; `hidden_init_head[0]` is re-created with D input channels under a fixed seed).
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
CKPT = os.path.join(REF_DIR, "checkpoints", "dtu", "model_000015.ckpt")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "models", "net.py")) and os.path.exists(CKPT)


def why_unavailable() -> str:
    return ("baseline/_ref is not installed (run `python tools/install_ref.py` in the build container; "
            "/root/reference does not exist on the GPU box)")


def _import_models():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    warnings.filterwarnings("ignore", message=".*torch.meshgrid.*")
    import models.net as net          # the reference package, unmodified
    assert os.path.realpath(net.__file__).startswith(os.path.realpath(REF_DIR)), net.__file__
    return net


def load_pipeline(iteration: int = 4, num_sample: int = 32, test: bool = True, seed: int = 0):
    import torch
    import torch.nn as nn
    net = _import_models()
    m = net.Pipeline(iteration=iteration, test=test)
    sd = torch.load(CKPT, map_location="cpu")["model"]
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    if num_sample != 32:
        m.iter_mvs.num_sample = num_sample
        m.iter_mvs.depth_initialization.num_sample = num_sample
        torch.manual_seed(seed)
        m.iter_mvs.update.hidden_init_head[0] = nn.Conv2d(num_sample, 64, 3, stride=1, padding=1, bias=False)
        key = "iter_mvs.update.hidden_init_head.0.weight"
        sd = {k: v for k, v in sd.items() if k != key}
        m.load_state_dict(sd, strict=False)
    else:
        m.load_state_dict(sd, strict=True)
    return m.eval()
