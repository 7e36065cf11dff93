#!/usr/bin/env python
"""Autograd memory of one training forward: bytes of tensors saved for the backward pass, reference vs itermvs_b200.

Tensor sizes do not depend on the device, so this runs where the reference runs -- the build container (needs
/root/reference): the reference's Pipeline.train() forward on CPU, and itermvs_b200's training path with its fused
operators executed through tests/cusim.  Saved tensors are counted once per storage (torch.autograd.graph.saved_tensors_hooks).

    python tools/saved_for_backward.py [--width 640 --height 512 --views 4 --iters 4]
"""
import argparse
import ctypes as C
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cusim"))
warnings.filterwarnings("ignore")


class Counter:
    def __init__(self):
        self.seen, self.bytes, self.largest = set(), 0, []

    def pack(self, t):
        key = (t.untyped_storage().data_ptr(), t.untyped_storage().nbytes())
        if key not in self.seen and t.untyped_storage().nbytes() > 0:
            self.seen.add(key)
            self.bytes += t.untyped_storage().nbytes()
            self.largest.append((t.untyped_storage().nbytes(), tuple(t.shape)))
        return t


def measure(fn):
    c = Counter()
    with torch.autograd.graph.saved_tensors_hooks(c.pack, lambda t: t):
        out = fn()
    c.largest.sort(reverse=True)
    return c, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    import numpy as np
    import cusim_build
    import itermvs_b200
    from itermvs_b200 import _lib, training
    from itermvs_b200.synthetic import make_sample
    with np.load(os.path.join(ROOT, "tests", "golden", "dtu_weights.npz")) as z:
        weights = {k: torch.from_numpy(z[k]) for k in z.files}
    s = make_sample(args.width, args.height, n_src=args.views, batch=1, seed=0, scene="plane")

    sys.path.insert(0, "/root/reference")
    from models.net import Pipeline as RefPipeline
    ref = RefPipeline(iteration=args.iters, test=False)
    ref.load_state_dict(weights, strict=True)
    ref.train()
    c_ref, _ = measure(lambda: ref(s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"]))

    lib = C.CDLL(cusim_build.build())
    for name in ("imvs_compose_projections", "imvs_warpcorr_init", "imvs_warpcorr_iter", "imvs_last_error"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib._SIGNATURES[name]
    training._L, training._st, training._chk = (lambda: lib), (lambda: None), (lambda t, n: t.float().contiguous())
    os.environ.setdefault("CUSIM_SMS", "16")
    m = itermvs_b200.Pipeline(iteration=args.iters, test=False)
    m.load_state_dict(weights, strict=True)
    m.train()
    c_new, _ = measure(lambda: training.pipeline_train_forward(m, s["imgs"], s["proj_matrices"], s["depth_min"], s["depth_max"]))
    print(json.dumps({
        "config": {"width": args.width, "height": args.height, "src_views": args.views, "iterations": args.iters, "batch": 1},
        "reference_saved_MB": round(c_ref.bytes / 1e6, 1), "itermvs_b200_saved_MB": round(c_new.bytes / 1e6, 1),
        "ratio": round(c_ref.bytes / max(c_new.bytes, 1), 2),
        "reference_largest": [{"MB": round(b / 1e6, 1), "shape": list(sh)} for b, sh in c_ref.largest[:6]],
        "itermvs_b200_largest": [{"MB": round(b / 1e6, 1), "shape": list(sh)} for b, sh in c_new.largest[:6]]}, indent=1))


if __name__ == "__main__":
    main()
