#!/usr/bin/env python
"""Micro-benchmark of the two fused plane-sweep kernels through the C ABI (imvs_warpcorr_init / imvs_warpcorr_iter)
at a BASELINE configuration, outside the pipeline: CUDA events around single launches, L2 flushed (cold) or not (warm,
the pyramids were just written by FeatureNet in the real pipeline and 35 MB of them sit in the 126 MB L2).

    python tools/bench_planesweep.py [--config 2|5] [--reps 30] [--tag name] [--save ref.pt | --check ref.pt]

Inputs: seeded N(0,1) feature pyramids (values do not influence timing), the synthetic cameras of
itermvs_b200.synthetic, the normalized depth of the synthetic plane (+ optional noise) as the current estimate -- the
access pattern of the real iterations on the consistent scene -- and uniform view weights.  Prints one JSON line.
Kernel variants are selected with IMVS_WC_* environment variables read by the library (see csrc/warpcorr.cu).
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--noise", type=float, default=0.0, help="uniform noise amplitude on the normalized depth")
    ap.add_argument("--tag", default="")
    ap.add_argument("--save", default=None)
    ap.add_argument("--check", default=None)
    ap.add_argument("--pad3", action="store_true", help="level 3 from the padded copy (imvs_pad_level3 + the *_padded entry points)")
    args = ap.parse_args()
    from itermvs_b200 import _lib, ops
    from itermvs_b200.synthetic import make_sample, plane_depth_map, DEPTH_MIN, DEPTH_MAX
    W, H, S, D = {2: (640, 512, 4, 32), 5: (1920, 1056, 7, 48), 1: (160, 128, 2, 8)}[args.config]
    dev = torch.device("cuda:0")
    L = _lib.lib()
    g = torch.Generator().manual_seed(0)
    V = S + 1
    fea = [torch.randn(1, V, H // s, W // s, c, generator=g).to(dev) for s, c in ((2, 16), (4, 32), (8, 48))]
    smp = make_sample(W, H, n_src=S, batch=1, seed=0, scene="noise")
    rts = [ops.compose_projections(smp["proj_matrices"][f"level_{l}"].float().to(dev)) for l in (1, 2, 3)]
    depth = torch.from_numpy(plane_depth_map(W, H).astype(np.float32))[::4, ::4].contiguous()
    inv_min, inv_max = 1.0 / DEPTH_MIN, 1.0 / DEPTH_MAX
    nd = ((1.0 / depth - inv_max) / (inv_min - inv_max)).clamp(0, 1)
    if args.noise > 0:
        nd = (nd + args.noise * (2 * torch.rand(nd.shape, generator=g) - 1)).clamp(0, 1)
    H2, W2, H3, W3 = H // 4, W // 4, H // 8, W // 8
    nd = nd.reshape(1, H2 * W2).contiguous().to(dev)
    vw = torch.rand(1, S, H2 * W2, generator=g).to(dev)
    dmin = torch.full((1,), DEPTH_MIN, device=dev)
    dmax = torch.full((1,), DEPTH_MAX, device=dev)
    agg = torch.zeros(1, 10, H2 * W2, 8, device=dev)
    corr = torch.zeros(1, S, D, H3 * W3, 8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    f3 = fea[2]
    iter_fn, init_fn = L.imvs_warpcorr_iter, L.imvs_warpcorr_init
    if args.pad3:
        f3 = torch.empty(1, V, H // 8, W // 8, 64, device=dev)
        iter_fn, init_fn = L.imvs_warpcorr_iter_padded, L.imvs_warpcorr_init_padded

    def run_pad():
        _lib.check(L.imvs_pad_level3(fea[2].data_ptr(), f3.data_ptr(), 1, V, H // 8, W // 8, st), "pad_level3")

    def run_iter():
        _lib.check(iter_fn(fea[0].data_ptr(), fea[1].data_ptr(), f3.data_ptr(), rts[0].data_ptr(), rts[1].data_ptr(),
                                        rts[2].data_ptr(), nd.data_ptr(), H2 * W2, 1, vw.data_ptr(), dmin.data_ptr(), dmax.data_ptr(),
                                        None, None, None, agg.data_ptr(), 1, V, H2, W2, st), "warpcorr_iter")

    def run_init():
        _lib.check(init_fn(f3.data_ptr(), rts[2].data_ptr(), dmin.data_ptr(), dmax.data_ptr(), None,
                                        corr.data_ptr(), 1, V, H3, W3, D, st), "warpcorr_init")

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, cold):
        ts = []
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        for _ in range(args.reps):
            if cold:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return {"median_us": ts[len(ts) // 2], "min_us": ts[0], "p90_us": ts[int(0.9 * len(ts))]}

    out = {"tag": args.tag, "config": args.config, "noise": args.noise,
           "env": {k: v for k, v in os.environ.items() if k.startswith("IMVS_")}}
    if args.pad3:
        run_pad()
        out["pad_warm"] = timed(run_pad, False)
    out["iter_warm"] = timed(run_iter, False)
    out["iter_cold"] = timed(run_iter, True)
    out["init_warm"] = timed(run_init, False)
    out["init_cold"] = timed(run_init, True)
    p2, p3 = H2 * W2, H3 * W3
    it_b, in_b = 4 * p2 * (108 * (S + 1) + (S + 1) + 80), 4 * p3 * (48 * (S + 1) + 8 * D)
    out["iter_GBs_warm"] = it_b / out["iter_warm"]["median_us"] / 1e3
    out["init_GBs_warm"] = in_b / out["init_warm"]["median_us"] / 1e3
    res = {"agg": agg.cpu(), "corr": corr.cpu()}
    if args.save:
        torch.save(res, args.save)
    if args.check and os.path.exists(args.check):
        ref = torch.load(args.check)
        out["max_abs_diff_vs_ref"] = {k: float((res[k] - ref[k]).abs().max()) for k in res}
        out["ref_abs_max"] = {k: float(ref[k].abs().max()) for k in res}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
