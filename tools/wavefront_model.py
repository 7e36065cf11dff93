#!/usr/bin/env python
"""Access-pattern model of the plane-sweep kernels WITHOUT a GPU (round-2 design aid, not a measurement).

Runs warpcorr_init_kernel / warpcorr_iter_kernel at the benchmark configuration (640x512, 4 source views, D=32,
plane scene) through tests/cusim with its load tracer on (CUSIM_TRACE): every __ldg of a warp is grouped into the
warp-level request the hardware would issue, and per launch the tool reports

    requests   warp-level global load instructions            <-> ncu l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
    sectors    distinct 32-byte sectors per request, summed   <-> ncu l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
    wavefronts max(distinct 128-byte lines, ceil(register bytes / 128)) per request, summed -- a model of the L1 data
               stage (one line and at most 128 bytes of register data per wavefront)
                                                              <-> ncu l1tex__data_pipe_lsu_wavefronts (global part)
    delivered  bytes written to registers; / 128 = the floor of the wavefront count for this thread mapping
    unique_texel_bytes   distinct feature bytes per (item, view) footprint -- the floor under register-level reuse

against the counters ncu measured on the B200 for the same launches (profiles/ncu_warpcorr_r01_v7.txt), so that a
changed thread mapping can be compared on requests / sectors / wavefronts before it is timed on the device.

    python tools/wavefront_model.py [--width 640 --height 512 --views 4 --sms 148]
"""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cusim"))

# ncu, B200, v7 tree (profiles/ncu_warpcorr_r01_v7.txt): requests, sectors, total LSU wavefronts = pct * cycles * 148, shared wavefronts
NCU = {"init": {"requests": 1994240, "sectors": 8273104, "lsu_wavefronts": 0.52371729 * 88163 * 148, "shared_wavefronts": 1221635},
       "iter": {"requests": 1259520, "sectors": 8223896, "lsu_wavefronts": 0.57166255 * 76163 * 148, "shared_wavefronts": 1308553}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--sms", type=int, default=148)
    args = ap.parse_args()
    import cusim_build
    from itermvs_b200 import _lib
    from itermvs_b200.synthetic import make_sample, random_feature_pyramids

    trace = tempfile.NamedTemporaryFile("w", suffix=".jsonl", delete=False).name
    os.environ["CUSIM_SMS"] = str(args.sms)
    lib = C.CDLL(cusim_build.build())
    for name in ("imvs_compose_projections", "imvs_warpcorr_init", "imvs_warpcorr_iter", "imvs_last_error"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib._SIGNATURES[name]
    s = make_sample(args.width, args.height, n_src=args.views, batch=1, seed=0, scene="plane")
    ref, srcs = random_feature_pyramids(args.width, args.height, args.views, 1, 0)      # values are irrelevant to the pattern
    P = lambda t: t.data_ptr()
    feas, rts = [], []
    for l in (1, 2, 3):
        k = f"level{l}"
        feas.append(torch.stack([ref[k]] + list(srcs[k]), dim=1).permute(0, 1, 3, 4, 2).contiguous())
        proj = s["proj_matrices"][f"level_{l}"].float().contiguous()
        rt = torch.empty(1, args.views, 12)
        assert lib.imvs_compose_projections(P(proj), 1, args.views + 1, P(rt), None, None) == 0
        rts.append(rt)
    h2, w2 = args.height // 4, args.width // 4
    h3, w3 = h2 // 2, w2 // 2
    dmin, dmax = s["depth_min"].float(), s["depth_max"].float()
    # the plane's normalized depth at level-2 resolution: what the estimator's hypotheses are centred on after convergence
    from itermvs_b200.synthetic import plane_depth_map
    import numpy as np
    import torch.nn.functional as F
    depth2 = F.interpolate(torch.from_numpy(plane_depth_map(args.width, args.height).astype(np.float32))[None, None], scale_factor=0.25,
                           mode="nearest")
    nd = ((1.0 / depth2 - 1.0 / dmax) / (1.0 / dmin - 1.0 / dmax)).contiguous()
    vw = torch.rand(1, args.views, h2, w2)
    corr = torch.empty(1, args.views, 32, h3 * w3, 8)
    agg = torch.empty(1, 10, h2 * w2, 8)
    out = {}
    for name, call in (("init", lambda: lib.imvs_warpcorr_init(P(feas[2]), P(rts[2]), P(dmin), P(dmax), None, P(corr), 1, args.views + 1,
                                                               h3, w3, 32, None)),
                       ("iter", lambda: lib.imvs_warpcorr_iter(P(feas[0]), P(feas[1]), P(feas[2]), P(rts[0]), P(rts[1]), P(rts[2]), P(nd),
                                                               h2 * w2, 1, P(vw), P(dmin), P(dmax), None, None, None, P(agg), 1,
                                                               args.views + 1, h2, w2, None))):
        open(trace, "w").close()
        os.environ["CUSIM_TRACE"] = trace
        t = time.time()
        assert call() == 0, lib.imvs_last_error()
        os.environ["CUSIM_TRACE"] = ""
        rec = json.loads(open(trace).read().splitlines()[-1])
        tot = {k: sum(site[k] for site in rec["sites"]) for k in ("requests", "lanes", "bytes", "sectors", "lines", "wavefronts")}
        by_width = {}
        for site in rec["sites"]:
            wdt = site["bytes"] // max(site["lanes"], 1)
            d = by_width.setdefault(wdt, {"requests": 0, "sectors": 0, "lines": 0, "wavefronts": 0, "bytes": 0})
            for k in d:
                d[k] += site[k]
        out[name] = {"model": tot, "by_bytes_per_lane": by_width, "delivered_floor_wavefronts": tot["bytes"] / 128.0,
                     "sim_seconds": round(time.time() - t, 1), "grid": rec["grid"], "block": rec["block"]}
        if (args.width, args.height, args.views, args.sms) == (640, 512, 4, 148):
            n = NCU[name]
            out[name]["ncu"] = {"requests": n["requests"], "sectors": n["sectors"],
                                "global_wavefronts": n["lsu_wavefronts"] - n["shared_wavefronts"]}
            out[name]["model_over_ncu"] = {"requests": tot["requests"] / n["requests"], "sectors": tot["sectors"] / n["sectors"],
                                           "wavefronts": tot["wavefronts"] / (n["lsu_wavefronts"] - n["shared_wavefronts"])}
    os.unlink(trace)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
