#!/usr/bin/env python
"""A/B sweep of the IMVS_TUNE_* tile-shape switches: one bench.py run per setting, prints refs/s and stage times."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWEEPS = [("", 0)] + [(k, v) for k, vals in (("PDL", (0,)), ("K8", (0,)), ("CONV0", (0,)), ("GRU", (0,)), ("HEAD", (0,)), ("FC", (0,)),
                                              ("FNET3", (0,))) for v in vals]


def setenv(env, k, v):
    env["IMVS_PDL" if k == "PDL" else f"IMVS_TUNE_{k}"] = str(v)      # PDL is a global switch, the rest tile-shape switches


if len(sys.argv) > 1:      # explicit settings: NAME=V,NAME=V ...
    SWEEPS = [(a, None) for a in sys.argv[1:]]
for name, val in SWEEPS:
    env = dict(os.environ)
    if val is None:
        for kv in name.split(","):
            k, v = kv.split("=")
            setenv(env, k, v)
        label = name
    else:
        if name:
            setenv(env, name, val)
        label = f"{name}={val}" if name else "base"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-cpu-baseline"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        st = d["stage_ms"]
        print(f"{label:24s} {d['value']:7.1f} refs/s  fnet {st['featurenet']:.4f} gru {st['gru']:.4f} head {st['head']:.4f} "
              f"corrnet {st['corrnet']:.4f} hinit {st['hidden_init']:.4f} pvw {st['pixel_view_weight']:.4f} ups {st['upsample']:.4f} "
              f"single {d['single_stream']['value']:.1f} e2e {d['e2e']['value']:.1f}", flush=True)
    except Exception as e:
        print(label, "FAILED", e, r.stderr[-400:], flush=True)
