#!/usr/bin/env python
"""Benchmark of the IterMVS hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --steps 3 --warmup 1        # CPU oracle port, all host threads

Workload (BASELINE.json configs[1]): one reference view = 640x512 image + 4 source views, D=32
initial hypotheses, 4 GRU iterations, batch 1 per GPU, synthetic consistent-plane scene, DTU
checkpoint weights (tests/golden/dtu_weights.npz).  A step = one `Pipeline.forward` (FeatureNet +
the whole estimator), test mode.  Metric = reference views per second.

value : K steps with device-resident inputs through graph.StreamingPipeline: `--in-flight` (default 4)
        reference views on the device at once, each a CUDA-graph replay on its own stream / workspace
        (independent units); the inputs rotate over 8 resident sets (> L2); one CUDA-event
        pair around the K steps, barrier + synchronize on both sides, max over ranks.
        `single_stream` in the JSON line is the latency figure: one forward at a time, L2 flushed
        between steps, per-step events.
e2e   : the same K steps from pinned HOST buffers: H2D of images/cameras + forward + D2H of the two
        full-resolution outputs of every step inside the timed region.
roofline : fused warp+correlate iteration kernel, algorithmic bytes / its in-step
        duration from the library's CUDA-event stage taps, against MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

W_IMG, H_IMG, N_SRC, D_HYP, ITERS = 640, 512, 4, 32, 4
METRIC = "reference-views/sec at 640x512, 4 src, D=32, 4 iters"
CONFIG_NAME = "BASELINE configs[1]"
CONFIGS = {        # BASELINE.json configs by index: (W, H, source views, D, iterations)
    1: (640, 512, 4, 32, 4),         # configs[1]: the metric's configuration (default; the only driver line)
    4: (1920, 1056, 7, 48, 4),       # configs[4]: Tanks&Temples-shape memory / throughput stress (--config 4)
}


def set_config(idx):
    global W_IMG, H_IMG, N_SRC, D_HYP, ITERS, METRIC, CONFIG_NAME
    W_IMG, H_IMG, N_SRC, D_HYP, ITERS = CONFIGS[idx]
    METRIC = f"reference-views/sec at {W_IMG}x{H_IMG}, {N_SRC} src, D={D_HYP}, {ITERS} iters"
    CONFIG_NAME = f"BASELINE configs[{idx}]"


def hidden_init_weight(d):
    """D is hard-coded to 32 in the reference (itermvs.py:237); for D != 32 hidden_init_head[0] is re-created with D input
    channels under a fixed seed (SURVEY 8c recipe) -- the same tensor for this path and for the reference arm."""
    g = torch.Generator().manual_seed(0)
    return (torch.rand(64, d, 3, 3, generator=g) * 2 - 1) * (1.0 / (d * 9)) ** 0.5


def build_model(test=True):
    import itermvs_b200
    model = itermvs_b200.Pipeline(iteration=ITERS, test=test)
    sd = load_weights()
    if D_HYP != 32:
        model.iter_mvs.update.hidden_init_head[0] = torch.nn.Conv2d(D_HYP, 64, 3, stride=1, padding=1, bias=False)
        sd["iter_mvs.update.hidden_init_head.0.weight"] = hidden_init_weight(D_HYP)
    model.load_state_dict(sd, strict=True)
    return model
HBM_FALLBACK_GBS = 6650.0
NROT = 8          # device-resident input sets the timed loop rotates over (8 x 19.7 MB > the 126 MB L2)


def algorithmic_bytes(h, w, s, d):
    p2, p3 = (h // 4) * (w // 4), (h // 8) * (w // 8)
    init = 4 * p3 * (48 * (s + 1) + 8 * d)
    it = 4 * p2 * (108 * (s + 1) + (s + 1) + 80)
    return init, it


def load_weights():
    with np.load(os.path.join(ROOT, "tests", "golden", "dtu_weights.npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def pick_cpu_threads(run, budget_s=10.0):
    """The CPU port is many small torch ops: using every hardware thread of a 128-core host is far slower
    than a moderate count (measured: 23 s/forward at 128 threads vs 0.6 s at 16).  Probe 8/16/32 threads on
    one pass each within a time budget and keep the fastest."""
    ncpu = os.cpu_count() or 1
    cands = [c for c in (16, 8, 32) if c <= ncpu] or [ncpu]
    best, best_t, spent = None, None, 0.0
    torch.set_num_threads(cands[0])
    t0 = time.perf_counter()
    run()                                         # warm-up (allocator, first-touch)
    spent += time.perf_counter() - t0
    for c in cands:
        if best is not None and spent > budget_s:
            break
        torch.set_num_threads(c)
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        spent += dt
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def reference_runner(device="cpu"):
    """(run, kind, note): one forward of the reference arm on the benchmark workload.  kind = "reference" when the
    UNMODIFIED reference installed in baseline/_ref (tools/install_ref.py) is importable -- its own
    models.net.Pipeline(test=True) with the DTU checkpoint, reference models/net.py:78 -- else "port" (oracle/)."""
    from itermvs_b200.synthetic import make_sample
    from oracle import reference_arm as RA
    s = make_sample(W_IMG, H_IMG, n_src=N_SRC, batch=1, seed=0, scene="plane")
    if RA.available():
        m = RA.load_pipeline(iteration=ITERS, num_sample=D_HYP)
        if D_HYP != 32:
            with torch.no_grad():
                m.iter_mvs.update.hidden_init_head[0].weight.copy_(hidden_init_weight(D_HYP))
        m = m.to(device)
        imgs = {k: v.to(device) for k, v in s["imgs"].items()}
        proj = {k: v.to(device) for k, v in s["proj_matrices"].items()}
        dmin, dmax = s["depth_min"].to(device), s["depth_max"].to(device)

        def run():
            with torch.no_grad():
                return m(imgs, proj, dmin, dmax)
        return run, "reference", ("the unmodified reference (baseline/_ref: models/net.py Pipeline(test=True), DTU checkpoint) "
                                  "through its own public API")
    from oracle import itermvs_oracle as O
    weights = load_weights()
    if D_HYP != 32:
        weights["iter_mvs.update.hidden_init_head.0.weight"] = hidden_init_weight(D_HYP)
    weights = {k: v.to(device) for k, v in weights.items()}
    imgs = {k: v.to(device) for k, v in s["imgs"].items()}
    proj = {k: v.to(device) for k, v in s["proj_matrices"].items()}
    dmin, dmax = s["depth_min"].to(device), s["depth_max"].to(device)
    run = lambda: O.pipeline_forward(weights, imgs, proj, dmin, dmax, iteration=ITERS, num_sample=D_HYP)
    return run, "port", "baseline/_ref absent: CPU port of the reference algorithm (oracle/itermvs_oracle.py)"


def run_reference(args, rank):
    """Reference arm: the reference's own CPU path (baseline/_ref) on the box's host cores; rank 0 only."""
    if rank != 0:
        return
    run, kind, note = reference_runner("cpu")
    pick_cpu_threads(run)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    v = args.steps / dt
    cores = torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "refs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{W_IMG}x{H_IMG}, {N_SRC} src views, D={D_HYP}, {ITERS} iters, batch 1 ({CONFIG_NAME})",
                   "note": note + "; torch CPU fp32, thread count probed over 8/16/32"},
        "cpu_baseline": {"value": v, "unit": "refs/s", "cores": cores, "kind": kind, "host_cpus": os.cpu_count(),
                         "sample": f"{args.steps} full forward passes of the workload"},
        "e2e": {"value": v, "unit": "refs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_train(args, rank, world, local):
    """BASELINE configs[3]: data-parallel training on synthetic DTU-shape samples, one process per GPU, one reference view
    per GPU and step.  A step is train.py:194-215 -- zero_grad, Pipeline.train() forward (fused plane sweep + cuDNN
    convolution stacks), full_loss, backward (CUDA backward of the plane sweep), ONE all-reduce of the flat 1.37 MB
    gradient bucket over NCCL, clip_grad_norm_(2.0), Adam.  Prints one JSON line (rank 0); the per-phase split comes
    from CUDA events on the training stream."""
    import torch.distributed as dist
    import itermvs_b200
    from itermvs_b200.ddp import FlatBucketDDP
    from itermvs_b200.synthetic import make_sample, plane_depth_map
    import torch.nn.functional as F

    assert torch.cuda.is_available(), "bench.py needs a GPU (the CUDA path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    model = itermvs_b200.Pipeline(iteration=ITERS, test=False)
    model.load_state_dict(load_weights(), strict=True)
    model = model.to(dev).train()
    ddp = FlatBucketDDP(model, grad_dtype=torch.bfloat16 if args.train_bf16_wire else None)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-5, betas=(0.9, 0.999))
    s = make_sample(W_IMG, H_IMG, n_src=N_SRC, batch=1, seed=rank, scene="plane")
    d0 = torch.from_numpy(plane_depth_map(W_IMG, H_IMG).astype("float32"))[None, None]
    gt = {"level_0": d0.to(dev), "level_2": F.interpolate(d0, scale_factor=0.25, mode="nearest").to(dev)}
    mask = {k: torch.ones_like(v) for k, v in gt.items()}
    imgs = {k: v.to(dev) for k, v in s["imgs"].items()}
    proj = {k: v.to(dev) for k, v in s["proj_matrices"].items()}
    dmin, dmax = s["depth_min"].to(dev), s["depth_max"].to(dev)
    marks = []

    def step(timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timed else None
        ddp.zero_grad()
        if timed: ev[0].record()
        out = ddp(imgs, proj, dmin, dmax)
        loss = itermvs_b200.full_loss(out["depths"], out["depths_upsampled"], out["confidences"], gt, mask, dmin, dmax)
        if timed: ev[1].record()
        loss.backward()
        if timed: ev[2].record()
        ddp.reduce_gradients()
        if timed: ev[3].record()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 2.0)
        opt.step()
        if timed:
            ev[4].record()
            marks.append(ev)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    e0.record()
    for _ in range(args.steps):
        loss = step(True)
    e1.record()
    barrier()
    clock_info = clocks.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    phases = [sum(ev[i].elapsed_time(ev[i + 1]) for ev in marks) / len(marks) for i in range(4)]
    if rank == 0:
        print(json.dumps({
            "metric": "training reference-views/sec at 640x512, 4 src, D=32, 4 iters (forward + full_loss + backward + gradient all-reduce + Adam)",
            "value": world * args.steps / (ms / 1000.0), "unit": "refs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (gradients on the wire: %s)" % ("bf16" if args.train_bf16_wire else "f32"), "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: train step, 640x512, 4 src views, D=32, 4 iters, 1 reference view per GPU",
                       "parallelism": f"data parallel x{world}: one flat {ddp.gradient_bucket.numel() * 4} byte gradient bucket, one all-reduce per step",
                       "convolutions": "ATen/cuDNN under torch autograd", "plane_sweep": "fused sm_100a kernels, forward and backward"},
            "phase_ms": {"forward+loss": phases[0], "backward": phases[1], "allreduce": phases[2], "clip+adam": phases[3]},
            "loss": float(loss.item()), "clocks": clock_info}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=4,
                    help="reference views in flight per GPU (graph slots with their own workspace and compute stream); "
                         "1 = one forward on the device at a time")
    ap.add_argument("--passes", type=int, default=4, choices=[1, 3, 4],
                    help="tensor-core conv precision: 4 = 3-product FP16 split (fp32-grade, default), 3 = 3xTF32 split (fp32-grade), 1 = single-pass TF32")
    ap.add_argument("--breakdown", action="store_true", help="print the per-stage timing table to stderr")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer (default) = the headline metric; train = BASELINE configs[3]: training steps with the "
                         "flat-bucket gradient all-reduce (not a driver line; see run_train)")
    ap.add_argument("--train-bf16-wire", action="store_true", help="train mode: all-reduce the gradient bucket in bf16")
    ap.add_argument("--no-u8", action="store_true", help="skip the extra e2e leg with uint8 images")
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS),
                    help="BASELINE.json configs index: 1 = the metric's configuration (default), 4 = 1920x1056 / 7 src / D=48 stress")
    args = ap.parse_args()
    set_config(args.config)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.mode == "train":
        run_train(args, rank, world, local)
        return

    import torch.distributed as dist
    import itermvs_b200
    from itermvs_b200 import _lib
    from itermvs_b200.graph import GraphedPipeline, profile_stages
    from itermvs_b200.synthetic import make_sample

    assert torch.cuda.is_available(), "bench.py needs a GPU (the CUDA path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")        # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)
        # optional (IMVS_BENCH_PIN=1): one slice of the host cores per rank.  Measured at 8 GPUs (gpurun calls r2c34): e2e with fp32
        # images 8085 -> 7824 refs/s, with 8-bit images 8269 -> 8349, device-resident 8366 -> 8349: no clear gain, off by default --
        # the fp32 e2e leg is bound by 8 x 21 GB/s of pinned-memory reads from one NUMA node, not by thread placement
        if os.environ.get("IMVS_BENCH_PIN", "0") == "1" and hasattr(os, "sched_setaffinity"):
            try:
                cpus = sorted(os.sched_getaffinity(0))
                per = len(cpus) // world
                if per >= 1:
                    os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
            except OSError:
                pass
    _lib.set_conv_passes(args.passes)

    model = build_model(test=True).to(dev).eval()
    s = make_sample(W_IMG, H_IMG, n_src=N_SRC, batch=1, seed=rank, scene="plane")   # one reference view per GPU
    host = {"imgs": {"level_0": s["imgs"]["level_0"].pin_memory()},
            "proj": {k: s["proj_matrices"][k].float().pin_memory() for k in ("level_1", "level_2", "level_3")},
            "dmin": s["depth_min"].pin_memory(), "dmax": s["depth_max"].pin_memory()}
    d_imgs = {"level_0": host["imgs"]["level_0"].to(dev)}
    d_proj = {k: v.to(dev) for k, v in host["proj"].items()}
    d_dmin, d_dmax = host["dmin"].to(dev), host["dmax"].to(dev)

    def eager():
        with torch.no_grad():
            return model(d_imgs, d_proj, d_dmin, d_dmax)

    eager()
    torch.cuda.synchronize()
    l0 = _lib.launches_total()
    eager()
    torch.cuda.synchronize()
    launches_per_step = _lib.launches_total() - l0
    model._last_nan_flag.raise_if_set()

    graphed = None if args.no_graph else GraphedPipeline(model, d_imgs, d_proj, d_dmin, d_dmax)
    step = eager if graphed is None else graphed.replay
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    # ---- stage breakdown + in-step duration of the fused warp+correlate kernels (eager, taps on)
    stage_ms = {}
    for _ in range(3):
        flush.zero_()
        recs = profile_stages(eager)
        acc = {}
        for name, ms in recs:
            acc.setdefault(name, []).append(ms)
        for name, lst in acc.items():
            stage_ms.setdefault(name, []).append(lst)
    iter_ms = [m for rep in stage_ms.get("warpcorr_iter", []) for m in rep]
    init_ms = [m for rep in stage_ms.get("warpcorr_init", []) for m in rep]

    # ---- single-stream figure: one forward on the device at a time, L2 flushed between steps (latency-oriented)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        step()
        b.record()
    barrier()
    serial_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- timed region (the metric): K reference views, device-resident inputs, `in_flight` of them on the device
    #      at once (independent units).  The inputs rotate over NROT resident sets (> L2 in total).
    sp = None
    if graphed is not None:
        from itermvs_b200.graph import StreamingPipeline
        nfl = max(1, args.in_flight)
        sp = StreamingPipeline(model, d_imgs, d_proj, d_dmin, d_dmax, n_slots=nfl, concurrent=nfl > 1)
        rot = [({"level_0": d_imgs["level_0"].clone()}, {k: v.clone() for k, v in d_proj.items()}, d_dmin.clone(), d_dmax.clone())
               for _ in range(NROT)]
        d_outs = [(torch.empty(1, 1, H_IMG, W_IMG, device=dev), torch.empty(1, 1, H_IMG, W_IMG, device=dev)) for _ in range(nfl)]
        for k in range(2 * nfl):
            sp.submit(*rot[k % NROT], *d_outs[k % nfl])
        sp.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        e0.record()
        for k in range(args.steps):
            sp.submit(*rot[k % NROT], *d_outs[k % nfl])
        sp.drain()
        e1.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
        total_ms = e0.elapsed_time(e1)
        ref_d = model(*rot[(args.steps - 1) % NROT])["depths_upsampled"]
        assert torch.equal(d_outs[(args.steps - 1) % nfl][0], ref_d), "in-flight replay differs from the plain forward"
        # spread of the figure: the same K-step region twice more (reported next to `value`, which stays the first region)
        repeat_ms = []
        for _ in range(2):
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for k in range(args.steps):
                sp.submit(*rot[k % NROT], *d_outs[k % nfl])
            sp.drain()
            r1.record()
            barrier()
            repeat_ms.append(r0.elapsed_time(r1))
    else:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        t_wall = time.perf_counter()
        for a, b in ev:
            flush.zero_()
            a.record()
            step()
            b.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
        total_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- end to end from pinned host buffers
    out_host = {"d": torch.empty(1, 1, H_IMG, W_IMG).pin_memory(), "c": torch.empty(1, 1, H_IMG, W_IMG).pin_memory()}
    h2d = host["imgs"]["level_0"].numel() * 4 + sum(v.numel() * 4 for v in host["proj"].values()) + 8
    d2h = 2 * H_IMG * W_IMG * 4

    def e2e_step():
        if graphed is not None:
            graphed.load_inputs(host["imgs"], host["proj"], host["dmin"], host["dmax"])
            out = graphed.replay()
        else:
            with torch.no_grad():
                out = model({"level_0": host["imgs"]["level_0"].to(dev, non_blocking=True)},
                            {k: v.to(dev, non_blocking=True) for k, v in host["proj"].items()},
                            host["dmin"].to(dev, non_blocking=True), host["dmax"].to(dev, non_blocking=True))
        out_host["d"].copy_(out["depths_upsampled"], non_blocking=True)
        out_host["c"].copy_(out["confidence_upsampled"], non_blocking=True)

    for _ in range(3):
        e2e_step()
    e2e_mode = "serial (H2D -> forward -> D2H per step)"
    if sp is not None:
        # the call a serving user makes: streaming from pinned host buffers; every step's H2D of its inputs and D2H of
        # its two result maps are inside the timed region, overlapped with other steps' compute on the copy engines
        ns = sp.n
        outs = [(torch.empty(1, 1, H_IMG, W_IMG).pin_memory(), torch.empty(1, 1, H_IMG, W_IMG).pin_memory()) for _ in range(ns)]
        for k in range(2 * ns):
            sp.submit(host["imgs"], host["proj"], host["dmin"], host["dmax"], *outs[k % ns])
        sp.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            sp.submit(host["imgs"], host["proj"], host["dmin"], host["dmax"], *outs[k % ns])
        sp.drain()
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        e2e_mode = (f"streaming from pinned host buffers through graph.StreamingPipeline: {ns} reference view(s) in flight "
                    "(own static buffers, workspace and compute stream per slot), H2D / forward / D2H of different steps overlap; "
                    "all copies inside the timed region")
        ref_d = model(d_imgs, d_proj, d_dmin, d_dmax)["depths_upsampled"]
        assert torch.allclose(outs[(args.steps - 1) % ns][0].to(dev), ref_d, rtol=0, atol=0), "streaming result differs"
    else:
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for a, b in ev2:
            flush.zero_()
            a.record()
            e2e_step()
            b.record()
        barrier()
        e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)
    # ---- the same serving loop fed with the raw 8-bit images (f-4: the first FeatureNet layer normalises 2 x / 255. - 1 as the
    #      reference's loaders do): a quarter of the H2D bytes per step.  Extra line, the contract's `e2e` stays fp32 images.
    e2e_u8 = None
    if sp is not None and not args.no_u8:
        u8 = ((host["imgs"]["level_0"] + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).pin_memory()
        sp8 = StreamingPipeline(model, {"level_0": u8.to(dev)}, d_proj, d_dmin, d_dmax, n_slots=sp.n, concurrent=sp.n > 1)
        h8 = {"level_0": u8}
        for k in range(2 * sp8.n):
            sp8.submit(h8, host["proj"], host["dmin"], host["dmax"], *outs[k % sp8.n])
        sp8.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            sp8.submit(h8, host["proj"], host["dmin"], host["dmax"], *outs[k % sp8.n])
        sp8.drain()
        e1.record()
        barrier()
        e2e_u8 = {"ms": e0.elapsed_time(e1), "h2d_bytes_per_step": u8.numel() + sum(v.numel() * 4 for v in host["proj"].values()) + 8}
        del sp8
    clk = clocks.stop()

    # replicas: every rank processed `steps` reference views; whole-job rate = all units / slowest rank
    from itermvs_b200 import replicas
    value = replicas.aggregate_throughput(args.steps, total_ms, device=dev)
    e2e_value = replicas.aggregate_throughput(args.steps, e2e_ms, device=dev)
    if e2e_u8 is not None:
        e2e_u8["value"] = replicas.aggregate_throughput(args.steps, e2e_u8.pop("ms"), device=dev)
    total_ms = replicas.max_over_ranks([total_ms], device=dev)[0]
    value_repeats = None
    if sp is not None:
        value_repeats = [replicas.aggregate_throughput(args.steps, m, device=dev) for m in repeat_ms]

    mem_ours = torch.cuda.max_memory_allocated(dev)
    torch.cuda.reset_peak_memory_stats(dev)
    mem_held = torch.cuda.memory_allocated(dev)          # this path's live tensors stay allocated while the reference runs
    # ---- CPU baseline: the reference itself (baseline/_ref) on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        run, kind, note = reference_runner("cpu")
        nthr = pick_cpu_threads(run)
        ts, t_begin = [], time.perf_counter()
        for _ in range(5):
            t0 = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0)
            if time.perf_counter() - t_begin > 20.0:      # bounded sample
                break
        cpu = {"value": 1.0 / statistics.median(ts), "unit": "refs/s", "cores": nthr, "kind": kind,
               "host_cpus": os.cpu_count(),
               "sample": f"{len(ts)} full forward passes of the workload (median) after a thread-count probe; " + note}

    # ---- the reference's stock PyTorch/cuDNN path on THIS GPU (BASELINE.md section 2, B-GPU): the same unmodified
    #      module .cuda(), cudnn.benchmark=True as eval.py:21, torch's default TF32 settings, inputs resident,
    #      CUDA events around every forward, >= 50 repetitions
    gpu_stock = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.backends.cudnn.benchmark = True
        run_g, kind_g, note_g = reference_runner(dev)
        for _ in range(5):
            run_g()
        torch.cuda.synchronize()
        reps = 50
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        t0 = time.perf_counter()
        for a_, b_ in evs:
            a_.record()
            run_g()
            b_.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = sorted(a_.elapsed_time(b_) for a_, b_ in evs)
        # parity of this path against the reference itself on the same GPU and inputs (rank 0: seed 0 for both)
        parity = None
        if kind_g == "reference":
            out_ref = run_g()
            out_new = eager()
            rel = ((out_new["depths_upsampled"] - out_ref["depths_upsampled"]).abs() / out_ref["depths_upsampled"].abs())
            dc = (out_new["confidence_upsampled"] - out_ref["confidence_upsampled"]).abs()
            parity = {"depth_rel_max": float(rel.max()), "depth_rel_median": float(rel.median()),
                      "depth_frac_gt_1e-3": float((rel > 1e-3).float().mean()), "confidence_abs_max": float(dc.max()),
                      "note": "reference runs cuDNN with torch's default TF32 convolutions here; this path is fp32-grade"}
        gpu_stock = {"value": reps / (sum(ms) / 1000.0), "unit": "refs/s", "kind": kind_g, "reps": reps,
                     "ms_median": statistics.median(ms), "ms_min": ms[0], "wall_refs_per_s": reps / wall, "parity": parity,
                     "max_allocated_MB": (torch.cuda.max_memory_allocated(dev) - mem_held) / 1e6,
                     "what": note_g + " on this GPU: .cuda(), cudnn.benchmark=True (eval.py:21), torch default TF32 flags, "
                             "device-resident inputs, CUDA events around each forward"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    init_b, iter_b = algorithmic_bytes(H_IMG, W_IMG, N_SRC, D_HYP)
    peak, peak_src = hbm_peak()
    mean_iter_ms = statistics.mean(iter_ms) if iter_ms else float("nan")
    achieved = iter_b / (mean_iter_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp) and args.config == 1:          # the capture was taken at configs[1]
        try:
            traffic = json.load(open(tp)).get("warpcorr_iter_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    breakdown = {k: round(statistics.mean(sum(rep) for rep in v), 4) for k, v in stage_ms.items()}
    if args.breakdown:
        print("stage breakdown (ms per forward, eager + event taps):", json.dumps(breakdown), file=sys.stderr)
    fwd_launches = int(launches_per_step)
    line = {
        "metric": METRIC, "value": value, "unit": "refs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {4: "f32 (fp32-grade tensor-core convolutions: 3-product FP16 hi/lo split, fp32 accumulate; sampling/softmax/regression fp32)",
                  3: "f32 (3xTF32 tensor-core convolutions, fp32 accumulate; sampling/softmax/regression fp32)",
                  1: "tf32 (single-pass TF32 tensor-core convolutions, fp32 accumulate; rest fp32)"}[args.passes], "data": "synthetic",
        "config": {"workload": f"{W_IMG}x{H_IMG}, {N_SRC} src views, D={D_HYP}, {ITERS} iters, batch 1 per GPU ({CONFIG_NAME})",
                   "step": "Pipeline.forward test mode: FeatureNet + estimator, all in hand-written sm_100a kernels (no cuDNN/cuBLAS)",
                   "launch": "eager" if graphed is None else "cuda-graph replay",
                   "in_flight": (sp.n if sp is not None else 1),
                   "l2": (f"device-resident inputs rotate over {NROT} sets ({NROT * h2d / 1e6:.0f} MB > 126 MB L2), no flush inside the timed region"
                          if sp is not None else "256 MiB memset between steps, outside the per-step CUDA-event windows"),
                   "weights": "DTU checkpoint", "parallelism": f"replicas x{world} (one reference view per GPU, no collectives)"},
        "e2e": {"value": e2e_value, "unit": "refs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "mode": e2e_mode},
        "single_stream": {"value": args.steps / (serial_ms * 1e-3), "unit": "refs/s", "ms_per_step": serial_ms / args.steps,
                          "what": "one forward on the device at a time (graph replay), L2 flushed by a 256 MiB memset between steps, "
                                  "per-step CUDA events: the latency-oriented figure"},
        "gpu_launches": fwd_launches * args.steps,
        "gpu_launches_per_step": fwd_launches,
        "roofline": {"kernel": "warpcorr_iter_kernel (fused warp+sample+group-corr+view-weighted aggregation)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_source": ("constant from one `ncu --set full` capture of this kernel at this configuration "
                                        "(profiles/ncu_traffic.json, profiles/ncu_warpcorr_r02.txt), not measured in this run") if traffic else None,
                     # bytes the gathers deliver to registers through the L1 data pipe: P2 * S * sum_l(R_l * C_l) * 4 taps * 4 B with
                     # (R, C) = (4, 16), (4, 32), (2, 48) -- what actually bounds this kernel
                     "l1_tap_bytes_per_launch": (H_IMG // 4) * (W_IMG // 4) * N_SRC * 288 * 16,
                     "l1_tap_bytes_over_algorithmic": (H_IMG // 4) * (W_IMG // 4) * N_SRC * 288 * 16 / iter_b,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": iter_b,
                     "avg_launch_ms": mean_iter_ms, "launches_timed": len(iter_ms),
                     "init_kernel": {"algorithmic_bytes": init_b, "avg_launch_ms": statistics.mean(init_ms) if init_ms else None}},
        "stage_ms": breakdown,
        "memory": {"max_allocated_MB": mem_ours / 1e6,
                   "what": "torch.cuda.max_memory_allocated up to the end of this path's timed regions: weights, the caller-owned "
                           "workspaces of all in-flight slots (FeatureNet + estimator), the rotating input sets, outputs"},
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"], "samples": clk["samples"]},
        "wall_s_timed_region": t_wall,
        "value_repeats": value_repeats,          # the same K-step timed region run twice more (spread of `value`)
    }
    if e2e_u8 is not None:
        line["e2e_uint8_images"] = {"value": e2e_u8["value"], "unit": "refs/s", "h2d_bytes_per_step": e2e_u8["h2d_bytes_per_step"],
                                    "d2h_bytes_per_step": d2h,
                                    "what": "the e2e loop fed with raw 8-bit images from pinned host memory; 2 x / 255. - 1 (the loaders' normalisation, "
                                            "dtu_yao_eval.py:63-64) happens in the first FeatureNet kernel"}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if gpu_stock is not None:
        line["gpu_stock_ref" if gpu_stock["kind"] == "reference" else "gpu_stock_port"] = gpu_stock
        line["speedup_vs_gpu_stock"] = {"e2e": e2e_value / gpu_stock["value"], "single_stream": line["single_stream"]["value"] / gpu_stock["value"],
                                        "what": "this path / the reference's stock PyTorch-cuDNN path on the same GPU (target >= 5x)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
