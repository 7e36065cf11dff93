"""CUDA-graph replay of `Pipeline.forward` for a fixed input shape, and the stage profiler.

The stock path issues ~1 500 kernel launches and 52 host synchronisations per reference view
(SURVEY.md section 3.1); the CUDA path issues ~70 launches and never synchronises, so the whole forward
(FeatureNet + estimator) can be captured once and replayed with a single launch.

    g = GraphedPipeline(model, sample)        # sample: a representative batch (device tensors)
    out = g(imgs, proj_matrices, depth_min, depth_max)   # same dict as Pipeline.forward
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib

STAGE_NAMES = ["compose", "warpcorr_init", "pixel_view_weight", "aggregate_init", "corrnet", "hidden_init", "head",
               "warpcorr_iter", "gru", "upsample", "featurenet"]


class GraphedPipeline:
    def __init__(self, model, imgs: Dict[str, torch.Tensor], proj_matrices: Dict[str, torch.Tensor],
                 depth_min: torch.Tensor, depth_max: torch.Tensor, warmup: int = 3, workspace_slot: int = 0):
        assert model.test and not model.training, "graph replay is for the inference pipeline (test=True, eval())"
        self.model = model
        self.workspace_slot = workspace_slot      # graphs with different slots own disjoint workspaces -> may overlap
        dev = imgs["level_0"].device
        # static inputs: only what Pipeline.forward reads (net.py:78-109): imgs['level_0'], proj level_1..3
        # (uint8 images stay uint8: the first FeatureNet layer normalises them, a quarter of the H2D bytes per step)
        img0 = imgs["level_0"].detach().clone()
        self.s_img = (img0 if img0.dtype == torch.uint8 else img0.float()).contiguous()
        self.s_proj = {k: proj_matrices[k].detach().clone().float().contiguous() for k in ("level_1", "level_2", "level_3")}
        self.s_dmin = depth_min.detach().clone().float().contiguous()
        self.s_dmax = depth_max.detach().clone().float().contiguous()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.s_out = self._run()
        self.nan_flag = model._last_nan_flag
        # the captured launches hold RAW pointers into the model's workspaces and packed weights: keep those objects alive
        # for the life of the graph (a later forward at another shape / a repack replaces the model's cache entries,
        # which would otherwise free the memory under the graph), and remember the weight versions to refuse stale replays
        from . import estimator as _est
        self._held = (model.feature_net._ws.get(workspace_slot), model.iter_mvs._workspaces.get(workspace_slot),
                      model.feature_net._packed(dev), model.iter_mvs.packed(dev))
        watched = [t for m in (model.feature_net, model.iter_mvs) for t in list(m.parameters()) + list(m.buffers())]
        self._watch = (watched, [t._version for t in watched])      # ~10 us per replay to compare

    def _run(self):
        self.model.set_workspace_slot(self.workspace_slot)
        try:
            return self.model({"level_0": self.s_img}, self.s_proj, self.s_dmin, self.s_dmax)
        finally:
            self.model.set_workspace_slot(0)

    def load_inputs(self, imgs, proj_matrices, depth_min, depth_max, non_blocking=True):
        """Copies (H2D when the sources are pinned host tensors) into the graph's static inputs."""
        self.s_img.copy_(imgs["level_0"], non_blocking=non_blocking)
        for k in self.s_proj:
            self.s_proj[k].copy_(proj_matrices[k], non_blocking=non_blocking)
        self.s_dmin.copy_(depth_min, non_blocking=non_blocking)
        self.s_dmax.copy_(depth_max, non_blocking=non_blocking)

    def weights_current(self) -> bool:
        tensors, versions = self._watch
        return all(t._version == v for t, v in zip(tensors, versions))

    def replay(self):
        if not self.weights_current():
            raise RuntimeError("GraphedPipeline: the model's parameters / buffers changed after capture (optimizer step, "
                               "load_state_dict, BatchNorm update); the graph still reads the old packed weights -- re-capture")
        self.graph.replay()
        return self.s_out

    def check_nan(self) -> None:
        """The reference asserts 'nan in proj' on every warp (module.py:83,87); here one device flag per forward.
        Synchronises the flag's stream; raises if a composed projection contained NaN in the last replay."""
        if self.nan_flag is not None:
            self.nan_flag.raise_if_set()

    def __call__(self, imgs, proj_matrices, depth_min, depth_max):
        self.load_inputs(imgs, proj_matrices, depth_min, depth_max)
        return self.replay()


def profile_stages(fn, capacity: int = 4096):
    """Run fn() with the library's per-stage CUDA-event taps on; returns [(stage_name, ms), ...] in
    issue order.  Synchronises; not capturable -- for bench.py's roofline / breakdown pass only."""
    L = _lib.lib()
    _lib.check(L.imvs_profile_begin(capacity), "profile_begin")
    try:
        fn()
    finally:
        ms = (C.c_float * capacity)()
        tags = (C.c_int * capacity)()
        rc = L.imvs_profile_end(ms, tags, capacity)
    if rc > 0:
        _lib.check(rc, "profile_end")
    n = min(-rc, capacity)
    return [(STAGE_NAMES[tags[i]], float(ms[i])) for i in range(n)]


class StreamingPipeline:
    """Serving loop from pinned HOST buffers with `n_slots` graph slots, each with its own static inputs,
    workspace and compute stream: the H2D copy of sample k+1 (copy engine 1), the D2H copy of result k-1
    (copy engine 2) and the graph replays of up to `n_slots` consecutive samples overlap on the device.
    Reference views are independent units, so two of them in flight fill the SMs that the
    small quarter-resolution kernels of one forward (GRU, heads: 2-4 CTAs per SM) leave idle.

        sp = StreamingPipeline(model, imgs, proj, dmin, dmax)        # device sample, shapes only
        sp.submit(host_imgs, host_proj, host_dmin, host_dmax, out_depth_pinned, out_conf_pinned)   # per sample
        sp.drain()                                                   # all results are in their host buffers
    """

    def __init__(self, model, imgs, proj_matrices, depth_min, depth_max, n_slots: int = 2, concurrent: bool = True):
        dev = imgs["level_0"].device
        self.dev = dev
        self.n = n_slots
        self.slots = [GraphedPipeline(model, imgs, proj_matrices, depth_min, depth_max, workspace_slot=i if concurrent else 0)
                      for i in range(n_slots)]
        main = torch.cuda.current_stream(dev)
        self.comp = [torch.cuda.Stream(device=dev) if concurrent else main for _ in range(n_slots)]
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        self.ev_in = [torch.cuda.Event() for _ in range(n_slots)]
        self.ev_done = [torch.cuda.Event() for _ in range(n_slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(n_slots)]
        self.k = 0
        for e in self.ev_done + self.ev_out:
            e.record(main)

    def submit(self, imgs, proj_matrices, depth_min, depth_max, out_depth, out_conf):
        s = self.k % self.n
        slot, comp = self.slots[s], self.comp[s]
        if self.k < self.n:
            comp.wait_stream(torch.cuda.current_stream(self.dev))      # work the caller queued before the first submit
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(self.ev_done[s])          # slot's previous replay has consumed its inputs
            slot.load_inputs(imgs, proj_matrices, depth_min, depth_max)
            self.ev_in[s].record(self.h2d)
        comp.wait_event(self.ev_in[s])
        comp.wait_event(self.ev_out[s])                    # slot's previous outputs have left the device
        with torch.cuda.stream(comp):
            out = slot.replay()
            self.ev_done[s].record(comp)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.ev_done[s])
            out_depth.copy_(out["depths_upsampled"], non_blocking=True)
            out_conf.copy_(out["confidence_upsampled"], non_blocking=True)
            self.ev_out[s].record(self.d2h)
        self.k += 1

    def drain(self, check_nan: bool = False):
        main = torch.cuda.current_stream(self.dev)
        for e in self.ev_out:
            main.wait_event(e)
        if check_nan:            # the host is about to read the results anyway: surface the reference's 'nan in proj' assert
            for slot in self.slots:
                slot.check_nan()
