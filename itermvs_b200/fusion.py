"""Depth-map filtering for fusion on the GPU: host-side mirror of reference eval.py:154-265.

    check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                geo_pixel_thres, geo_depth_thres)  ->  mask, depth_reprojected, x2d_src, y2d_src
        same name, argument order and return tuple as eval.py:199; numpy arrays in -> numpy arrays out (what
        filter_depth passes), CUDA tensors in -> CUDA tensors out.
    filter_depth_view(depth_ref, confidence, intrinsics_ref, extrinsics_ref, depth_srcs, intrinsics_srcs,
                      extrinsics_srcs, geo_pixel_thres, geo_depth_thres, photo_thres, geo_mask_thres=3)
        the body of filter_depth's loop over reference views (eval.py:238-265) in S + 1 launches:
        -> depth_est_averaged (float64), photo_mask, geo_mask, final_mask

The camera algebra (two 3x3 / 4x4 inverses and two 4x4 products per pair) is done on the host with numpy in
float32, exactly the expressions of eval.py:162-190, so the kernels see bit-identical matrices; everything per
pixel runs in imvs_check_geometric_consistency / imvs_filter_depth_view.  No CPU fallback.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib
from . import ops


def pair_cameras(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src) -> np.ndarray:
    """The six matrices of one (reference, source) pair as the 68-float block of include/itermvs_b200.h."""
    kr, er, ks, es = (np.asarray(m) for m in (intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src))
    blocks = [np.linalg.inv(kr), np.matmul(es, np.linalg.inv(er)), ks, np.linalg.inv(ks), np.matmul(er, np.linalg.inv(es)), kr]
    return np.concatenate([np.asarray(b, dtype=np.float32).reshape(-1) for b in blocks])


def _dev_map(a, device) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) if isinstance(a, np.ndarray) else a
    t = t.to(device=device, dtype=torch.float32)
    if t.dim() == 3 and t.shape[-1] == 1:          # read_pfm returns [H,W,1]
        t = t[..., 0]
    return ops._chk(t, "depth map")


def _device_of(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise RuntimeError("itermvs_b200.fusion: a CUDA device is required (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                geo_pixel_thres, geo_depth_thres):
    """Drop-in for eval.py:199."""
    as_numpy = isinstance(depth_ref, np.ndarray)
    dev = _device_of(depth_ref, depth_src)
    dr, ds = _dev_map(depth_ref, dev), _dev_map(depth_src, dev)
    h, w = dr.shape
    if ds.shape != dr.shape:
        raise ValueError("reference and source depth maps must have the same shape (eval.py remaps onto the reference grid)")
    cams = np.ascontiguousarray(pair_cameras(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src))
    mask = torch.empty(h, w, dtype=torch.uint8, device=dev)
    rep, xs, ys = (torch.empty(h, w, device=dev) for _ in range(3))
    _lib.check(_lib.lib().imvs_check_geometric_consistency(dr.data_ptr(), ds.data_ptr(), cams.ctypes.data, float(geo_pixel_thres),
                                                           float(geo_depth_thres), mask.data_ptr(), rep.data_ptr(), xs.data_ptr(),
                                                           ys.data_ptr(), None, None, h, w, ops._stream()),
               "check_geometric_consistency")
    mask = mask.bool()
    if as_numpy:
        return mask.cpu().numpy(), rep.cpu().numpy(), xs.cpu().numpy(), ys.cpu().numpy()
    return mask, rep, xs, ys


def filter_depth_view(depth_ref, confidence, intrinsics_ref, extrinsics_ref, depth_srcs: Sequence, intrinsics_srcs: Sequence,
                      extrinsics_srcs: Sequence, geo_pixel_thres, geo_depth_thres, photo_thres, geo_mask_thres: int = 3):
    """eval.py:238-265 for one reference view.  Returns (depth_est_averaged [H,W] float64, photo_mask, geo_mask,
    final_mask), numpy if depth_ref is numpy, CUDA tensors otherwise."""
    as_numpy = isinstance(depth_ref, np.ndarray)
    dev = _device_of(depth_ref, *depth_srcs)
    dr, cf = _dev_map(depth_ref, dev), _dev_map(confidence, dev)
    h, w = dr.shape
    srcs = torch.stack([_dev_map(d, dev) for d in depth_srcs]).contiguous()
    s = srcs.shape[0]
    if s == 0 or srcs.shape[1:] != dr.shape or cf.shape != dr.shape:
        raise ValueError("filter_depth_view: need >= 1 source depth map, all maps of the reference's shape")
    cams = np.ascontiguousarray(np.stack([pair_cameras(intrinsics_ref, extrinsics_ref, k, e)
                                          for k, e in zip(intrinsics_srcs, extrinsics_srcs)]))
    acc = torch.empty(h, w, device=dev)
    cnt = torch.empty(h, w, dtype=torch.int32, device=dev)
    avg = torch.empty(h, w, dtype=torch.float64, device=dev)
    pm, gm, fm = (torch.empty(h, w, dtype=torch.uint8, device=dev) for _ in range(3))
    _lib.check(_lib.lib().imvs_filter_depth_view(dr.data_ptr(), cf.data_ptr(), srcs.data_ptr(), cams.ctypes.data, s,
                                                 float(geo_pixel_thres), float(geo_depth_thres), float(photo_thres),
                                                 int(geo_mask_thres), acc.data_ptr(), cnt.data_ptr(), avg.data_ptr(), pm.data_ptr(),
                                                 gm.data_ptr(), fm.data_ptr(), h, w, ops._stream()), "filter_depth_view")
    out = (avg, pm.bool(), gm.bool(), fm.bool())
    return tuple(t.cpu().numpy() for t in out) if as_numpy else out
