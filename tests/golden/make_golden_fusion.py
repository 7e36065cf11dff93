#!/usr/bin/env python
"""Golden vectors for the depth-map filtering step, made by EXECUTING THE REFERENCE'S OWN FUNCTIONS
(build container only: needs /root/reference and OpenCV).

    python tests/golden/make_golden_fusion.py     ->  tests/golden/fusion_kat.npz

eval.py cannot be imported (it parses the command line and imports plyfile at import time), so the two function
definitions `reproject_with_depth` and `check_geometric_consistency` are located by name in its syntax tree and
executed in a namespace holding numpy and cv2 -- the reference's code runs unmodified, nothing is copied into this
repository.  The accumulation over source views follows eval.py:243-265 statement by statement.

Scene: the synthetic plane of itermvs_b200.synthetic seen by 1 reference + 4 source cameras at 160 x 128;
per-view depth maps = exact plane depth + noise, with outliers, zeros (invalid depth) and a band that projects
outside the source images, so every branch of the mask is exercised.
"""
import ast
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from itermvs_b200.synthetic import PLANE_DEPTH, PLANE_NORMAL, extrinsics, intrinsics_full  # noqa: E402

W, H, NSRC = 160, 128, 4


def reference_functions():
    tree = ast.parse(open("/root/reference/eval.py").read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("reproject_with_depth", "check_geometric_consistency")]
    assert len(keep) == 2
    ns = {"np": np, "cv2": cv2}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "/root/reference/eval.py", "exec"), ns)
    return ns["check_geometric_consistency"], ns["reproject_with_depth"]


def plane_depth_in_view(K, E, width, height):
    """depth of the plane n.X = d (reference frame) along the rays of camera E (world == reference frame)."""
    n = np.array(PLANE_NORMAL) / np.linalg.norm(PLANE_NORMAL)
    R, t = E[:3, :3], E[:3, 3]
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing="ij")
    rays = np.linalg.inv(K) @ np.stack([xs.ravel(), ys.ravel(), np.ones(height * width)])       # camera frame, z = 1
    # X_world = R^T (z * ray - t);  n.X_world = d  ->  z = (d + n.R^T t) / (n.R^T ray)
    nr = n @ R.T
    z = (PLANE_DEPTH + nr @ t) / (nr @ rays)
    return z.reshape(height, width)


def main():
    check, reproject = reference_functions()
    rng = np.random.RandomState(3)
    K = intrinsics_full(W, H).astype(np.float32)
    cams = [(K.copy(), extrinsics(v).astype(np.float32)) for v in range(NSRC + 1)]
    depths = []
    for v, (k, e) in enumerate(cams):
        d = plane_depth_in_view(k.astype(np.float64), e.astype(np.float64), W, H)
        d = d * (1 + 0.004 * rng.randn(H, W))                 # around the 1 % relative threshold after reprojection
        out = rng.rand(H, W) < 0.05
        d[out] *= 1 + 0.2 * rng.randn(int(out.sum()))         # outliers
        d[rng.rand(H, W) < 0.02] = 0                          # invalid estimates
        depths.append(d.astype(np.float32))
    depths[0][:, :10] *= 0.5                                  # a band that reprojects far away / outside
    conf = rng.rand(H, W).astype(np.float32)
    g = {"width": W, "height": H, "n_src": NSRC, "confidence": conf, "geo_pixel_thres": 1.0, "geo_depth_thres": 0.01,
         "photo_thres": 0.3, "geo_mask_thres": 3}
    for v, ((k, e), d) in enumerate(zip(cams, depths)):
        g[f"K{v}"], g[f"E{v}"], g[f"depth{v}"] = k, e, d
    # eval.py:243-265
    geo_mask_sum = 0
    all_srcview_depth_ests = []
    for v in range(1, NSRC + 1):
        geo_mask, depth_reprojected, x2d_src, y2d_src = check(depths[0], cams[0][0], cams[0][1], depths[v], cams[v][0], cams[v][1], 1.0, 0.01)
        g[f"mask{v}"], g[f"reprojected{v}"], g[f"x_src{v}"], g[f"y_src{v}"] = geo_mask, depth_reprojected, x2d_src, y2d_src
        geo_mask_sum += geo_mask.astype(np.int32)
        all_srcview_depth_ests.append(depth_reprojected)
    depth_est_averaged = (sum(all_srcview_depth_ests) + depths[0]) / (geo_mask_sum + 1)
    photo_mask = conf > 0.3
    geo_mask = geo_mask_sum >= 3
    g["geo_mask_sum"], g["depth_est_averaged"] = geo_mask_sum, depth_est_averaged
    g["photo_mask"], g["geo_mask"], g["final_mask"] = photo_mask, geo_mask, np.logical_and(photo_mask, geo_mask)
    # raw reprojection outputs of one pair (before masking)
    rep = reproject(depths[0], cams[0][0], cams[0][1], depths[1], cams[1][0], cams[1][1])
    g["raw_depth_reprojected1"], g["raw_x_reprojected1"], g["raw_y_reprojected1"] = rep[0], rep[1], rep[2]
    # cv2.remap on its own, incl. far-out-of-range and non-finite coordinates
    img = rng.rand(24, 32).astype(np.float32)
    mx = (rng.rand(40, 50) * 40 - 4).astype(np.float32)
    my = (rng.rand(40, 50) * 30 - 3).astype(np.float32)
    mx[0, :6] = [-1e9, 1e9, np.inf, -np.inf, np.nan, 31.0]
    my[1, :4] = [23.0, 23.5, -0.5, 1e30]
    g["remap_img"], g["remap_x"], g["remap_y"] = img, mx, my
    g["remap_out"] = cv2.remap(img, mx, my, interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(os.path.join(HERE, "fusion_kat.npz"), **g)
    print("fusion_kat.npz: consistent fraction per source", [float(g[f"mask{v}"].mean()) for v in range(1, NSRC + 1)],
          "geo", float(geo_mask.mean()), "final", float(g["final_mask"].mean()), "dtype avg", depth_est_averaged.dtype,
          os.path.getsize(os.path.join(HERE, "fusion_kat.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
