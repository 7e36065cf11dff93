"""CPU oracle for the depth-map filtering step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of reference eval.py:154-215 (reproject_with_depth, check_geometric_consistency) and of the
accumulation in filter_depth (eval.py:243-265), with `cv2.remap(..., INTER_LINEAR)` restated explicitly (OpenCV's
float maps are converted to fixed point with 5 fractional bits, weights come from a 32 x 32 float table, constant
border 0), so the test does not depend on OpenCV being installed on the GPU box.

PINNING: tests/golden/make_golden_fusion.py executes the reference's own two functions (extracted from
/root/reference/eval.py by name, not copied) with the real cv2.remap in the build container and stores inputs and
outputs in tests/golden/fusion_kat.npz; tests/test_fusion.py replays them against this file.
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def remap_linear(img: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    """cv2.remap(img, map_x, map_y, cv2.INTER_LINEAR) for float32 single-channel `img` (BORDER_CONSTANT, 0)."""
    h, w = img.shape

    def fixed(m):
        with np.errstate(invalid="ignore", over="ignore"):
            v = m.astype(np.float32) * np.float32(32.0)
            ok = np.isfinite(v) & (v >= -2147483648.0) & (v < 2147483648.0)
            r = np.where(ok, np.rint(np.where(ok, v, 0)), -2147483648.0).astype(np.int64)     # cvRound, ties to even
        return r

    ix, iy = fixed(map_x), fixed(map_y)
    sx = np.clip(ix >> 5, -32768, 32767)
    sy = np.clip(iy >> 5, -32768, 32767)
    ax = ((ix & 31).astype(np.float32) * np.float32(1 / 32))
    ay = ((iy & 31).astype(np.float32) * np.float32(1 / 32))
    one = np.float32(1)
    w00, w01, w10, w11 = (one - ay) * (one - ax), (one - ay) * ax, ay * (one - ax), ay * ax

    def px(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return np.where(ok, img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], np.float32(0)).astype(np.float32)

    r = px(sy, sx) * w00
    r = r + px(sy, sx + 1) * w01
    r = r + px(sy + 1, sx) * w10
    r = r + px(sy + 1, sx + 1) * w11
    return r.astype(np.float32)


def reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """eval.py:154-196."""
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    x_ref, y_ref = x_ref.reshape(-1), y_ref.reshape(-1)
    ones = np.ones_like(x_ref)
    with np.errstate(all="ignore"):
        xyz_ref = np.matmul(np.linalg.inv(intrinsics_ref), np.vstack((x_ref, y_ref, ones)) * depth_ref.reshape(-1))
        xyz_src = np.matmul(np.matmul(extrinsics_src, np.linalg.inv(extrinsics_ref)), np.vstack((xyz_ref, ones)))[:3]
        k_src = np.matmul(intrinsics_src, xyz_src)
        xy_src = k_src[:2] / k_src[2:3]
        x_src = xy_src[0].reshape(height, width).astype(np.float32)
        y_src = xy_src[1].reshape(height, width).astype(np.float32)
        sampled = remap_linear(depth_src, x_src, y_src)
        xyz_src = np.matmul(np.linalg.inv(intrinsics_src), np.vstack((xy_src, ones)) * sampled.reshape(-1))
        xyz_rep = np.matmul(np.matmul(extrinsics_ref, np.linalg.inv(extrinsics_src)), np.vstack((xyz_src, ones)))[:3]
        depth_rep = xyz_rep[2].reshape(height, width).astype(np.float32)
        k_rep = np.matmul(intrinsics_ref, xyz_rep)
        xy_rep = k_rep[:2] / (k_rep[2:3] + 1e-6)
    return (depth_rep, xy_rep[0].reshape(height, width).astype(np.float32), xy_rep[1].reshape(height, width).astype(np.float32),
            x_src, y_src)


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                geo_pixel_thres, geo_depth_thres):
    """eval.py:199-215."""
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_rep, x_rep, y_rep, x_src, y_src = reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src,
                                                                 intrinsics_src, extrinsics_src)
    with np.errstate(all="ignore"):
        dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
        rel = np.abs(depth_rep - depth_ref) / depth_ref
        mask = np.logical_and(dist < geo_pixel_thres, rel < np.float32(geo_depth_thres))
    depth_rep[~mask] = 0
    return mask, depth_rep, x_src, y_src


def filter_depth_view(depth_ref, confidence, intrinsics_ref, extrinsics_ref, depth_srcs, intrinsics_srcs, extrinsics_srcs,
                      geo_pixel_thres, geo_depth_thres, photo_thres, geo_mask_thres=3):
    """eval.py:238-265 for one reference view -> (depth_est_averaged float64, photo_mask, geo_mask, final_mask)."""
    photo_mask = confidence > np.float32(photo_thres)
    geo_sum = np.zeros(depth_ref.shape, np.int32)
    total = np.zeros(depth_ref.shape, np.float32)
    for d, k, e in zip(depth_srcs, intrinsics_srcs, extrinsics_srcs):
        m, rep, _, _ = check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, d, k, e, geo_pixel_thres, geo_depth_thres)
        geo_sum += m.astype(np.int32)
        total = total + rep
    averaged = (total + depth_ref) / (geo_sum + 1)
    geo_mask = geo_sum >= geo_mask_thres
    return averaged, photo_mask, geo_mask, np.logical_and(photo_mask, geo_mask)
