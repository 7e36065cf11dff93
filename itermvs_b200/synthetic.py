"""Seeded synthetic multi-view inputs in the exact dict format `Pipeline.forward` takes.

The reference ships no data; its loaders (reference datasets/dtu_yao_eval.py:61-158) produce
    imgs          {'level_0': [B,V,3,H,W], ... 'level_3'}     float32 in [-1,1]
    proj_matrices {'level_0'..'level_3': [B,V,4,4]}           4x4 = [K_l @ E[:3,:4]; 0 0 0 1]
    depth_min / depth_max  [B]
This module builds the same structure from a seed:

* cameras: DTU-like pinhole K scaled to (W,H); per level K[:2] *= 0.125 * 2^k exactly as
  dtu_yao_eval.py:106-126; reference extrinsic = identity; source i rotated about y by
  +-0.06*i rad (and 0.02*i about z) and translated so the cameras converge on depth ~650.
* 'plane' scene (geometrically consistent): a smooth random texture on the plane n.X = 650,
  source images rendered through the exact plane-induced homography  K (R + t n^T / d) K^-1.
  On this scene the estimator's arg-max is well conditioned, so end-to-end parity is well posed.
* 'noise' scene: i.i.d. U(-1,1) images (stage-wise checks / throughput only; arg-max chaotic).

Only numpy + torch are used so it runs identically here and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DEPTH_MIN = 425.0
DEPTH_MAX = 935.0
PLANE_DEPTH = 650.0
PLANE_NORMAL = (0.1, 0.05, 1.0)


def _rodrigues(rvec: np.ndarray) -> np.ndarray:
    theta = float(np.linalg.norm(rvec))
    if theta < 1e-12:
        return np.eye(3)
    k = rvec / theta
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=np.float64)
    return np.eye(3) + math.sin(theta) * K + (1 - math.cos(theta)) * (K @ K)


def intrinsics_full(width: int, height: int) -> np.ndarray:
    return np.array([[2892.33 * width / 1600.0, 0.0, width / 2.0],
                     [0.0, 2883.18 * height / 1200.0, height / 2.0],
                     [0.0, 0.0, 1.0]], dtype=np.float64)


def extrinsics(view: int) -> np.ndarray:
    """view 0 = reference (identity); view i>=1 = i-th source."""
    E = np.eye(4, dtype=np.float64)
    if view == 0:
        return E
    i = (view + 1) // 2
    sgn = 1.0 if view % 2 == 1 else -1.0
    E[:3, :3] = _rodrigues(np.array([0.0, sgn * 0.06 * i, 0.02 * i]))
    E[:3, 3] = np.array([-sgn * 0.06 * i * PLANE_DEPTH, 5.0 * i, 3.0 * i])
    return E


def projection_pyramid(width: int, height: int, n_views: int) -> Dict[str, np.ndarray]:
    """proj['level_k'] : [V,4,4] float64, built the way dtu_yao_eval.py:106-126 does."""
    out = {f"level_{k}": [] for k in range(4)}
    for v in range(n_views):
        E = extrinsics(v)
        K = intrinsics_full(width, height).copy()
        K[:2, :] *= 0.125
        for k in (3, 2, 1, 0):
            P = E.copy()
            P[:3, :4] = K @ P[:3, :4]
            out[f"level_{k}"].append(P)
            K[:2, :] *= 2
    return {k: np.stack(v) for k, v in out.items()}


def _texture(rng: np.random.RandomState, height: int, width: int) -> torch.Tensor:
    """Textured RGB image in [-1,1], [3,H,W]: band-limited uniform noise summed over octaves
    (coarsest ~1/4 of the image, finest ~3 px) so that every pyramid level has matchable detail."""
    tex = torch.zeros(3, height, width, dtype=torch.float32)
    amp_total = 0.0
    cells = 4
    amp = 1.0
    while min(height, width) / cells >= 3.0:
        h, w = max(2, round(height * cells / min(height, width))), max(2, round(width * cells / min(height, width)))
        n = torch.from_numpy(rng.uniform(-1, 1, size=(1, 3, h, w)).astype(np.float32))
        tex += amp * F.interpolate(n, size=(height, width), mode="bicubic", align_corners=True)[0]
        amp_total += amp * amp
        cells *= 2
        amp *= 0.8
    tex = tex / (1.2 * math.sqrt(amp_total))
    return tex.clamp_(-1, 1)


def _render_through_homography(ref_img: torch.Tensor, H_ref_to_src: np.ndarray) -> torch.Tensor:
    """src(x) = ref(H^-1 x), bilinear, zeros outside. ref_img [3,H,W]."""
    _, h, w = ref_img.shape
    Hinv = np.linalg.inv(H_ref_to_src)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    pts = np.stack([xs.ravel(), ys.ravel(), np.ones(h * w)])
    q = Hinv @ pts
    u = (q[0] / q[2]).reshape(h, w)
    v = (q[1] / q[2]).reshape(h, w)
    grid = np.stack([u / ((w - 1) / 2.0) - 1.0, v / ((h - 1) / 2.0) - 1.0], axis=-1).astype(np.float32)
    return F.grid_sample(ref_img[None], torch.from_numpy(grid)[None], mode="bilinear",
                         padding_mode="border", align_corners=True)[0]


def _image_pyramid(img0: torch.Tensor) -> Dict[str, torch.Tensor]:
    """level_k by bilinear resize (the loader uses cv2.INTER_LINEAR, dtu_yao_eval.py:69-76)."""
    out = {"level_0": img0}
    for k in (1, 2, 3):
        out[f"level_{k}"] = F.interpolate(img0[None], scale_factor=1.0 / 2 ** k, mode="bilinear",
                                          align_corners=False)[0]
    return out


def plane_depth_map(width: int, height: int) -> np.ndarray:
    """Ground-truth reference-view depth of the synthetic plane at full resolution, [H,W]."""
    K = intrinsics_full(width, height)
    n = np.array(PLANE_NORMAL) / np.linalg.norm(PLANE_NORMAL)
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing="ij")
    rays = np.linalg.inv(K) @ np.stack([xs.ravel(), ys.ravel(), np.ones(height * width)])
    depth = PLANE_DEPTH / (n @ rays)
    return depth.reshape(height, width)


def make_sample(width: int = 640, height: int = 512, n_src: int = 4, batch: int = 1, seed: int = 0,
                scene: str = "plane") -> Dict[str, object]:
    """Return {'imgs','proj_matrices','depth_min','depth_max'} exactly as a DataLoader batch."""
    assert width % 32 == 0 and height % 32 == 0, "H and W must be multiples of 32 (CorrNet strides)"
    n_views = n_src + 1
    imgs = {f"level_{k}": [] for k in range(4)}
    for b in range(batch):
        rng = np.random.RandomState(seed + 7919 * b)
        views = []
        if scene == "plane":
            ref = _texture(rng, height, width)
            K = intrinsics_full(width, height)
            n = np.array(PLANE_NORMAL) / np.linalg.norm(PLANE_NORMAL)
            views.append(ref)
            for v in range(1, n_views):
                E = extrinsics(v)
                Hm = K @ (E[:3, :3] + np.outer(E[:3, 3], n) / PLANE_DEPTH) @ np.linalg.inv(K)
                views.append(_render_through_homography(ref, Hm))
        elif scene == "noise":
            for v in range(n_views):
                views.append(torch.from_numpy(rng.uniform(-1, 1, size=(3, height, width)).astype(np.float32)))
        else:
            raise ValueError(f"unknown scene {scene!r}")
        pyr = [_image_pyramid(v) for v in views]
        for k in range(4):
            imgs[f"level_{k}"].append(torch.stack([p[f"level_{k}"] for p in pyr]))
    imgs = {k: torch.stack(v).contiguous() for k, v in imgs.items()}
    proj_np = projection_pyramid(width, height, n_views)
    proj = {k: torch.from_numpy(np.broadcast_to(v[None], (batch,) + v.shape).copy()) for k, v in proj_np.items()}
    return {
        "imgs": imgs,
        "proj_matrices": proj,
        "depth_min": torch.full((batch,), DEPTH_MIN, dtype=torch.float32),
        "depth_max": torch.full((batch,), DEPTH_MAX, dtype=torch.float32),
    }


def random_feature_pyramids(width: int, height: int, n_src: int, batch: int = 1, seed: int = 0
                            ) -> Tuple[Dict[str, torch.Tensor], Dict[str, list]]:
    """N(0,1) feature pyramids (NCHW) for stage-wise checks of the estimator without FeatureNet."""
    g = torch.Generator().manual_seed(seed)
    dims = {"level1": (16, 2), "level2": (32, 4), "level3": (48, 8)}
    ref, srcs = {}, {}
    for name, (c, s) in dims.items():
        ref[name] = torch.randn(batch, c, height // s, width // s, generator=g)
        srcs[name] = [torch.randn(batch, c, height // s, width // s, generator=g) for _ in range(n_src)]
    return ref, srcs
