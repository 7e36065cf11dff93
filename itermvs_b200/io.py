"""On-disk formats either side of the hot path, byte-compatible with the reference:

    read_pfm / save_pfm            datasets/data_io.py:6-73   (depth_est/*.pfm, confidence/*.pfm)
    read_cam_file                  datasets/dtu_yao_eval.py:42-53 (cams_1/*_cam.txt: extrinsic 4x4, intrinsic 3x3, depth range)
    read_pair_file                 eval.py:90-100 (pair.txt)
    projection_pyramid             datasets/dtu_yao_eval.py:106-126 (4x4 = [K_l @ E[:3,:4]; E[3]] for levels 3..0)
    image_pyramid / read_img       datasets/dtu_yao_eval.py:61-76
    load_views                     datasets/dtu_yao_eval.py:78-158 (__getitem__): the dict Pipeline.forward takes
    backproject_points / save_ply  eval.py:281-309 (the filtered depth maps as a coloured point cloud; read_ply reads it back)

Host-side numpy; nothing here touches the GPU (Pipeline.forward reads imgs['level_0'] and proj_matrices['level_1..3'] only).
"""
from __future__ import annotations

import os
import re
import sys
from typing import Dict, List, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------------------------------- PFM --
def read_pfm(filename: str) -> Tuple[np.ndarray, float]:
    """-> (data [H,W,1] (Pf) or [H,W,3] (PF), float32, top row first; scale).  Byte order from the sign of the scale."""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header == "PF":
            channels = 3
        elif header == "Pf":
            channels = 1
        else:
            raise Exception("Not a PFM file.")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise Exception("Malformed PFM header.")
        width, height = int(m.group(1)), int(m.group(2))
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.fromfile(f, endian + "f")
    return np.flipud(data.reshape(height, width, channels)), abs(scale)


def save_pfm(filename: str, image: np.ndarray, scale: float = 1) -> None:
    """float32 [H,W], [H,W,1] (written as 'Pf') or [H,W,3] ('PF'); rows bottom-up, scale negative = little endian."""
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        tag = b"PF\n"
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        tag = b"Pf\n"
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    order = image.dtype.byteorder
    if order == "<" or (order == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(filename, "wb") as f:
        f.write(tag)
        f.write(("%d %d\n" % (image.shape[1], image.shape[0])).encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        np.flipud(image).tofile(f)


# ----------------------------------------------------------------------------------------- cameras --
def read_cam_file(filename: str):
    """-> intrinsics [3,3] f32, extrinsics [4,4] f32, depth_min, depth_max (first / last number of line 11)."""
    with open(filename) as f:
        lines = [line.rstrip() for line in f.readlines()]
    extrinsics = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape(4, 4)
    intrinsics = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape(3, 3)
    rng = lines[11].split()
    return intrinsics, extrinsics, float(rng[0]), float(rng[-1])


def read_pair_file(filename: str) -> List[Tuple[int, List[int]]]:
    """pair.txt: view count, then per view its id and '<n> id score id score ...'; views without sources are dropped."""
    out = []
    with open(filename) as f:
        n = int(f.readline())
        for _ in range(n):
            ref = int(f.readline().rstrip())
            srcs = [int(x) for x in f.readline().rstrip().split()[1::2]]
            if srcs:
                out.append((ref, srcs))
    return out


def projection_pyramid(intrinsics: np.ndarray, extrinsics: np.ndarray, img_wh: Sequence[int], orig_wh: Sequence[int] = (1600, 1200)
                       ) -> Dict[str, np.ndarray]:
    """Per-level 4x4 projection matrices: intrinsics rescaled from the original to the working resolution, then
    K[:2] *= 1/8, 1/4, 1/2, 1 for levels 3..0 (successive float32 doublings, as the loader does)."""
    k = np.array(intrinsics, dtype=np.float32, copy=True)
    e = np.asarray(extrinsics, dtype=np.float32)
    k[0] *= img_wh[0] / orig_wh[0]
    k[1] *= img_wh[1] / orig_wh[1]
    out = {}
    k[:2, :] *= 0.125
    for level in (3, 2, 1, 0):
        p = e.copy()
        p[:3, :4] = np.matmul(k, p[:3, :4])
        out[f"level_{level}"] = p
        k[:2, :] *= 2
    return out


# ------------------------------------------------------------------------------------------ images --
def image_pyramid(img0: np.ndarray) -> Dict[str, np.ndarray]:
    """[H,W,3] float32 -> level_0..3 by cv2.resize(INTER_LINEAR) to (W/2^k, H/2^k)."""
    import cv2
    h, w = img0.shape[:2]
    out = {"level_0": img0}
    for k in (1, 2, 3):
        out[f"level_{k}"] = cv2.resize(img0, (w // 2 ** k, h // 2 ** k), interpolation=cv2.INTER_LINEAR)
    return out


def read_img(filename: str, img_wh: Sequence[int]) -> Dict[str, np.ndarray]:
    """8-bit image -> [-1, 1] float32, resized to img_wh, 4-level pyramid."""
    import cv2
    from PIL import Image
    img = 2 * np.array(Image.open(filename), dtype=np.float32) / 255.0 - 1
    img = cv2.resize(img, tuple(img_wh), interpolation=cv2.INTER_LINEAR)
    return image_pyramid(img)


def load_views(datapath: str, scan: str, ref_view: int, src_views: Sequence[int], nviews: int = 5,
               img_wh: Sequence[int] = (1600, 1152), orig_wh: Sequence[int] = (1600, 1200)) -> Dict[str, object]:
    """One sample in the loader's format: imgs {'level_k': [V,3,H_k,W_k]}, proj_matrices {'level_k': [V,4,4]},
    depth_min / depth_max of the reference view, filename pattern."""
    view_ids = [ref_view] + list(src_views)[: nviews - 1]
    imgs = {f"level_{k}": [] for k in range(4)}
    proj = {f"level_{k}": [] for k in range(4)}
    depth_min = depth_max = None
    for i, vid in enumerate(view_ids):
        pyr = read_img(os.path.join(datapath, "{}/images/{:0>8}.jpg".format(scan, vid)), img_wh)
        k, e, dmin, dmax = read_cam_file(os.path.join(datapath, "{}/cams_1/{:0>8}_cam.txt".format(scan, vid)))
        pp = projection_pyramid(k, e, img_wh, orig_wh)
        for lv in imgs:
            imgs[lv].append(pyr[lv])
            proj[lv].append(pp[lv])
        if i == 0:
            depth_min, depth_max = dmin, dmax
    return {"imgs": {lv: np.stack(v).transpose([0, 3, 1, 2]) for lv, v in imgs.items()},
            "proj_matrices": {lv: np.stack(v) for lv, v in proj.items()},
            "depth_min": depth_min, "depth_max": depth_max,
            "filename": scan + "/{}/" + "{:0>8}".format(view_ids[0]) + "{}"}


# ------------------------------------------------------------------ images on the device (f-4) --
def image_pyramid_device(img_u8, img_wh: Sequence[int], levels: int = 4, stream=None):
    """`read_img` (dtu_yao_eval.py:61-76) after the decode, on the GPU: `img_u8` is the raw 8-bit image as a CUDA uint8 tensor
    [H0, W0, 3]; returns {'level_k': float32 CUDA tensor [3, H >> k, W >> k]} with level_0 = cv2.resize(2 * img / 255. - 1,
    img_wh, INTER_LINEAR) and level_k = cv2.resize(level_0, (W >> k, H >> k), INTER_LINEAR) (C entry point
    `imvs_image_pyramid_u8`, csrc/imageprep.cu).  There is no CPU path: a CPU tensor raises."""
    import torch
    from . import _lib
    if not (isinstance(img_u8, torch.Tensor) and img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.shape[2] == 3):
        raise RuntimeError("itermvs_b200.io.image_pyramid_device: expected a CUDA uint8 tensor [H0, W0, 3] (there is no CPU path)")
    img_u8 = img_u8.contiguous()
    w, h = int(img_wh[0]), int(img_wh[1])
    out = {f"level_{k}": torch.empty(3, h >> k, w >> k, device=img_u8.device, dtype=torch.float32) for k in range(levels)}
    ptr = lambda k: out[f"level_{k}"].data_ptr() if k < levels else None
    st = stream if stream is not None else torch.cuda.current_stream(img_u8.device).cuda_stream
    _lib.check(_lib.lib().imvs_image_pyramid_u8(img_u8.data_ptr(), img_u8.shape[0], img_u8.shape[1], ptr(0), ptr(1), ptr(2), ptr(3), h, w, st),
               "image_pyramid_u8")
    return out


class PrefetchLoader:
    """Serving-side replacement of the evaluation DataLoader (eval.py:45-50 / dtu_yao_eval.py `__getitem__`): a background
    thread decodes the views of reference view i + 1 (PIL, host) into PINNED uint8 buffers and reads its cameras while
    reference view i is on the GPU; `__next__` uploads the raw 8-bit images on a copy stream and runs normalisation, resize and
    the pyramid there (`image_pyramid_device`), so the host never touches float images.  Yields the loader's sample dict with
    CUDA tensors: imgs {'level_k': [1, V, 3, H_k, W_k]}, proj_matrices {'level_k': [1, V, 4, 4]}, depth_min / depth_max [1].

    items: sequence of (image paths of the V views, camera-file paths of the V views), reference view first."""

    def __init__(self, items, img_wh: Sequence[int], orig_wh: Sequence[int] = (1600, 1200), device="cuda", depth: int = 2):
        import queue
        import threading
        import torch
        self.items, self.img_wh, self.orig_wh = list(items), tuple(img_wh), tuple(orig_wh)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._q = queue.Queue(maxsize=depth)
        self._pool = {}                 # (V, H0, W0) -> ring of pinned buffers, reused: no pinned allocation per sample
        self._depth = depth
        self._t = threading.Thread(target=self._work, daemon=True)
        self._t.start()

    def _pinned(self, slot: int, shape):
        import torch
        key = (slot,) + tuple(shape)
        if key not in self._pool:
            self._pool[key] = torch.empty(shape, dtype=torch.uint8).pin_memory()
        return self._pool[key]

    def _work(self):
        from PIL import Image
        try:
            for n, (img_paths, cam_paths) in enumerate(self.items):
                raw = [np.array(Image.open(f).convert("RGB"), dtype=np.uint8) for f in img_paths]
                h0, w0 = raw[0].shape[:2]
                buf = self._pinned(n % (self._depth + 1), (len(raw), h0, w0, 3))
                for v, r in enumerate(raw):
                    if r.shape[:2] != (h0, w0):
                        raise ValueError(f"view {v} of sample {n} has a different size")
                    buf[v].numpy()[...] = r
                proj = {f"level_{k}": [] for k in range(4)}
                dmin = dmax = None
                for v, f in enumerate(cam_paths):
                    k_, e_, dmin_, dmax_ = read_cam_file(f)
                    pp = projection_pyramid(k_, e_, self.img_wh, self.orig_wh)
                    for lv in proj:
                        proj[lv].append(pp[lv])
                    if v == 0:
                        dmin, dmax = dmin_, dmax_
                self._q.put((buf, {lv: np.stack(p) for lv, p in proj.items()}, dmin, dmax))
            self._q.put(None)
        except BaseException as e:      # surfaced by __next__
            self._q.put(e)

    def __iter__(self):
        return self

    def __next__(self):
        import torch
        item = self._q.get()
        if item is None:
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        buf, proj, dmin, dmax = item
        with torch.cuda.stream(self.copy_stream):
            dev_u8 = buf.to(self.device, non_blocking=True)
            per_view = [image_pyramid_device(dev_u8[v], self.img_wh, stream=self.copy_stream.cuda_stream) for v in range(dev_u8.shape[0])]
            imgs = {lv: torch.stack([p[lv] for p in per_view]).unsqueeze(0) for lv in per_view[0]}
            pm = {lv: torch.from_numpy(p).unsqueeze(0).to(self.device, non_blocking=True) for lv, p in proj.items()}
            dm = torch.tensor([dmin], dtype=torch.float32).to(self.device, non_blocking=True)
            dx = torch.tensor([dmax], dtype=torch.float32).to(self.device, non_blocking=True)
        torch.cuda.current_stream(self.device).wait_stream(self.copy_stream)
        dev_u8.record_stream(torch.cuda.current_stream(self.device))
        return {"imgs": imgs, "proj_matrices": pm, "depth_min": dm, "depth_max": dx}


# --------------------------------------------------------------------------------- point cloud --
def backproject_points(depth_est_averaged: np.ndarray, final_mask: np.ndarray, ref_img: np.ndarray, ref_intrinsics: np.ndarray,
                       ref_extrinsics: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """The tail of filter_depth's loop body (eval.py:281-296): the pixels that survive the photometric + geometric
    filter (what `filter_depth_view` returns) lifted to world coordinates, with their colours.
    -> (vertices [N,3] float64 world xyz, colours [N,3] uint8), same arithmetic / promotions as the numpy original."""
    height, width = depth_est_averaged.shape[:2]
    xs, ys = np.meshgrid(np.arange(0, width), np.arange(0, height))
    valid = np.asarray(final_mask, dtype=bool)
    xs, ys, depth = xs[valid], ys[valid], depth_est_averaged[valid]
    xyz_ref = np.matmul(np.linalg.inv(ref_intrinsics), np.vstack((xs, ys, np.ones_like(xs))) * depth)
    xyz_world = np.matmul(np.linalg.inv(ref_extrinsics), np.vstack((xyz_ref, np.ones_like(xs))))[:3]
    return xyz_world.transpose((1, 0)), (ref_img[valid] * 255).astype(np.uint8)


_PLY_HEADER = ("ply\nformat binary_little_endian 1.0\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\n"
               "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
_PLY_VERTEX = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])


def save_ply(filename: str, vertices: np.ndarray, colors: np.ndarray) -> None:
    """The fused point cloud as eval.py:298-309 writes it through plyfile (`PlyData([PlyElement.describe(vertex_all,
    'vertex')]).write(...)`): binary little-endian, one 15-byte record per point (x, y, z float32; red, green, blue
    uchar).  plyfile is not installed in the build image, so the header text is restated from its writer (format line,
    element line, one property line per field, end_header) rather than compared byte for byte -- parity unpinned for
    this one function; read_ply below and any PLY reader (MeshLab, Open3D, the DTU evaluation scripts) accept it."""
    vertices = np.asarray(vertices)
    colors = np.asarray(colors)
    if vertices.ndim != 2 or vertices.shape[1] != 3 or colors.shape != vertices.shape:
        raise ValueError("save_ply: vertices and colors must both be [N,3]")
    rec = np.empty(len(vertices), dtype=_PLY_VERTEX)
    for i, name in enumerate(("x", "y", "z")):
        rec[name] = vertices[:, i].astype(np.float32)
    for i, name in enumerate(("red", "green", "blue")):
        rec[name] = colors[:, i].astype(np.uint8)
    with open(filename, "wb") as f:
        f.write(_PLY_HEADER.format(n=len(rec)).encode("ascii"))
        rec.tofile(f)


def read_ply(filename: str) -> Tuple[np.ndarray, np.ndarray]:
    """Reader for the files save_ply writes: -> (vertices [N,3] float32, colours [N,3] uint8)."""
    with open(filename, "rb") as f:
        header = b""
        while not header.endswith(b"end_header\n"):
            line = f.readline()
            if not line:
                raise ValueError("read_ply: no end_header")
            header += line
        text = header.decode("ascii")
        if "format binary_little_endian 1.0" not in text:
            raise ValueError("read_ply: only the binary little-endian vertex files of save_ply are supported")
        n = int(re.search(r"element vertex (\d+)", text).group(1))
        rec = np.fromfile(f, dtype=_PLY_VERTEX, count=n)
    return (np.stack([rec["x"], rec["y"], rec["z"]], axis=1), np.stack([rec["red"], rec["green"], rec["blue"]], axis=1))
