#!/usr/bin/env python
"""Golden vectors for the on-disk formats, made with the REFERENCE'S OWN reader / writer / loader
(build container only).  ->  tests/golden/io_kat.npz

* PFM: bytes written by datasets/data_io.py:save_pfm for gray [H,W], [H,W,1] and colour [H,W,3] maps, and what its
  read_pfm returns for them and for a big-endian file.
* a miniature scan folder (3 views, 64 x 48 JPEGs, DTU-format cam files, pair.txt) and what the reference loader
  datasets/dtu_yao_eval.py:MVSDataset.__getitem__ and eval.py:read_pair_file return for it.  The folder's files are
  stored as bytes so that the test can rebuild it.
"""
import ast
import io
import os
import sys
import tempfile

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from datasets.data_io import read_pfm, save_pfm  # noqa: E402  (the reference)
from datasets.dtu_yao_eval import MVSDataset  # noqa: E402


def ref_read_pair_file():
    tree = ast.parse(open("/root/reference/eval.py").read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "read_pair_file"]
    ns = {}
    exec(compile(ast.Module(body=fn, type_ignores=[]), "/root/reference/eval.py", "exec"), ns)
    return ns["read_pair_file"]


def cam_text(k, e, dmin, interval, ndepth, dmax):
    rows = ["extrinsic"] + [" ".join("%.8g" % v for v in r) for r in e] + ["", "intrinsic"] + \
           [" ".join("%.8g" % v for v in r) for r in k] + ["", "%g %g %d %g" % (dmin, interval, ndepth, dmax)]
    return "\n".join(rows) + "\n"


def main():
    rng = np.random.RandomState(11)
    g = {}
    tmp = tempfile.mkdtemp()
    # ---- PFM
    arrays = {"gray": (rng.rand(7, 5) * 900).astype(np.float32), "gray1": (rng.rand(4, 6, 1) * 2 - 1).astype(np.float32),
              "color": rng.rand(3, 4, 3).astype(np.float32)}
    for name, a in arrays.items():
        fn = os.path.join(tmp, name + ".pfm")
        save_pfm(fn, a) if name != "gray1" else save_pfm(fn, a, scale=2.5)
        g[f"pfm_{name}_in"] = a
        g[f"pfm_{name}_bytes"] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
        d, s = read_pfm(fn)
        g[f"pfm_{name}_read"], g[f"pfm_{name}_scale"] = np.ascontiguousarray(d), np.float64(s)
    be = os.path.join(tmp, "be.pfm")
    with open(be, "wb") as f:
        f.write(b"Pf\n3 2\n1.000000\n")
        np.arange(6, dtype=">f4").tofile(f)
    d, s = read_pfm(be)
    g["pfm_be_bytes"] = np.frombuffer(open(be, "rb").read(), dtype=np.uint8)
    g["pfm_be_read"], g["pfm_be_scale"] = np.ascontiguousarray(d), np.float64(s)
    # ---- miniature scan
    scan = "scan1"
    os.makedirs(os.path.join(tmp, scan, "images"))
    os.makedirs(os.path.join(tmp, scan, "cams_1"))
    for v in range(3):
        img = (rng.rand(48, 64, 3) * 255).astype(np.uint8)
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="JPEG", quality=95)
        open(os.path.join(tmp, scan, "images", "%08d.jpg" % v), "wb").write(buf.getvalue())
        g[f"jpg{v}"] = np.frombuffer(buf.getvalue(), dtype=np.uint8)
        k = np.array([[2892.33, 0, 823.2 + v], [0, 2883.18, 619.07], [0, 0, 1]])
        e = np.eye(4)
        e[:3, 3] = [10.0 * v, -3.0, 2.5]
        e[0, 1], e[1, 0] = 0.01 * v, -0.01 * v
        txt = cam_text(k, e, 425.0 + v, 2.5, 192, 935.0 + v)
        open(os.path.join(tmp, scan, "cams_1", "%08d_cam.txt" % v), "w").write(txt)
        g[f"cam{v}"] = np.frombuffer(txt.encode(), dtype=np.uint8)
    pair = "3\n0\n2 1 2036.53 2 1243.89\n1\n2 0 2036.53 2 1113.2\n2\n0\n"
    open(os.path.join(tmp, scan, "pair.txt"), "w").write(pair)
    g["pair"] = np.frombuffer(pair.encode(), dtype=np.uint8)
    lst = os.path.join(tmp, "list.txt")
    open(lst, "w").write(scan + "\n")
    ds = MVSDataset(tmp, lst, nviews=3, img_wh=(64, 32))
    g["n_metas"] = len(ds)
    s = ds[0]
    for lv in ("level_0", "level_1", "level_2", "level_3"):
        g[f"imgs_{lv}"], g[f"proj_{lv}"] = s["imgs"][lv], s["proj_matrices"][lv]
    g["depth_min"], g["depth_max"] = np.float64(s["depth_min"]), np.float64(s["depth_max"])
    g["filename"] = np.frombuffer(s["filename"].encode(), dtype=np.uint8)
    k, e, dmin, dmax = ds.read_cam_file(os.path.join(tmp, scan, "cams_1", "%08d_cam.txt" % 1))
    g["cam1_K"], g["cam1_E"], g["cam1_range"] = k, e, np.array([dmin, dmax])
    pairs = ref_read_pair_file()(os.path.join(tmp, scan, "pair.txt"))
    g["pairs_ref"] = np.array([p[0] for p in pairs])
    g["pairs_src"] = np.array([p[1] for p in pairs])
    np.savez_compressed(os.path.join(HERE, "io_kat.npz"), **g)
    print("io_kat.npz", len(g), "arrays", os.path.getsize(os.path.join(HERE, "io_kat.npz")) / 1e3, "kB; pairs", pairs)


if __name__ == "__main__":
    main()
