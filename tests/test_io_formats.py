"""On-disk formats against bytes / values produced by the reference's own reader, writer and loader
(tests/golden/make_golden_io.py).  CPU only."""
import os

import numpy as np
import pytest

from itermvs_b200 import io as mio


@pytest.fixture(scope="module")
def kat():
    with np.load(os.path.join(os.path.dirname(__file__), "golden", "io_kat.npz")) as z:
        return {k: z[k] for k in z.files}


def test_save_pfm_is_byte_identical(kat, tmp_path):
    for name, scale in (("gray", 1), ("gray1", 2.5), ("color", 1)):
        fn = str(tmp_path / (name + ".pfm"))
        mio.save_pfm(fn, kat[f"pfm_{name}_in"], scale) if scale != 1 else mio.save_pfm(fn, kat[f"pfm_{name}_in"])
        assert open(fn, "rb").read() == kat[f"pfm_{name}_bytes"].tobytes(), name
    with pytest.raises(Exception):
        mio.save_pfm(str(tmp_path / "bad.pfm"), np.zeros((2, 2), np.float64))
    with pytest.raises(Exception):
        mio.save_pfm(str(tmp_path / "bad.pfm"), np.zeros((2, 2, 2), np.float32))


def test_read_pfm_matches_reference(kat, tmp_path):
    for name in ("gray", "gray1", "color", "be"):
        fn = str(tmp_path / (name + ".pfm"))
        open(fn, "wb").write(kat[f"pfm_{name}_bytes"].tobytes())
        data, scale = mio.read_pfm(fn)
        assert data.dtype.kind == "f" and data.shape == kat[f"pfm_{name}_read"].shape
        assert np.array_equal(data, kat[f"pfm_{name}_read"]) and scale == float(kat[f"pfm_{name}_scale"])
    # write -> read round trip returns the array ([H,W] comes back as [H,W,1])
    assert np.array_equal(mio.read_pfm(str(tmp_path / "gray.pfm"))[0][..., 0], kat["pfm_gray_in"])
    bad = str(tmp_path / "bad.pfm")
    open(bad, "wb").write(b"P6\n1 1\n-1\n")
    with pytest.raises(Exception):
        mio.read_pfm(bad)


def _rebuild_scan(kat, root):
    os.makedirs(root / "scan1" / "images")
    os.makedirs(root / "scan1" / "cams_1")
    for v in range(3):
        open(root / "scan1" / "images" / ("%08d.jpg" % v), "wb").write(kat[f"jpg{v}"].tobytes())
        open(root / "scan1" / "cams_1" / ("%08d_cam.txt" % v), "wb").write(kat[f"cam{v}"].tobytes())
    open(root / "scan1" / "pair.txt", "wb").write(kat["pair"].tobytes())


def test_cam_and_pair_files(kat, tmp_path):
    _rebuild_scan(kat, tmp_path)
    k, e, dmin, dmax = mio.read_cam_file(str(tmp_path / "scan1" / "cams_1" / "00000001_cam.txt"))
    assert k.dtype == np.float32 and e.dtype == np.float32
    assert np.array_equal(k, kat["cam1_K"]) and np.array_equal(e, kat["cam1_E"])
    assert [dmin, dmax] == list(kat["cam1_range"])
    pairs = mio.read_pair_file(str(tmp_path / "scan1" / "pair.txt"))
    assert [p[0] for p in pairs] == list(kat["pairs_ref"]) and [p[1] for p in pairs] == [list(r) for r in kat["pairs_src"]]
    assert len(pairs) == 2            # the view without source views is dropped (eval.py:98)


def test_load_views_matches_reference_loader(kat, tmp_path):
    pytest.importorskip("cv2")
    pytest.importorskip("PIL")
    _rebuild_scan(kat, tmp_path)
    s = mio.load_views(str(tmp_path), "scan1", 0, [1, 2], nviews=3, img_wh=(64, 32))
    for lv in ("level_0", "level_1", "level_2", "level_3"):
        assert s["imgs"][lv].dtype == np.float32 and s["imgs"][lv].shape == kat[f"imgs_{lv}"].shape
        assert np.array_equal(s["imgs"][lv], kat[f"imgs_{lv}"]), lv
        assert s["proj_matrices"][lv].dtype == np.float32 and np.array_equal(s["proj_matrices"][lv], kat[f"proj_{lv}"]), lv
    assert s["depth_min"] == float(kat["depth_min"]) and s["depth_max"] == float(kat["depth_max"])
    assert s["filename"] == kat["filename"].tobytes().decode()


def test_point_cloud_tail_matches_reference_lines(tmp_path):
    """backproject_points == eval.py:281-296 executed verbatim (tests/golden/make_golden_points.py); save_ply / read_ply
    round trip with the float32 / uint8 casts eval.py:300-301 applies before plyfile writes."""
    with np.load(os.path.join(os.path.dirname(__file__), "golden", "points_kat.npz")) as z:
        k = {n: z[n] for n in z.files}
    v, c = mio.backproject_points(k["depth_est_averaged"], k["final_mask"], k["ref_img"], k["ref_intrinsics"], k["ref_extrinsics"])
    assert v.dtype == k["vertices"].dtype and np.array_equal(v, k["vertices"])
    assert c.dtype == np.uint8 and np.array_equal(c, k["colors"])
    fn = str(tmp_path / "fused.ply")
    mio.save_ply(fn, v, c)
    raw = open(fn, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\n" % len(v))
    assert len(body) == 15 * len(v)                                  # 3 x float32 + 3 x uchar per point, packed
    v2, c2 = mio.read_ply(fn)
    assert np.array_equal(v2, v.astype(np.float32)) and np.array_equal(c2, c)
    with pytest.raises(ValueError):
        mio.save_ply(fn, v[:, :2], c)
