"""ctypes binding of the C ABI declared in include/itermvs_b200.h.

The product path has NO fallback: if `libitermvs_b200.so` cannot be loaded (and cannot be built
because nvcc is absent) importing an operator raises.  The library is built in-tree by
`itermvs_b200._build` so that it travels with the source tree.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _build

_lock = threading.Lock()
_lib = None

vp = C.c_void_p
ci = C.c_int
sz = C.c_size_t


class CorrNetWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv5_b")]


class Weights(C.Structure):
    _fields_ = ([(n, vp) for n in ("pvw_conv0", "pvw_conv1", "pvw_conv1_b")]
                + [("corrnet", CorrNetWeights * 3)]
                + [(n, vp) for n in ("gru_zr", "gru_zr_b", "gru_q", "gru_q_b",
                                     "head_conv0", "head_fc1", "head_fc2", "head_fc2_b", "conf_fc", "conf_fc_b",
                                     "hinit_conv0", "hinit_fc", "hinit_fc_b", "ups_conv0", "ups_fc")])


class Problem(C.Structure):
    _fields_ = [(n, ci) for n in ("B", "V", "H", "W", "D", "iterations")]


class FeatureNetWeights(C.Structure):
    _fields_ = [("w", vp * 32), ("b", vp * 32)]


_SIGNATURES = {
    "imvs_abi_version": (ci, []),
    "imvs_last_error": (C.c_char_p, []),
    "imvs_launches_total": (C.c_longlong, []),
    "imvs_profile_begin": (ci, [ci]),
    "imvs_profile_end": (ci, [vp, vp, ci]),
    "imvs_compose_projections": (ci, [vp, ci, ci, vp, vp, vp]),
    "imvs_differentiable_warping": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]),
    "imvs_nchw_to_nhwc": (ci, [vp, vp, ci, ci, ci, ci, vp]),
    "imvs_nhwc_to_nchw": (ci, [vp, vp, ci, ci, ci, ci, vp]),
    "imvs_warpcorr_init": (ci, [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "imvs_pixel_view_weight": (ci, [C.POINTER(Weights), vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "imvs_aggregate_init": (ci, [vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_warpcorr_iter": (ci, [vp, vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_corrnet_scratch_floats": (sz, [ci, ci, ci]),
    "imvs_corrnet": (ci, [C.POINTER(CorrNetWeights), ci, ci, ci, vp, vp, sz, vp, ci, ci, ci, vp]),
    "imvs_hidden_init": (ci, [C.POINTER(Weights), vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_conv_gru": (ci, [C.POINTER(Weights), vp, vp, vp, ci, ci, ci, vp]),
    "imvs_depth_head": (ci, [C.POINTER(Weights), vp, vp, sz, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]),
    "imvs_upsample_outputs": (ci, [C.POINTER(Weights), vp, vp, sz, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]),
    "imvs_forward_workspace_bytes": (sz, [C.POINTER(Problem)]),
    "imvs_forward_launch_count": (ci, [C.POINTER(Problem)]),
    "imvs_itermvs_forward": (ci, [C.POINTER(Problem), C.POINTER(Weights), vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                  vp, sz, vp, vp, vp, vp, vp, vp]),
}
# entry points that later revisions add; bound when present
_OPTIONAL = {
    "imvs_featurenet_workspace_bytes": (sz, [ci, ci, ci]),
    "imvs_featurenet_forward": (ci, [C.POINTER(FeatureNetWeights), vp, vp, vp, vp, vp, vp, sz, ci, ci, ci, vp]),
    "imvs_featurenet_launch_count": (ci, []),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class LibraryMissing(RuntimeError):
    pass


def lib():
    """Load (building first if the in-tree .so is missing or stale and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        try:
            if _build.is_stale():
                path = _build.build()
        except Exception as e:  # no nvcc on the box and no prebuilt library
            if not os.path.exists(path):
                raise LibraryMissing(
                    f"itermvs_b200: CUDA library {path} is missing and could not be built ({e}). "
                    "There is no CPU fallback; run `python -m itermvs_b200._build` where nvcc is available.") from e
        handle = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        for name, (res, args) in _OPTIONAL.items():
            if hasattr(handle, name):
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
        if handle.imvs_abi_version() != 1:
            raise LibraryMissing(f"itermvs_b200: ABI version mismatch in {path}")
        _lib = handle
        return _lib


def library_path() -> str:
    return _build.LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().imvs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"itermvs_b200 {what}: {msg}")


def launches_total() -> int:
    return int(lib().imvs_launches_total())
