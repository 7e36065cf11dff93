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
ABI_VERSION = 5
FNET_CONVS = 24


class WPair(C.Structure):
    _fields_ = [("tf32", vp), ("fp32", vp), ("umma", vp), ("f16x3", vp), ("f16umma", vp), ("f16ummai", vp)]


class CorrNetWeights(C.Structure):
    _fields_ = [(n, WPair) for n in ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5")] + [("conv5_b", vp)]


class Weights(C.Structure):
    _fields_ = [("pvw_conv0", WPair), ("pvw_conv1", vp), ("pvw_conv1_b", vp),
                ("corrnet", CorrNetWeights * 3),
                ("gru_zr", WPair), ("gru_zr_b", vp), ("gru_q", WPair), ("gru_q_b", vp),
                ("head_conv0", WPair), ("head_fc1", WPair), ("head_fc2", WPair), ("head_fc2_b", vp),
                ("conf_fc", vp), ("conf_fc_b", vp),
                ("hinit_conv0", WPair), ("hinit_fc", WPair), ("hinit_fc_b", vp),
                ("ups_conv0", WPair), ("ups_fc", vp), ("head_fused", vp)]


class Problem(C.Structure):
    _fields_ = [(n, ci) for n in ("B", "V", "H", "W", "D", "iterations")]


class FeatureNetWeights(C.Structure):
    _fields_ = [("w", WPair * FNET_CONVS), ("b", vp * FNET_CONVS)]


PW = C.POINTER(Weights)
_SIGNATURES = {
    "imvs_abi_version": (ci, []),
    "imvs_last_error": (C.c_char_p, []),
    "imvs_launches_total": (C.c_longlong, []),
    "imvs_set_conv_passes": (ci, [ci]),
    "imvs_get_conv_passes": (ci, []),
    "imvs_set_tcgen05": (ci, [ci]),
    "imvs_tcgen05_status": (ci, []),
    "imvs_device_status": (ci, [ci]),
    "imvs_profile_begin": (ci, [ci]),
    "imvs_profile_end": (ci, [vp, vp, ci]),
    "imvs_compose_projections": (ci, [vp, ci, ci, vp, vp, vp]),
    "imvs_differentiable_warping": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]),
    "imvs_differentiable_warping_backward": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]),
    "imvs_nchw_to_nhwc": (ci, [vp, vp, ci, ci, ci, ci, vp]),
    "imvs_nhwc_to_nchw": (ci, [vp, vp, ci, ci, ci, ci, vp]),
    "imvs_warpcorr_init": (ci, [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "imvs_pixel_view_weight": (ci, [PW, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "imvs_aggregate_init": (ci, [vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_warpcorr_iter": (ci, [vp, vp, vp, vp, vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_pad_level3": (ci, [vp, vp, ci, ci, ci, ci, vp]),
    "imvs_warpcorr_init_padded": (ci, [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "imvs_warpcorr_iter_padded": (ci, [vp, vp, vp, vp, vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_warpcorr_init_backward": (ci, [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "imvs_warpcorr_iter_backward": (ci, [vp, vp, vp, vp, vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_corrnet_scratch_floats": (sz, [ci, ci, ci]),
    "imvs_corrnet": (ci, [C.POINTER(CorrNetWeights), ci, ci, ci, vp, vp, sz, sz, vp, ci, ci, ci, vp]),
    "imvs_hidden_init": (ci, [PW, vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_conv_gru": (ci, [PW, vp, vp, vp, ci, ci, ci, vp]),
    "imvs_depth_head": (ci, [PW, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]),
    "imvs_upsample_outputs": (ci, [PW, vp, sz, vp, sz, sz, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]),
    "imvs_check_geometric_consistency": (ci, [vp, vp, vp, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, ci, ci, vp]),
    "imvs_filter_depth_view": (ci, [vp, vp, vp, vp, ci, C.c_float, C.c_float, C.c_float, ci, vp, vp, vp, vp, vp, vp, ci, ci, vp]),
    "imvs_init_depth": (ci, [vp, sz, sz, sz, vp, vp, vp, vp, ci, ci, ci, ci, vp]),
    "imvs_conv3x3_tcgen05_workspace_bytes": (sz, [ci, ci, ci, ci, ci]),
    "imvs_conv3x3_tcgen05": (ci, [vp, vp, vp, vp, vp, vp, sz, ci, ci, ci, ci, ci, ci, ci, ci, vp]),
    "imvs_image_pyramid_u8": (ci, [vp, ci, ci, vp, vp, vp, vp, ci, ci, vp]),
    "imvs_set_sm_share": (ci, [ci]),
    "imvs_forward_workspace_bytes": (sz, [C.POINTER(Problem)]),
    "imvs_forward_launch_count": (ci, [C.POINTER(Problem)]),
    "imvs_itermvs_forward": (ci, [C.POINTER(Problem), PW, vp, vp, vp, vp, vp, vp, vp, vp,
                                  vp, sz, vp, vp, vp, vp, vp, vp]),
    "imvs_featurenet_workspace_bytes": (sz, [ci, ci, ci]),
    "imvs_featurenet_forward": (ci, [C.POINTER(FeatureNetWeights), vp, vp, vp, vp, vp, sz, ci, ci, ci, vp]),
    "imvs_featurenet_forward_u8": (ci, [C.POINTER(FeatureNetWeights), vp, vp, vp, vp, vp, sz, ci, ci, ci, vp]),
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
            if os.path.exists(path):
                import sys
                print(f"itermvs_b200: WARNING: {path} is older than its sources and rebuilding failed ({str(e)[:300]}...); "
                      "loading the existing library", file=sys.stderr)
            if not os.path.exists(path):
                raise LibraryMissing(
                    f"itermvs_b200: CUDA library {path} is missing and could not be built ({e}). "
                    "There is no CPU fallback; run `python -m itermvs_b200._build` where nvcc is available.") from e
        handle = C.CDLL(path)
        handle.imvs_abi_version.restype = ci
        if handle.imvs_abi_version() != ABI_VERSION:        # before binding: a stale library may lack newer symbols
            raise LibraryMissing(f"itermvs_b200: ABI version mismatch in {path} (rebuild with python -m itermvs_b200._build --force)")
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
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


def set_conv_passes(passes: int) -> None:
    """1 = single-pass TF32 tensor-core convolutions, 3 = 3xTF32 error-compensated (fp32-grade, default)."""
    check(lib().imvs_set_conv_passes(int(passes)), "set_conv_passes")


def device_status(clear: bool = True) -> int:
    """Bitmask (synchronises): 1 = tcgen05 mbarrier time-out, 2 = a convolution output left the fp16 range in mode 4."""
    return int(lib().imvs_device_status(1 if clear else 0))


def check_device_status() -> None:
    st = device_status(clear=True)
    if st & 1:
        raise RuntimeError("itermvs_b200: a tcgen05 kernel timed out on its mbarrier")
    if st & 2:
        raise OverflowError("itermvs_b200: a convolution output exceeded +-65504 in the fp32-grade fp16-split mode (mode 4): "
                            "results are invalid; use imvs_set_conv_passes(3) (3xTF32) for inputs of this magnitude")
    if st & 4:
        raise RuntimeError("itermvs_b200: CUDA error while reading the device status word")


def get_conv_passes() -> int:
    return int(lib().imvs_get_conv_passes())
