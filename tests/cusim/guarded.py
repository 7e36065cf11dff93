"""Guard-page buffers for the emulated kernels (TEST INFRASTRUCTURE): a tensor whose last (or first) byte abuts a
PROT_NONE page, so that a kernel that reads or writes one element past either end of a buffer dies with SIGSEGV
instead of silently reading a neighbour -- a poor man's compute-sanitizer memcheck for the CPU suite."""
import ctypes
import mmap

import numpy as np
import torch

_libc = ctypes.CDLL(None, use_errno=True)
_PAGE = mmap.PAGESIZE
_keep = []


def guarded(t: torch.Tensor, side: str = "end") -> torch.Tensor:
    """Copy `t` (float32 / int32, contiguous) into a fresh mapping with an inaccessible page right after its last byte
    (side='end') or right before its first byte (side='start')."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    assert nbytes % 16 == 0, "guarded(): sizes must keep the 16-byte alignment the kernels require"
    body = (nbytes + _PAGE - 1) // _PAGE * _PAGE
    mm = mmap.mmap(-1, body + 2 * _PAGE)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    for guard in (base, base + _PAGE + body):
        if _libc.mprotect(ctypes.c_void_p(guard), ctypes.c_size_t(_PAGE), 0) != 0:
            raise OSError(ctypes.get_errno(), "mprotect failed")
    off = _PAGE + (body - nbytes if side == "end" else 0)
    arr = np.frombuffer(mm, dtype=np.uint8, count=nbytes, offset=off)
    out = torch.from_numpy(arr).view(t.dtype).view(t.shape)
    out.copy_(t)
    _keep.append(mm)
    return out
