"""Build the CPU simulation of the non-tensor-core kernel sources (TEST INFRASTRUCTURE, never used by the
product path -- see tests/cusim/cuda_runtime.h).

    python tests/cusim/cusim_build.py          # -> tests/cusim/_build/libitermvs_sim.so

The *.cu files are taken from itermvs_b200/csrc as they are; three textual rewrites make them host C++:
  * `extern __shared__ T name[];`  ->  `T* name = (T*)cusim::dyn_smem();`
  * `asm volatile(...)` / `asm(...)` -> `CUSIM_ASM(...)` (only the griddepcontrol hints of common.cuh occur
    in the simulated files; sources with data-path PTX -- mma.sync, tcgen05, cp.async -- are not simulated)
  * `#include <cuda_runtime.h>` resolves to tests/cusim/cuda_runtime.h (include path order).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "itermvs_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libitermvs_sim.so")
SIM_SOURCES = ["warp.cu", "warpcorr.cu", "warpcorr_bwd.cu", "fusion.cu", "evalnets.cu", "update.cu", "upsample.cu", "forward.cu",
               "featurenet.cu", "imageprep.cu"]
HEADERS = ["common.cuh", "sampling.cuh", "mmaconv.cuh", "tc5conv.cuh", "headfused.cuh", "corrnet_tile.cuh"]
def _isa_flags():
    """F16C / FMA when this CPU has them (hardware half<->float conversion for the tensor-core emulation)."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return []
    return [f for f, name in (("-mf16c", " f16c"), ("-mfma", " fma")) if name in flags]


CXXFLAGS = ["-std=c++17", "-O2", *_isa_flags(), "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes",
            "-fno-strict-aliasing"]

_DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?((?:unsigned\s+)?\w+)\s+(\w+)\s*\[\s*\]\s*;")
_ASM = re.compile(r"\basm\s+volatile\s*\(|\basm\s*\(")


def rewrite(text: str) -> str:
    text = _DYN.sub(r"\1* \2 = (\1*)::cusim::dyn_smem();", text)
    return _ASM.sub("CUSIM_ASM(", text)


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    inputs = [os.path.join(CSRC, f) for f in SIM_SOURCES + HEADERS] + \
             [os.path.join(HERE, f) for f in ("cuda_runtime.h", "cuda_fp16.h", "cusim.cpp", "cusim_build.py")] + \
             [os.path.join(ROOT, "include", "itermvs_b200.h")]
    stamp = os.path.join(OUT, "stamp")
    digest = _digest(inputs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    # mirror the relative layout (csrc includes "../../include/itermvs_b200.h")
    gen = os.path.join(OUT, "gen", "itermvs_b200", "csrc")
    os.makedirs(gen, exist_ok=True)
    os.makedirs(os.path.join(OUT, "gen", "include"), exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "itermvs_b200.h"), os.path.join(OUT, "gen", "include", "itermvs_b200.h"))
    for f in SIM_SOURCES + HEADERS:
        with open(os.path.join(CSRC, f)) as src:
            text = rewrite(src.read())
        with open(os.path.join(gen, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w") as dst:
            dst.write(text)
    cpps = [os.path.join(gen, f.replace(".cu", ".cpp")) for f in SIM_SOURCES] + [os.path.join(HERE, "cusim.cpp")]

    def compile_one(src):
        obj = os.path.join(OUT, os.path.basename(src) + ".o")
        r = subprocess.run([gxx, *CXXFLAGS, "-I", HERE, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"cusim build failed on {src}:\n" + r.stdout + r.stderr[-8000:])
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(cpps))) as ex:
        objs = list(ex.map(compile_one, cpps))
    r = subprocess.run([gxx, "-shared", "-o", LIB + ".tmp", *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cusim link failed:\n" + r.stdout + r.stderr[-8000:])
    os.replace(LIB + ".tmp", LIB)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv))
