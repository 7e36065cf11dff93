"""Build the CPU simulation of the non-tensor-core kernel sources (TEST INFRASTRUCTURE, never used by the
product path -- see tests/cusim/cuda_runtime.h).

    python tests/cusim/build.py          # -> tests/cusim/_build/libitermvs_sim.so

The *.cu files are taken from itermvs_b200/csrc as they are; three textual rewrites make them host C++:
  * `extern __shared__ T name[];`  ->  `T* name = (T*)cusim::dyn_smem();`
  * `asm volatile(...)` / `asm(...)` -> `CUSIM_ASM(...)` (only the griddepcontrol hints of common.cuh occur
    in the simulated files; sources with data-path PTX -- mma.sync, tcgen05, cp.async -- are not simulated)
  * `#include <cuda_runtime.h>` resolves to tests/cusim/cuda_runtime.h (include path order).
"""
from __future__ import annotations

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
SIM_SOURCES = ["warp.cu", "warpcorr.cu", "warpcorr_bwd.cu", "fusion.cu"]
HEADERS = ["common.cuh", "sampling.cuh"]
CXXFLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes",
            "-fno-strict-aliasing"]

_DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\s*\[\s*\]\s*;")
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
             [os.path.join(HERE, f) for f in ("cuda_runtime.h", "cusim.cpp", "build.py")] + \
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
    cmd = [gxx, *CXXFLAGS, "-I", HERE, "-shared", "-o", LIB + ".tmp", *cpps]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cusim build failed:\n" + r.stdout + r.stderr[-8000:])
    os.replace(LIB + ".tmp", LIB)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv))
