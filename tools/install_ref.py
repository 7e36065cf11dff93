#!/usr/bin/env python
"""Install the UNMODIFIED reference model package next to the repository so that it travels to the GPU box.

    python tools/install_ref.py            # copy if /root/reference is present, verify otherwise

The reference (FangjinhuaWang/IterMVS) is pure Python with no setup.py, so `pip install --target baseline/_ref`
has nothing to build; this recipe is the install.  It copies, byte for byte,

    /root/reference/models/{__init__,module,itermvs,net}.py   -> baseline/_ref/models/
    /root/reference/checkpoints/dtu/model_000015.ckpt          -> baseline/_ref/checkpoints/dtu/

into the git-ignored (NOT gpurun-ignored) directory `baseline/_ref/` and writes `MANIFEST.json` with the sha256 of
every file.  Nothing under `baseline/_ref/` is tracked, and nothing in `itermvs_b200/` imports it: it is the timing
baseline (`bench.py --impl reference`, the `gpu_stock_ref` leg of the default bench) and a parity witness
(`tests/test_gpu_parity.py::test_reference_itself_on_this_gpu`), i.e. the reference's own stock PyTorch/cuDNN path
(SURVEY 8c "GPU oracle", BASELINE.md section 2 B-CPU / B-GPU).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/__init__.py", "models/module.py", "models/itermvs.py", "models/net.py",
         "checkpoints/dtu/model_000015.ckpt"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def installed() -> bool:
    man = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(man):
        return False
    try:
        m = json.load(open(man))
    except Exception:
        return False
    return all(os.path.exists(os.path.join(DST, f)) and _sha(os.path.join(DST, f)) == h for f, h in m["sha256"].items())


def install(verbose: bool = True) -> str:
    """Returns 'installed', 'present' (already there and intact) or 'unavailable' (no /root/reference and no copy)."""
    if not os.path.isdir(REF):
        state = "present" if installed() else "unavailable"
        if verbose:
            print(f"install_ref: {REF} absent; baseline/_ref {state}")
        return state
    sums = {}
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and _sha(dst) == _sha(src)):
            shutil.copyfile(src, dst)
        sums[f] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF, "what": "unmodified copies, see tools/install_ref.py", "sha256": sums}, fh, indent=1)
    if verbose:
        print(f"install_ref: {len(FILES)} files -> {DST}")
    return "installed"


if __name__ == "__main__":
    s = install()
    sys.exit(0 if s != "unavailable" else 1)
