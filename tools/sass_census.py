#!/usr/bin/env python
"""SASS census of the shipped library: which kernels contain the Blackwell-native instructions.

    python tools/sass_census.py > profiles/sass_census_r02.txt

Counts, per kernel of itermvs_b200/csrc/libitermvs_b200.so (cuobjdump -sass): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld /
.st), UBLKCP / UTMALDG (bulk / tensor TMA copies), SYNCS (mbarrier), HMMA (legacy mma.sync), LDGSTS (cp.async), REDG (vector
atomics of the backward kernels)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "itermvs_b200", "csrc", "libitermvs_b200.so")
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UBLKCP", r"\bUBLKCP\b"),
                               ("UTMALDG", r"\bUTMALDG\b"), ("SYNCS", r"\bSYNCS\b"), ("UTCBAR", r"\bUTCBAR\b"), ("HMMA", r"\bHMMA\b"),
                               ("LDGSTS", r"\bLDGSTS\b"), ("REDG", r"\bREDG\b")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, p in PAT.items():
            if re.search(p, line):
                kernels[cur][k] += 1
    try:
        names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    except Exception:
        names = list(kernels)
    print("# cuobjdump -sass itermvs_b200/csrc/libitermvs_b200.so (sm_100a): static instruction counts per kernel")
    print("# " + "  ".join(f"{k:>8s}" for k in PAT) + "  kernel")
    tot = collections.Counter()
    for (mangled, c), name in zip(kernels.items(), names):
        tot.update(c)
        if not any(c[k] for k in PAT):
            continue
        name = re.sub(r"\(.*", "", name).replace("imvs::", "")
        print("  " + "  ".join(f"{c[k]:8d}" for k in PAT) + "  " + name[:150])
    print("# total: " + ", ".join(f"{k} {tot[k]}" for k in PAT) + f"; {len(kernels)} kernels")


if __name__ == "__main__":
    main()
