#!/usr/bin/env python
"""SASS census of every kernel in libqpb200.so: which instructions prove what the docs claim
(UBLKCP / SYNCS = bulk asynchronous copies through the TMA unit and mbarrier waits; FFMA2 =
packed FP32 pairs; LDL / STL = register spills).  Runs where the library was built
(cuobjdump, no GPU):  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "embedded_qp_solver_b200", "lib", "libqpb200.so")
COLS = ["UBLKCP", "SYNCS", "FFMA2", "FFMA", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "LDL", "STL", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)),
                           capture_output=True, text=True).stdout.split("\n")
    kernels, cur = [], None
    it = iter(names)
    for line in out.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels.append((next(it), cur))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for c in COLS[:-1]:
                if op == c or (c in ("LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "LDL", "STL", "SYNCS", "UBLKCP") and op.startswith(c)):
                    cur[c] += 1
    w = max(len(k) for k, _ in kernels)
    w = min(w, 70)
    print(f"{'kernel':{w}s} " + " ".join(f"{c:>7s}" for c in COLS))
    for k, c in sorted(kernels):
        k = re.sub(r"\(qpb::KernelArgs\)|\(qpb::\w+Args\)|void |qpb::", "", k)
        print(f"{k[:w]:{w}s} " + " ".join(f"{c[x]:7d}" for x in COLS))
    spilled = [k for k, c in kernels if c["LDL"] or c["STL"]]
    print(f"\n{len(kernels)} kernels; with local-memory instructions (spills): {len(spilled)}")


if __name__ == "__main__":
    main()
