"""Build libqpb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m embedded_qp_solver_b200.build [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libqpb200.so")
SOURCES = [os.path.join(CSRC, "qp_batch.cu"), os.path.join(CSRC, "qp_file.cpp"),
           os.path.join(CSRC, "qp_hostpack.cpp")]
HEADERS = [os.path.join(CSRC, f) for f in ("qp_common.cuh", "qp_tpq.cuh", "qp_gpq.cuh", "qp_bpq.cuh", "qp_cpq.cuh", "qp_newton.cuh", "qp_gen.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "qp_batch.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=default",
    "-shared", "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [__file__])


def build(verbose: bool = False, force: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    extra = os.environ.get("QPB_NVCC_EXTRA", "").split()
    cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libqpb200.so")
    return LIB


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv, force=True))
