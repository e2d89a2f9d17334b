"""Build libqpb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m embedded_qp_solver_b200.build [--verbose]

One translation unit per kernel family, compiled in parallel into build/*.o (only the ones
whose sources changed), then linked.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "libqpb200.so")
PUBLIC = os.path.join(os.path.dirname(HERE), "include", "qp_batch.h")
COMMON = ["qp_common.cuh", "qp_internal.h", "qp_aux_args.h"]
# translation unit -> headers it includes besides COMMON
UNITS = {
    "k_tpq.cu": [],
    "k_tpq4.cu": ["qp_tpq.cuh", "k_tpq_impl.cuh"],
    "k_tpq8.cu": ["qp_tpq.cuh", "k_tpq_impl.cuh"],
    "k_tpq12.cu": ["qp_tpq.cuh", "k_tpq_impl.cuh"],
    "k_tpq16.cu": ["qp_tpq.cuh", "k_tpq_impl.cuh"],
    "k_bpq.cu": ["qp_bpq.cuh", "qp_gpq.cuh"],
    "k_bpqs.cu": ["qp_bpq_stream.cuh", "qp_bpq.cuh", "qp_gpq.cuh"],
    "k_gpq.cu": ["qp_gpq.cuh"],
    "k_cpq.cu": ["qp_cpq.cuh"],
    "k_aux.cu": ["qp_gen.cuh", "qp_newton.cuh"],
    "k_genr.cu": ["qp_genr.cuh"],
    "k_f64.cu": ["qp_f64.cuh"],
    "qp_batch.cu": [],
    "qp_file.cpp": [],
    "qp_hostpack.cpp": [],
}
SOURCES = [os.path.join(CSRC, u) for u in UNITS]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=default",
]


def _deps(unit: str):
    return [os.path.join(CSRC, unit), PUBLIC, __file__] + \
        [os.path.join(CSRC, h) for h in COMMON + UNITS[unit]]


def _obj(unit: str) -> str:
    return os.path.join(OBJ, os.path.splitext(unit)[0] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return any(_stale(LIB, _deps(u)) for u in UNITS)


def build(verbose: bool = False, force: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    extra = os.environ.get("QPB_NVCC_EXTRA", "").split()
    stamp = os.path.join(OBJ, ".flags")
    flags_now = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(stamp) or open(stamp).read() != flags_now:
        force = True

    def compile_unit(unit):
        if not force and not _stale(_obj(unit), _deps(unit)):
            return unit, 0, ""
        cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", "-o", _obj(unit), os.path.join(CSRC, unit)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return unit, res.returncode, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_unit, UNITS))
    failed = [u for u, rc, _ in results if rc != 0]
    for u, rc, out in results:
        if out and (verbose or rc != 0):
            sys.stderr.write(f"---- {u} ----\n{out}")
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
           "-o", LIB] + [_obj(u) for u in UNITS]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed linking libqpb200.so")
    with open(stamp, "w") as f:
        f.write(flags_now)
    return LIB


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv, force=True))
