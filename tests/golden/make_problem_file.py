#!/usr/bin/env python
"""Golden problem file: the bytes the UNMODIFIED reference's test_reference()
(test/test.c:108-130) writes for a seeded problem, captured through oracle/_ref/
libqpcompat.so (the reference built with its own config.h, N_DIM = 48).

    python tests/golden/make_problem_file.py     (needs /root/reference; run in the build container)

Writes tests/golden/problem_file_n48.bin (the captured file) and problem_file_n48.npz
(the P and q the reference held in memory when it wrote it).
"""
import ctypes as C
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

STUB = "import shutil, sys\nshutil.copy(sys.argv[1], 'captured.bin')\n"


def capture(seed: int):
    """Returns (bytes the reference wrote, P, q)."""
    import oracle
    oracle.build()
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libqpcompat.so"))
    lib.compat_dim.restype = C.c_uint
    n = lib.compat_dim()
    d = C.POINTER(C.c_double)
    lib.compat_write_reference_file.restype = C.c_int
    lib.compat_write_reference_file.argtypes = [C.c_uint, d, d]
    P, q = np.zeros((n, n)), np.zeros(n)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    try:
        os.makedirs(os.path.join(tmp, "test"))
        with open(os.path.join(tmp, "test", "qp_ref.py"), "w") as f:
            f.write(STUB)  # stands in for the reference's script: keeps the file it is given
        os.chdir(tmp)
        rc = lib.compat_write_reference_file(seed, P.ctypes.data_as(d), q.ctypes.data_as(d))
        assert rc == 0
        with open("captured.bin", "rb") as f:
            data = f.read()
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    return data, P, q


if __name__ == "__main__":
    data, P, q = capture(2026)
    with open(os.path.join(HERE, "problem_file_n48.bin"), "wb") as f:
        f.write(data)
    np.savez_compressed(os.path.join(HERE, "problem_file_n48.npz"), P=P, q=q)
    print(f"captured {len(data)} bytes, n = {P.shape[0]}")
