import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build(ref=False)
    return oracle.Oracle()


def golden_files():
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith("problem_file"))


def load_golden(path):
    d = np.load(path)
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def qp():
    import embedded_qp_solver_b200 as m
    from embedded_qp_solver_b200 import build
    build.build()
    m.load_library()
    return m


def nondegenerate(margins, tol=1e-3):
    """Problems whose stop test was not within `tol` (relative) of flipping at
    the exit iteration or the one before: the set on which FP32 and FP64 must
    agree on iteration count, status and active set."""
    return np.all(np.abs(margins) > tol, axis=1)
