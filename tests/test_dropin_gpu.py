"""GPU: the reference's own program structure with the drop-in solver swapped in.

oracle/_ref/libqpcompat.so is the UNMODIFIED reference built with ITS OWN config.h
(N_DIM = 48, box +-1e12) plus embedded_qp_solver_b200/csrc/compat/admm_b200.c, linked
against libqpb200.so (oracle/Makefile target `compat`).  compat_run() draws a problem
with the reference's generators (matirx_random_pos_def, matrix_random; main.c:36-40),
calls the reference's test_optimizer() through the solver function pointer
(test/test.c:89-106) with `admm` and with `admm_b200`, and returns both solutions.
"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libqpcompat.so")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(LIB), reason="libqpcompat.so not built")]


def test_reference_harness_runs_with_admm_b200(qp, orc, capfd):
    qp.load_library()
    lib = C.CDLL(LIB)
    lib.compat_dim.restype = C.c_uint
    n = lib.compat_dim()
    assert n == 48  # the reference's own config.h:5
    d = C.POINTER(C.c_double)
    lib.compat_run.restype = C.c_int
    lib.compat_run.argtypes = [C.c_uint, d, d, d, d, d]
    for seed in (1, 2, 3):
        x_ref, x_gpu = np.zeros(n), np.zeros(n)
        f = np.zeros(2)
        P, q = np.zeros((n, n)), np.zeros(n)
        rc = lib.compat_run(seed, x_ref.ctypes.data_as(d), x_gpu.ctypes.data_as(d),
                            f.ctypes.data_as(d), P.ctypes.data_as(d), q.ctypes.data_as(d))
        assert rc == 0
        # the comparison the reference makes by eye: printed objectives agree
        # (test/test.c:100-102 prints %e, 7 significant digits)
        assert abs(f[1] - f[0]) <= 1e-3 * abs(f[0]), (seed, f)
        # and the restatement agrees with what the reference's admm returned here
        o = orc.admm(P[None], q[None], -1e12, 1e12)
        assert np.array_equal(o["x"][0], x_ref)
        assert np.all(np.isfinite(x_gpu))
    out = capfd.readouterr().out
    assert out.count("admm_b200 gave:") == 3 and out.count("admm gave:") == 3


def test_binary64_drop_in_reproduces_the_reference(qp, capfd):
    """admm_b200_f64 (qp_options.precision = QP_PRECISION_F64) inside the reference's harness, on
    problems drawn by the reference's own generators: the reference's x to 1e-9, i.e. the same
    iteration count (a binary32 run stops an iteration early or late on one such problem in three)."""
    qp.load_library()
    lib = C.CDLL(LIB)
    lib.compat_dim.restype = C.c_uint
    n = lib.compat_dim()
    d = C.POINTER(C.c_double)
    lib.compat_run_f64.restype = C.c_int
    lib.compat_run_f64.argtypes = [C.c_uint, d, d]
    for seed in range(1, 9):
        x_ref, x_gpu = np.zeros(n), np.zeros(n)
        assert lib.compat_run_f64(seed, x_ref.ctypes.data_as(d), x_gpu.ctypes.data_as(d)) == 0
        err = np.max(np.abs(x_gpu - x_ref)) / (1 + np.max(np.abs(x_ref)))
        assert err <= 1e-9, (seed, err)
    assert capfd.readouterr().out.count("admm_b200_f64 gave:") == 8


def test_reference_harness_runs_with_newton_b200(qp, capfd):
    """The reference's second solver through the same function pointer
    (main.c:46-48): newton_method_with_line_search vs the drop-in newton_b200."""
    qp.load_library()
    lib = C.CDLL(LIB)
    lib.compat_dim.restype = C.c_uint
    n = lib.compat_dim()
    d = C.POINTER(C.c_double)
    lib.compat_run_newton.restype = C.c_int
    lib.compat_run_newton.argtypes = [C.c_uint, d, d, d]
    for seed in (4, 5, 6):
        x_ref, x_gpu, f = np.zeros(n), np.zeros(n), np.zeros(2)
        rc = lib.compat_run_newton(seed, x_ref.ctypes.data_as(d), x_gpu.ctypes.data_as(d),
                                   f.ctypes.data_as(d))
        assert rc == 0
        assert abs(f[1] - f[0]) <= 1e-3 * abs(f[0]), (seed, f)
        assert np.all(np.isfinite(x_gpu))
    out = capfd.readouterr().out
    assert out.count("newton_b200 gave:") == 3 and out.count("newton gave:") == 3


def test_reference_harness_runs_with_gradient_b200(qp, capfd):
    """The reference's first solver through the same function pointer (main.c:42-44):
    gradient_descent_with_line_search vs the drop-in gradient_b200.  The reference's own
    test distribution is ill-conditioned (cond ~ 1e4), so both run into GRAD_ITERATIONS = 1e4
    without meeting the gradient test, 1e4 discrete step-size decisions apart: the printed
    objectives (what the reference compares by eye) agree to a few per cent, not to digits
    (tests/test_newton_gpu.py holds the tight comparison on problems that converge)."""
    qp.load_library()
    lib = C.CDLL(LIB)
    lib.compat_dim.restype = C.c_uint
    n = lib.compat_dim()
    d = C.POINTER(C.c_double)
    lib.compat_run_gradient.restype = C.c_int
    lib.compat_run_gradient.argtypes = [C.c_uint, d, d, d]
    for seed in (7, 8):
        x_ref, x_gpu, f = np.zeros(n), np.zeros(n), np.zeros(2)
        rc = lib.compat_run_gradient(seed, x_ref.ctypes.data_as(d), x_gpu.ctypes.data_as(d),
                                     f.ctypes.data_as(d))
        assert rc == 0
        assert abs(f[1] - f[0]) <= 5e-2 * abs(f[0]), (seed, f)
        assert np.all(np.isfinite(x_gpu))
    out = capfd.readouterr().out
    assert out.count("gradient_b200 gave:") == 2 and out.count("gradient gave:") == 2
