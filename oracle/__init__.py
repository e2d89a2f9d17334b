"""oracle/ -- TEST INFRASTRUCTURE, not product code.

ctypes loaders for
  * ``liboracle.so``            our CPU restatement of the reference's ADMM path
                                (oracle/admm_oracle.c), and
  * ``_ref/libqpref_n<N>.so``   the UNMODIFIED reference compiled per shape
                                (oracle/Makefile, oracle/shim/).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``embedded_qp_solver_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"

# reference constants, qp_solvers/qp_solvers.c:17-18 and :281-282
REF_RHO = 1.0
REF_ALPHA = 1.0
REF_ABSTOL = 1e-4
REF_RELTOL = 1e-2
REF_MAX_ITER = 10000  # ADMM_ITERATIONS, config.h:38

STATUS_CONVERGED = 0
STATUS_MAX_ITER = 1
STATUS_SINGULAR = 2
STATUS_NONFINITE = 3


def build(ref: bool | None = None) -> None:
    """Compile liboracle.so and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref is None:
        ref = os.path.isdir(REFERENCE_ROOT)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref", "reffast"], check=True)
        if os.path.exists(os.path.join(HERE, "..", "embedded_qp_solver_b200", "lib",
                                       "libqpb200.so")):
            subprocess.run(["make", "-s", "-C", HERE, "compat"], check=True)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


class Oracle:
    """The restatement.  dtype float64 is the pinned oracle; float32 is its
    same-operation-order twin used to separate precision effects from bugs."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        for suf, ct in (("f64", C.c_double), ("f32", C.c_float)):
            p = C.POINTER(ct)
            f = getattr(self.lib, f"qp_oracle_admm_batch_{suf}")
            f.restype = None
            f.argtypes = [C.c_uint, C.c_ulong, p, p, p, p, ct, ct, ct, ct,
                          C.c_uint, p, p, p, p, p, C.POINTER(C.c_int),
                          C.POINTER(C.c_int), p]
            f = getattr(self.lib, f"qp_oracle_invert_{suf}")
            f.restype = C.c_int
            f.argtypes = [p, C.c_uint]
            f = getattr(self.lib, f"qp_oracle_eval_{suf}")
            f.restype = ct
            f.argtypes = [p, p, ct, p, C.c_uint]
            for name in ("newton", "gradient"):
                f = getattr(self.lib, f"qp_oracle_{name}_{suf}")
                f.restype = C.c_int
                f.argtypes = [C.c_uint, p, p, p, C.c_uint, p]
        f = self.lib.qp_oracle_admm_timed_f64
        f.restype = C.c_double
        f.argtypes = [C.c_uint, C.c_ulong] + [C.POINTER(C.c_double)] * 4 + \
            [C.c_uint, C.POINTER(C.c_double)]

    @staticmethod
    def _suf(dtype):
        dtype = np.dtype(dtype)
        if dtype == np.float64:
            return "f64", C.c_double
        if dtype == np.float32:
            return "f32", C.c_float
        raise TypeError(dtype)

    def admm(self, P, q, lb, ub, rho=REF_RHO, alpha=REF_ALPHA,
             abstol=REF_ABSTOL, reltol=REF_RELTOL, max_iter=REF_MAX_ITER,
             z0=None, u0=None, dtype=np.float64):
        """Solve a packed batch.  P [count,n,n], q/lb/ub [count,n] (lb/ub may
        be scalars).  Returns dict(x,z,u,lam,iters,status,margins)."""
        suf, ct = self._suf(dtype)
        P = np.ascontiguousarray(P, dtype=dtype)
        count, n, _ = P.shape
        q = np.ascontiguousarray(q, dtype=dtype).reshape(count, n)
        lb = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=dtype), (count, n)))
        ub = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=dtype), (count, n)))
        if z0 is not None:
            z0 = np.ascontiguousarray(z0, dtype=dtype).reshape(count, n)
        if u0 is not None:
            u0 = np.ascontiguousarray(u0, dtype=dtype).reshape(count, n)
        x = np.zeros((count, n), dtype)
        z = np.zeros((count, n), dtype)
        u = np.zeros((count, n), dtype)
        iters = np.zeros(count, np.int32)
        status = np.zeros(count, np.int32)
        margins = np.zeros((count, 4), dtype)
        getattr(self.lib, f"qp_oracle_admm_batch_{suf}")(
            n, count, _ptr(P, ct), _ptr(q, ct), _ptr(lb, ct), _ptr(ub, ct),
            ct(rho), ct(alpha), ct(abstol), ct(reltol), int(max_iter),
            _ptr(z0, ct), _ptr(u0, ct), _ptr(x, ct), _ptr(z, ct), _ptr(u, ct),
            _ptr(iters, C.c_int), _ptr(status, C.c_int), _ptr(margins, ct))
        return dict(x=x, z=z, u=u, lam=(np.asarray(rho, dtype) * u), iters=iters,
                    status=status, margins=margins)

    def invert(self, A, dtype=np.float64):
        suf, ct = self._suf(dtype)
        A = np.array(A, dtype=dtype, order="C")
        s = getattr(self.lib, f"qp_oracle_invert_{suf}")(_ptr(A, ct), A.shape[0])
        return A, s

    def eval(self, P, q, r, x, dtype=np.float64):
        suf, ct = self._suf(dtype)
        P = np.ascontiguousarray(P, dtype=dtype)
        q = np.ascontiguousarray(q, dtype=dtype)
        x = np.ascontiguousarray(x, dtype=dtype)
        return float(getattr(self.lib, f"qp_oracle_eval_{suf}")(
            _ptr(P, ct), _ptr(q, ct), ct(r), _ptr(x, ct), P.shape[0]))

    def newton(self, P, q, x0, max_iter=10, dtype=np.float64):
        suf, ct = self._suf(dtype)
        P = np.ascontiguousarray(P, dtype=dtype)
        q = np.ascontiguousarray(q, dtype=dtype)
        x0 = np.ascontiguousarray(x0, dtype=dtype)
        x = np.zeros_like(q)
        steps = getattr(self.lib, f"qp_oracle_newton_{suf}")(
            P.shape[0], _ptr(P, ct), _ptr(q, ct), _ptr(x0, ct), int(max_iter), _ptr(x, ct))
        return x, steps

    def gradient(self, P, q, x0, max_iter=10000, dtype=np.float64):
        """gradient_descent_with_line_search (qp_solvers.c:65-101); returns (x, steps)."""
        suf, ct = self._suf(dtype)
        P = np.ascontiguousarray(P, dtype=dtype)
        q = np.ascontiguousarray(q, dtype=dtype)
        x0 = np.ascontiguousarray(x0, dtype=dtype)
        x = np.zeros_like(q)
        steps = getattr(self.lib, f"qp_oracle_gradient_{suf}")(
            P.shape[0], _ptr(P, ct), _ptr(q, ct), _ptr(x0, ct), int(max_iter), _ptr(x, ct))
        return x, steps

    def admm_timed(self, P, q, lb, ub, max_iter=REF_MAX_ITER):
        """Seconds for solving the packed f64 batch with the port; x returned."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        count, n, _ = P.shape
        q = np.ascontiguousarray(q, dtype=np.float64)
        lb = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=np.float64), (count, n)))
        ub = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=np.float64), (count, n)))
        x = np.zeros((count, n))
        d = C.c_double
        secs = self.lib.qp_oracle_admm_timed_f64(
            n, count, _ptr(P, d), _ptr(q, d), _ptr(lb, d), _ptr(ub, d), int(max_iter), _ptr(x, d))
        return secs, x


def ref_available(n: int, fast: bool = False) -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libqpref_n{n}{'_fast' if fast else ''}.so"))


# compiler flags of the two builds of the unmodified reference (oracle/Makefile)
REF_FLAGS = "gcc -O3 -ffp-contract=off"                 # parity build, bit-comparable with the port
REF_FAST_FLAGS = "gcc -O3 -march=x86-64-v3"             # timed build (bench.py CPU arm)


class Reference:
    """The unmodified reference for one compile-time dimension N_DIM = n.

    The reference keeps global static pools (kmalloc/kmalloc.c:37-42): one
    instance per process and per n, not thread-safe; parallel use = fork."""

    def __init__(self, n: int, fast: bool = False):
        path = os.path.join(HERE, "_ref", f"libqpref_n{n}{'_fast' if fast else ''}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle ref` where {REFERENCE_ROOT} exists")
        self.n = n
        self.lib = C.CDLL(path)
        d = C.POINTER(C.c_double)
        self.lib.qpref_dim.restype = C.c_uint
        assert self.lib.qpref_dim() == n
        self.lib.qpref_init.restype = None
        self.lib.qpref_admm.restype = C.c_int
        self.lib.qpref_admm.argtypes = [d, d, d, d, C.c_uint, d]
        self.lib.qpref_admm_many.restype = C.c_double
        self.lib.qpref_admm_many.argtypes = [d, d, d, d, C.c_ulong, C.c_uint, d]
        self.lib.qpref_newton.restype = C.c_int
        self.lib.qpref_newton.argtypes = [d, d, d, C.c_uint, d]
        self.lib.qpref_gradient.restype = C.c_int
        self.lib.qpref_gradient.argtypes = [d, d, d, C.c_uint, d]
        self.lib.qpref_eval.restype = C.c_double
        self.lib.qpref_eval.argtypes = [d, d, C.c_double, d]
        self.lib.qpref_invert.restype = C.c_int
        self.lib.qpref_invert.argtypes = [d]
        self.lib.qpref_layout.restype = None
        self.lib.qpref_layout.argtypes = [C.POINTER(C.c_ulong)]
        self.lib.qpref_selftests.restype = None
        self.lib.qpref_selftests.argtypes = [C.c_uint]
        self.lib.qpref_init()

    def admm_many(self, P, q, lb, ub, max_iter=REF_MAX_ITER):
        """Returns (x [count,n] float64, seconds inside the reference's admm)."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        count, n, _ = P.shape
        assert n == self.n
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(count, n)
        lb = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=np.float64), (count, n)))
        ub = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=np.float64), (count, n)))
        x = np.zeros((count, n))
        d = C.c_double
        secs = self.lib.qpref_admm_many(_ptr(P, d), _ptr(q, d), _ptr(lb, d), _ptr(ub, d),
                                        count, int(max_iter), _ptr(x, d))
        return x, secs

    def newton(self, P, q, x0, max_iter=10):
        d = C.c_double
        P = np.ascontiguousarray(P, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        x = np.zeros(self.n)
        rc = self.lib.qpref_newton(_ptr(P, d), _ptr(q, d), _ptr(x0, d), int(max_iter), _ptr(x, d))
        assert rc == 0
        return x

    def gradient(self, P, q, x0, max_iter=10000):
        d = C.c_double
        P = np.ascontiguousarray(P, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        x = np.zeros(self.n)
        rc = self.lib.qpref_gradient(_ptr(P, d), _ptr(q, d), _ptr(x0, d), int(max_iter), _ptr(x, d))
        assert rc == 0
        return x

    def eval(self, P, q, r, x):
        d = C.c_double
        P = np.ascontiguousarray(P, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        return float(self.lib.qpref_eval(_ptr(P, d), _ptr(q, d), float(r), _ptr(x, d)))

    def invert(self, A):
        A = np.array(A, dtype=np.float64, order="C")
        rc = self.lib.qpref_invert(_ptr(A, C.c_double))
        assert rc == 0
        return A

    def layout(self):
        out = (C.c_ulong * 6)()
        self.lib.qpref_layout(out)
        return list(out)

    def selftests(self, seed=1):
        self.lib.qpref_selftests(seed)
