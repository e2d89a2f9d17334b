"""Host-side mirror of the C ABI (include/qp_batch.h) over ctypes.

The reference is a C library; its hosts call ``admm(x0, iterations, qf)``
(/root/reference/qp_solvers/qp_solvers.h:14-16).  This module exposes the same
operator (``admm``/``QuadraticForm``) plus the batched entry points, all of
them thin calls into ``libqpb200.so``.  PyTorch is used only as the owner of
device memory and streams.  There is no CPU fallback: if the CUDA library is
missing or no GPU is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libqpb200.so")

QP_MAX_DIM = 64        # Newton, gradient, general rows
QP_MAX_DIM_BOX = 128   # box ADMM (solve_batch, admm)
QP_MAX_ROWS = 128
STATUS_CONVERGED, STATUS_MAX_ITER, STATUS_SINGULAR, STATUS_NONFINITE = 0, 1, 2, 3
QP_OK, QP_ERR_INVALID, QP_ERR_UNSUPPORTED, QP_ERR_CUDA, QP_ERR_NO_DEVICE = 0, -1, -2, -3, -4
QP_ERR_IO = -5
MEM_HOST, MEM_DEVICE = 0, 1
KERNEL_AUTO, KERNEL_TPQ, KERNEL_GPQ, KERNEL_BPQ, KERNEL_CPQ = 0, 1, 2, 3, 4
KERNEL_GEN_SMEM, KERNEL_GEN_REG = 5, 6  # solve_general_batch only
SCHEDULE_AUTO, SCHEDULE_SINGLE, SCHEDULE_PASSES = 0, 1, 2
P_DENSE, P_UPPER = 0, 1

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


PRECISION_F32, PRECISION_F64 = 0, 1


class qp_batch(C.Structure):
    _fields_ = [("n", C.c_uint32), ("mem", C.c_uint32), ("count", C.c_size_t),
                ("P", C.c_void_p), ("q", C.c_void_p), ("lb", C.c_void_p), ("ub", C.c_void_p),
                ("lb_scalar", C.c_float), ("ub_scalar", C.c_float),
                ("z0", C.c_void_p), ("u0", C.c_void_p), ("rho", C.c_void_p),
                ("p_layout", C.c_uint32), ("p_share", C.c_uint32)]


class qp_batch_solution(C.Structure):
    _fields_ = [("x", C.c_void_p), ("z", C.c_void_p), ("lam", C.c_void_p),
                ("status", C.c_void_p), ("iters", C.c_void_p)]


class qp_general_batch(C.Structure):
    _fields_ = [("n", C.c_uint32), ("m", C.c_uint32), ("mem", C.c_uint32), ("pad", C.c_uint32),
                ("count", C.c_size_t), ("P", C.c_void_p), ("q", C.c_void_p), ("A", C.c_void_p),
                ("l", C.c_void_p), ("u", C.c_void_p), ("z0", C.c_void_p), ("u0", C.c_void_p)]


class qp_general_solution(C.Structure):
    _fields_ = [("x", C.c_void_p), ("z", C.c_void_p), ("y", C.c_void_p),
                ("status", C.c_void_p), ("iters", C.c_void_p)]


class qp_options(C.Structure):
    _fields_ = [("rho", C.c_float), ("alpha", C.c_float), ("abstol", C.c_float),
                ("reltol", C.c_float), ("max_iter", C.c_uint32), ("device_mask", C.c_uint32),
                ("device", C.c_int32), ("stream", C.c_void_p), ("chunk", C.c_uint32),
                ("kernel", C.c_uint32), ("schedule", C.c_uint32), ("precision", C.c_uint32)]


class qp_ref_matrix(C.Structure):
    """Layout of the reference's ``struct _matrix`` (matrix_type.h:20-23)."""
    _fields_ = [("dimensions", C.c_uint), ("elements", C.POINTER(C.c_double))]


class qp_ref_quadratic_form(C.Structure):
    """Layout of the reference's ``struct _quadratic_form`` (qp.h:8-13)."""
    _fields_ = [("p", C.POINTER(qp_ref_matrix)), ("q", C.POINTER(qp_ref_matrix)),
                ("r", C.c_double)]


# every symbol include/qp_batch.h declares (tests check the library exports them all)
ABI_SYMBOLS = [
    "qp_options_default", "qp_solve_batch", "qp_workspace_create", "qp_workspace_destroy",
    "qp_solve_batch_ws", "qp_eval_batch", "qp_admm_qf", "qp_admm_qf_batch", "qp_last_error",
    "qp_device_count", "qp_kernel_for", "qp_probe_fp32_tflops", "qp_probe_hbm_gbs",
    "qp_newton_batch", "qp_newton_qf",
    "qp_problem_file_write", "qp_problem_file_count", "qp_problem_file_read",
    "qp_solve_general_batch", "qp_host_path_packs", "qp_host_path_h2d_bytes", "qp_host_pack_upper", "qp_gradient_batch", "qp_gradient_qf",
    "qp_packed_stride", "qp_probe_copy_seconds", "qp_schedule_stats",
]

_lib = None


class QpError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what} failed with code {code}: {_describe(code)}")
        self.code = code


def _describe(code: int) -> str:
    names = {QP_ERR_INVALID: "invalid argument", QP_ERR_UNSUPPORTED: "unsupported dimension",
             QP_ERR_CUDA: "CUDA error", QP_ERR_NO_DEVICE: "no usable GPU",
             QP_ERR_IO: "problem file I/O"}
    s = names.get(code, "?")
    if code == QP_ERR_CUDA and _lib is not None:
        s += " (" + _lib.qp_last_error().decode() + ")"
    return s


def load_library(path: Optional[str] = None):
    """dlopen libqpb200.so.  Raises loudly when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the CUDA library is not built. Run "
            "`python -m embedded_qp_solver_b200.build` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.qp_options_default.argtypes = [C.POINTER(qp_options)]
    lib.qp_options_default.restype = None
    lib.qp_solve_batch.argtypes = [C.POINTER(qp_batch), C.POINTER(qp_batch_solution),
                                   C.POINTER(qp_options)]
    lib.qp_solve_batch.restype = C.c_int
    lib.qp_workspace_create.argtypes = [C.c_uint32, C.c_size_t, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.qp_workspace_create.restype = C.c_int
    lib.qp_workspace_destroy.argtypes = [C.c_void_p]
    lib.qp_workspace_destroy.restype = None
    lib.qp_solve_batch_ws.argtypes = [C.c_void_p, C.POINTER(qp_batch),
                                      C.POINTER(qp_batch_solution), C.POINTER(qp_options)]
    lib.qp_solve_batch_ws.restype = C.c_int
    lib.qp_eval_batch.argtypes = [C.POINTER(qp_batch), C.c_void_p, C.c_float, C.c_void_p,
                                  C.POINTER(qp_options)]
    lib.qp_eval_batch.restype = C.c_int
    lib.qp_admm_qf.argtypes = [C.POINTER(qp_ref_quadratic_form), C.c_uint, C.c_double, C.c_double,
                               C.POINTER(C.c_double), _ip, _ip]
    lib.qp_admm_qf.restype = C.c_int
    lib.qp_admm_qf_batch.argtypes = [C.POINTER(C.POINTER(qp_ref_quadratic_form)), C.c_size_t,
                                     C.c_uint, C.c_double, C.c_double,
                                     C.POINTER(C.POINTER(C.c_double)), _ip, _ip,
                                     C.POINTER(qp_options)]
    lib.qp_admm_qf_batch.restype = C.c_int
    lib.qp_newton_batch.argtypes = [C.POINTER(qp_batch), C.c_void_p, C.c_uint, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                    C.POINTER(qp_options)]
    lib.qp_newton_batch.restype = C.c_int
    lib.qp_gradient_batch.argtypes = lib.qp_newton_batch.argtypes
    lib.qp_gradient_batch.restype = C.c_int
    for fn in (lib.qp_newton_qf, lib.qp_gradient_qf):
        fn.argtypes = [C.POINTER(qp_ref_quadratic_form), C.POINTER(C.c_double), C.c_uint,
                       C.POINTER(C.c_double), _ip, _ip]
        fn.restype = C.c_int
    _dp = C.POINTER(C.c_double)
    lib.qp_problem_file_write.argtypes = [C.c_char_p, C.c_uint32, C.c_size_t, _dp, _dp, C.c_int]
    lib.qp_problem_file_write.restype = C.c_int
    lib.qp_problem_file_count.argtypes = [C.c_char_p, C.POINTER(C.c_uint32)]
    lib.qp_problem_file_count.restype = C.c_longlong
    lib.qp_problem_file_read.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_uint32, _dp, _dp]
    lib.qp_problem_file_read.restype = C.c_int
    lib.qp_solve_general_batch.argtypes = [C.POINTER(qp_general_batch),
                                           C.POINTER(qp_general_solution), C.POINTER(qp_options)]
    lib.qp_solve_general_batch.restype = C.c_int
    lib.qp_host_path_packs.argtypes = [C.c_uint32]
    lib.qp_host_path_packs.restype = C.c_int
    lib.qp_host_pack_upper.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int]
    lib.qp_host_pack_upper.restype = C.c_int
    lib.qp_host_path_h2d_bytes.argtypes = []
    lib.qp_host_path_h2d_bytes.restype = C.c_ulonglong
    lib.qp_packed_stride.argtypes = [C.c_uint32]
    lib.qp_packed_stride.restype = C.c_size_t
    lib.qp_probe_copy_seconds.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_int]
    lib.qp_probe_copy_seconds.restype = C.c_double
    lib.qp_schedule_stats.argtypes = [C.c_int, C.POINTER(C.c_ulonglong)]
    lib.qp_schedule_stats.restype = C.c_int
    lib.qp_last_error.restype = C.c_char_p
    lib.qp_device_count.restype = C.c_int
    lib.qp_kernel_for.argtypes = [C.c_uint32, C.c_int]
    lib.qp_kernel_for.restype = C.c_char_p
    lib.qp_probe_fp32_tflops.argtypes = [C.c_int, C.c_int]
    lib.qp_probe_fp32_tflops.restype = C.c_double
    lib.qp_probe_hbm_gbs.argtypes = [C.c_int, C.c_size_t, C.c_int]
    lib.qp_probe_hbm_gbs.restype = C.c_double
    _lib = lib
    return lib


@dataclass
class Options:
    """Mirror of ``qp_options``; defaults are the reference's constants
    (qp_solvers.c:17-18, :281-282; config.h:38)."""
    rho: float = 1.0
    alpha: float = 1.0
    abstol: float = 1e-4
    reltol: float = 1e-2
    max_iter: int = 10000
    device_mask: int = 0
    device: int = -1
    stream: int = 0
    chunk: int = 0
    kernel: int = 0   # KERNEL_AUTO / KERNEL_TPQ / KERNEL_GPQ / KERNEL_BPQ / KERNEL_CPQ (general rows: KERNEL_GEN_*)
    schedule: int = 0  # SCHEDULE_AUTO / SCHEDULE_SINGLE / SCHEDULE_PASSES
    precision: int = 0  # PRECISION_F32 / PRECISION_F64 (qf entry points only)

    def to_c(self) -> qp_options:
        return qp_options(self.rho, self.alpha, self.abstol, self.reltol, int(self.max_iter),
                          int(self.device_mask), int(self.device), C.c_void_p(self.stream or None),
                          int(self.chunk), int(self.kernel), int(self.schedule), int(self.precision))


def _is_torch(t) -> bool:
    return type(t).__module__.startswith("torch")


def _as_f32(a, shape, what):
    """Return (holder, address, is_cuda) for a float32 contiguous array-like."""
    if a is None:
        return None, None, None
    if _is_torch(a):
        import torch
        if a.dtype != torch.float32 or not a.is_contiguous():
            raise TypeError(f"{what}: need a contiguous float32 tensor")
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{what}: shape {tuple(a.shape)} != {tuple(shape)}")
        return a, a.data_ptr(), a.is_cuda
    arr = np.ascontiguousarray(a, dtype=np.float32)
    if arr.shape != tuple(shape):
        raise ValueError(f"{what}: shape {arr.shape} != {tuple(shape)}")
    return arr, arr.ctypes.data, False


def solve_batch(P, q, lb=None, ub=None, z0=None, u0=None, options: Optional[Options] = None,
                out: Optional[dict] = None, workspace: Optional["Workspace"] = None,
                want=("x", "z", "lam", "status", "iters"), rho=None, p_layout: int = P_DENSE,
                p_share: int = 1) -> dict:
    """Solve ``count`` box QPs: the batched replacement of a loop of the
    reference's ``admm(x0, iterations, qf)`` (qp_solvers.c:255-319).

    P [count,n,n], q [count,n] float32; lb/ub each an array [count,n], a Python
    scalar, or None (= -/+1e12, the reference's host box, config.h:29-30).
    rho: None (``options.rho`` for every problem, the reference's RHO_ADMM) or
    an array [count] of per-problem penalties.
    p_layout=P_UPPER: P is [count, packed_stride(n)], the row-major upper triangle of every
    matrix (``pack_upper``), n = 8, 12, 16.
    p_share=k > 1: problems i*k .. i*k+k-1 share P[i] (``qp_batch.p_share``: one model, many
    right-hand sides); P then holds ceil(count / k) matrices and count is q's.
    CUDA torch tensors take the device path (kernel launched on the current
    torch stream, no synchronisation, outputs are torch tensors on that
    device); numpy arrays / CPU tensors take the host path (copies, sharded
    over ``options.device_mask``) and return numpy arrays.
    """
    lib = load_library()
    opt = options or Options()
    p_share = max(1, int(p_share))
    count = int(q.shape[0]) if p_share > 1 else int(P.shape[0])
    nmat = (count + p_share - 1) // p_share
    if p_layout == P_UPPER:
        n = int(q.shape[1])
        pshape = (nmat, packed_stride(n))
    else:
        n = int(P.shape[1])
        pshape = (nmat, n, n)
    if tuple(P.shape) != pshape:
        raise ValueError(f"P must be {pshape}")
    hP, aP, dev = _as_f32(P, pshape, "P")
    hq, aq, dq = _as_f32(q, (count, n), "q")
    if dq != dev:
        raise ValueError("P and q must live on the same side (host or device)")

    pb = qp_batch()
    pb.n, pb.count = n, count
    pb.p_layout = int(p_layout)
    pb.p_share = p_share
    pb.mem = MEM_DEVICE if dev else MEM_HOST
    pb.P, pb.q = aP, aq
    pb.lb_scalar, pb.ub_scalar = -1e12, 1e12
    keep = [hP, hq]
    for name, val in (("lb", lb), ("ub", ub)):
        if val is None:
            continue
        if isinstance(val, (int, float)):
            setattr(pb, name + "_scalar", float(val))
        else:
            h, a, d = _as_f32(val, (count, n), name)
            if d != dev:
                raise ValueError(f"{name} must live where P lives")
            setattr(pb, name, a)
            keep.append(h)
    for name, val in (("z0", z0), ("u0", u0)):
        if val is not None:
            h, a, d = _as_f32(val, (count, n), name)
            if d != dev:
                raise ValueError(f"{name} must live where P lives")
            setattr(pb, name, a)
            keep.append(h)
    if rho is not None:
        h, a, d = _as_f32(rho, (count,), "rho")
        if d != dev:
            raise ValueError("rho must live where P lives")
        pb.rho = a
        keep.append(h)

    res = {} if out is None else out
    sol = qp_batch_solution()
    copt = opt.to_c()
    if dev:
        import torch
        device = hP.device
        copt.device = device.index if device.index is not None else torch.cuda.current_device()
        if not opt.stream:
            copt.stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream or None)
        for k in want:
            if k not in res:
                shape = (count,) if k in ("status", "iters") else (count, n)
                dt = torch.int32 if k in ("status", "iters") else torch.float32
                res[k] = torch.empty(shape, dtype=dt, device=device)
        for k in want:
            setattr(sol, k, res[k].data_ptr())
    else:
        for k in want:
            if k not in res:
                shape = (count,) if k in ("status", "iters") else (count, n)
                dt = np.int32 if k in ("status", "iters") else np.float32
                res[k] = np.empty(shape, dtype=dt)
        for k in want:
            v = res[k]
            setattr(sol, k, v.data_ptr() if _is_torch(v) else v.ctypes.data)

    if workspace is not None:
        rc = lib.qp_solve_batch_ws(workspace.handle, C.byref(pb), C.byref(sol), C.byref(copt))
    else:
        rc = lib.qp_solve_batch(C.byref(pb), C.byref(sol), C.byref(copt))
    if rc != QP_OK:
        raise QpError(rc, "qp_solve_batch")
    del keep
    return res


class Workspace:
    """Preallocated device workspace (``qp_workspace_create``), the replacement
    of the reference's static pools (kmalloc/kmalloc.c:37-66)."""

    def __init__(self, n: int, max_count: int, device_mask: int = 0):
        lib = load_library()
        h = C.c_void_p()
        rc = lib.qp_workspace_create(n, max_count, device_mask, C.byref(h))
        if rc != QP_OK:
            raise QpError(rc, "qp_workspace_create")
        self.handle = h

    def close(self):
        if self.handle:
            load_library().qp_workspace_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def eval_batch(P, q, x, r: float = 0.0):
    """f(x) = x'Px/2 + q'x + r per problem (quadratic_form_eval, qp.c:9-27)."""
    lib = load_library()
    count, n = int(P.shape[0]), int(P.shape[1])
    hP, aP, dev = _as_f32(P, (count, n, n), "P")
    hq, aq, _ = _as_f32(q, (count, n), "q")
    hx, ax, _ = _as_f32(x, (count, n), "x")
    pb = qp_batch()
    pb.n, pb.count, pb.mem, pb.P, pb.q = n, count, (MEM_DEVICE if dev else MEM_HOST), aP, aq
    copt = Options().to_c()
    if dev:
        import torch
        f = torch.empty(count, dtype=torch.float32, device=hP.device)
        copt.device = hP.device.index
        copt.stream = C.c_void_p(torch.cuda.current_stream(hP.device).cuda_stream or None)
        af = f.data_ptr()
    else:
        f = np.empty(count, np.float32)
        af = f.ctypes.data
    rc = lib.qp_eval_batch(C.byref(pb), ax, r, af, C.byref(copt))
    if rc != QP_OK:
        raise QpError(rc, "qp_eval_batch")
    return f


def solve_general_batch(P, q, A, l, u, z0=None, u0=None, options: Optional[Options] = None,
                        out: Optional[dict] = None):
    """``l <= A x <= u`` rows (``qp_solve_general_batch``): the reference's ADMM iteration for
    the splitting A x = z.  Returns dict(x, z, y, status, iters); device tensors in, device
    tensors out; numpy in, numpy out.  Host arrays: ``out`` may hold preallocated result arrays
    (numpy or pinned torch tensors, any subset of the keys); the library cuts the batch into
    chunks that alternate between two device buffers and two streams."""
    lib = load_library()
    count, n = int(P.shape[0]), int(P.shape[1])
    m = int(A.shape[1])
    hP, aP, dev = _as_f32(P, (count, n, n), "P")
    hq, aq, _ = _as_f32(q, (count, n), "q")
    hA, aA, _ = _as_f32(A, (count, m, n), "A")
    hl, al, _ = _as_f32(l, (count, m), "l")
    hu, au, _ = _as_f32(u, (count, m), "u")
    hz0, az0, _ = _as_f32(z0, (count, m), "z0")
    hu0, au0, _ = _as_f32(u0, (count, m), "u0")
    pb = qp_general_batch()
    pb.n, pb.m, pb.count, pb.mem = n, m, count, (MEM_DEVICE if dev else MEM_HOST)
    pb.P, pb.q, pb.A, pb.l, pb.u, pb.z0, pb.u0 = aP, aq, aA, al, au, az0, au0
    copt = (options or Options()).to_c()
    shapes = {"x": ((count, n), "f"), "z": ((count, m), "f"), "y": ((count, m), "f"),
              "status": ((count,), "i"), "iters": ((count,), "i")}
    if dev:
        import torch
        d = hP.device
        out = {k: torch.empty(sh, dtype=torch.float32 if t == "f" else torch.int32, device=d)
               for k, (sh, t) in shapes.items()}
        copt.device = d.index
        copt.stream = C.c_void_p(torch.cuda.current_stream(d).cuda_stream or None)
        addr = {k: v.data_ptr() for k, v in out.items()}
    else:
        # host arrays; `out` may bring preallocated ones (e.g. pinned torch tensors, which the copy
        # engine then writes in place), missing keys are allocated here
        res = {} if out is None else out
        for k, (sh, t) in shapes.items():
            if k not in res:
                res[k] = np.empty(sh, np.float32 if t == "f" else np.int32)
            elif tuple(res[k].shape) != sh:
                raise ValueError(f"out[{k!r}] must have shape {sh}")
        out = res
        addr = {k: (v.data_ptr() if _is_torch(v) else v.ctypes.data) for k, v in out.items()}
    sol = qp_general_solution(addr["x"], addr["z"], addr["y"], addr["status"], addr["iters"])
    rc = lib.qp_solve_general_batch(C.byref(pb), C.byref(sol), C.byref(copt))
    if rc != QP_OK:
        raise QpError(rc, "qp_solve_general_batch")
    return out


def gradient_batch(P, q, x0=None, iterations: int = 10000, r: float = 0.0, want_f: bool = False):
    """Batched ``gradient_descent_with_line_search`` (qp_solvers.c:65-101); same interface
    as :func:`newton_batch`; ``iterations`` defaults to the reference's GRAD_ITERATIONS."""
    return newton_batch(P, q, x0, iterations, r, want_f, _entry="qp_gradient_batch")


def newton_batch(P, q, x0=None, iterations: int = 10, r: float = 0.0, want_f: bool = False,
                 _entry: str = "qp_newton_batch"):
    """Batched ``newton_method_with_line_search`` (qp_solvers.c:103-144): returns a
    dict(x, status, steps[, f]).  Device tensors in, device tensors out; numpy in,
    numpy out.  ``iterations`` defaults to the reference's HESS_ITERATIONS."""
    lib = load_library()
    count, n = int(P.shape[0]), int(P.shape[1])
    hP, aP, dev = _as_f32(P, (count, n, n), "P")
    hq, aq, _ = _as_f32(q, (count, n), "q")
    hx0, ax0, _ = _as_f32(x0, (count, n), "x0")
    pb = qp_batch()
    pb.n, pb.count, pb.mem, pb.P, pb.q = n, count, (MEM_DEVICE if dev else MEM_HOST), aP, aq
    copt = Options().to_c()
    if dev:
        import torch
        d = hP.device
        out = {"x": torch.empty((count, n), dtype=torch.float32, device=d),
               "status": torch.empty(count, dtype=torch.int32, device=d),
               "steps": torch.empty(count, dtype=torch.int32, device=d)}
        if want_f:
            out["f"] = torch.empty(count, dtype=torch.float32, device=d)
        copt.device = d.index
        copt.stream = C.c_void_p(torch.cuda.current_stream(d).cuda_stream or None)
        addr = {k: v.data_ptr() for k, v in out.items()}
    else:
        out = {"x": np.empty((count, n), np.float32), "status": np.empty(count, np.int32),
               "steps": np.empty(count, np.int32)}
        if want_f:
            out["f"] = np.empty(count, np.float32)
        addr = {k: v.ctypes.data for k, v in out.items()}
    rc = getattr(lib, _entry)(C.byref(pb), ax0, int(iterations), addr["x"], addr["status"],
                              addr["steps"], addr.get("f"), float(r), C.byref(copt))
    if rc != QP_OK:
        raise QpError(rc, _entry)
    return out


# ---- the reference's problem file (test/test.c:108-126, test/qp_ref.py:8-30) --------------

def write_problem_file(path, P, q, append: bool = False) -> None:
    """[double n][P][q] per problem; one problem gives the reference's exact bytes."""
    lib = load_library()
    P = np.ascontiguousarray(P, np.float64)
    if P.ndim == 2:
        P = P[None]
    count, n = P.shape[0], P.shape[1]
    q = np.ascontiguousarray(q, np.float64).reshape(count, n)
    dp = C.POINTER(C.c_double)
    rc = lib.qp_problem_file_write(os.fsencode(path), n, count, P.ctypes.data_as(dp),
                                   q.ctypes.data_as(dp), int(append))
    if rc != QP_OK:
        raise QpError(rc, "qp_problem_file_write")


def read_problem_file(path, first: int = 0, count: Optional[int] = None):
    """Returns (P [count,n,n], q [count,n]) as float64."""
    lib = load_library()
    n = C.c_uint32()
    total = lib.qp_problem_file_count(os.fsencode(path), C.byref(n))
    if total < 0:
        raise QpError(int(total), "qp_problem_file_count")
    count = total - first if count is None else count
    P = np.empty((count, n.value, n.value))
    q = np.empty((count, n.value))
    dp = C.POINTER(C.c_double)
    rc = lib.qp_problem_file_read(os.fsencode(path), first, count, n.value,
                                  P.ctypes.data_as(dp), q.ctypes.data_as(dp))
    if rc != QP_OK:
        raise QpError(rc, "qp_problem_file_read")
    return P, q


# ---- the reference's own operator ------------------------------------------------------

class Matrix:
    """Python holder of a reference ``struct _matrix`` (double, row-major;
    rows in bits 31:16 and cols in bits 15:0 of ``dimensions``)."""

    def __init__(self, array):
        a = np.array(array, dtype=np.float64, order="C")
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        self.array = a
        self.c = qp_ref_matrix((a.shape[0] << 16) | a.shape[1],
                               a.ctypes.data_as(C.POINTER(C.c_double)))


class QuadraticForm:
    """Mirror of ``struct _quadratic_form {p, q, r}`` (qp.h:8-13): non-owning."""

    def __init__(self, p, q, r: float = 0.0):
        self.p = p if isinstance(p, Matrix) else Matrix(p)
        self.q = q if isinstance(q, Matrix) else Matrix(q)
        self.r = float(r)
        self.c = qp_ref_quadratic_form(C.pointer(self.p.c), C.pointer(self.q.c), self.r)


def admm(x0, iterations: int, qf: QuadraticForm, lb: float = -1e12, ub: float = 1e12,
         info: Optional[dict] = None) -> np.ndarray:
    """``struct _matrix *admm(x0, iterations, qf)`` (qp_solvers.c:255-319) on the
    GPU.  Like the reference, ``x0`` is accepted and ignored (qp_solvers.c:256),
    the box is the scalar pair (config.h:29-30) and x is returned as an n x 1
    float64 array."""
    lib = load_library()
    n = qf.p.array.shape[0]
    x = np.zeros((n, 1))
    st, it = C.c_int32(), C.c_int32()
    rc = lib.qp_admm_qf(C.byref(qf.c), int(iterations), float(lb), float(ub),
                        x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(st), C.byref(it))
    if rc != QP_OK:
        raise QpError(rc, "qp_admm_qf")
    if info is not None:
        info["status"], info["iters"] = st.value, it.value
    return x


def gradient_descent_with_line_search(x0, iterations: int, qf: QuadraticForm,
                                      info: Optional[dict] = None) -> np.ndarray:
    """``struct _matrix *gradient_descent_with_line_search(x0, iterations, qf)``
    (qp_solvers.c:65-101) on the GPU; x returned as an n x 1 float64 array."""
    return newton_method_with_line_search(x0, iterations, qf, info, _entry="qp_gradient_qf")


def newton_method_with_line_search(x0, iterations: int, qf: QuadraticForm,
                                   info: Optional[dict] = None,
                                   _entry: str = "qp_newton_qf") -> np.ndarray:
    """``struct _matrix *newton_method_with_line_search(x0, iterations, qf)``
    (qp_solvers.c:103-144) on the GPU; x returned as an n x 1 float64 array."""
    lib = load_library()
    n = qf.p.array.shape[0]
    x = np.zeros((n, 1))
    st, it = C.c_int32(), C.c_int32()
    x0a = None
    if x0 is not None:
        x0a = np.ascontiguousarray(x0.array if isinstance(x0, Matrix) else x0, np.float64).reshape(-1)
    rc = getattr(lib, _entry)(C.byref(qf.c),
                              x0a.ctypes.data_as(C.POINTER(C.c_double)) if x0a is not None else None,
                              int(iterations), x.ctypes.data_as(C.POINTER(C.c_double)),
                              C.byref(st), C.byref(it))
    if rc != QP_OK:
        raise QpError(rc, _entry)
    if info is not None:
        info["status"], info["steps"] = st.value, it.value
    return x


def admm_many(qfs, iterations: int, lb: float = -1e12, ub: float = 1e12, options: Optional[Options] = None):
    """Array-of-structs entry (``qp_admm_qf_batch``): a list of QuadraticForm
    solved as one batch.  Returns (x [count,n] float64, status, iters).
    ``options.precision = PRECISION_F64`` solves in the reference's own binary64."""
    lib = load_library()
    count = len(qfs)
    n = qfs[0].p.array.shape[0]
    arr = (C.POINTER(qp_ref_quadratic_form) * count)(*[C.pointer(f.c) for f in qfs])
    x = np.zeros((count, n))
    xp = (C.POINTER(C.c_double) * count)(
        *[x[i].ctypes.data_as(C.POINTER(C.c_double)) for i in range(count)])
    st = np.zeros(count, np.int32)
    it = np.zeros(count, np.int32)
    copt = options.to_c() if options is not None else None
    rc = lib.qp_admm_qf_batch(arr, count, int(iterations), float(lb), float(ub), xp,
                              st.ctypes.data_as(_ip), it.ctypes.data_as(_ip),
                              C.byref(copt) if copt is not None else None)
    if rc != QP_OK:
        raise QpError(rc, "qp_admm_qf_batch")
    return x, st, it


def packed_stride(n: int) -> int:
    """Floats per problem of a P_UPPER batch (``qp_packed_stride``)."""
    return int(load_library().qp_packed_stride(n))


def pack_upper(P) -> np.ndarray:
    """[count, n, n] -> [count, packed_stride(n)]: the row-major upper triangle of each
    matrix, zero padded (the P_UPPER layout).  numpy, any n the layout takes."""
    P = np.asarray(P, dtype=np.float32)
    count, n = P.shape[0], P.shape[1]
    iu = np.triu_indices(n)
    out = np.zeros((count, packed_stride(n)), np.float32)
    out[:, :len(iu[0])] = P[:, iu[0], iu[1]]
    return out


def probe_copy_seconds(h2d_bytes: int, d2h_bytes: int, device: int = -1, pinned: bool = True,
                       piece_bytes: int = 0, repeats: int = 3) -> float:
    """Seconds to move those bytes both ways at once with no kernel in between
    (``qp_probe_copy_seconds``): the ceiling of the host-memory path."""
    t = float(load_library().qp_probe_copy_seconds(device, h2d_bytes, d2h_bytes, int(pinned),
                                                   piece_bytes, repeats))
    if t < 0:
        raise QpError(int(t), "qp_probe_copy_seconds")
    return t


def schedule_stats(device: int = -1) -> dict:
    """Counters of the schedules (``qp_schedule_stats``): passes -- problems that outlived the
    first window, problems looked at, evictions by later passes; AUTO's probing single launches --
    sum of iteration counts, sum of what the warps waited for, problems looked at."""
    v = (C.c_ulonglong * 6)()
    rc = load_library().qp_schedule_stats(device, v)
    if rc != QP_OK:
        raise QpError(rc, "qp_schedule_stats")
    return {"outlived": int(v[0]), "seen": int(v[1]), "evicted_later": int(v[2]),
            "probe_iterations": int(v[3]), "probe_waited_for": int(v[4]), "probe_problems": int(v[5])}


def host_path_h2d_bytes() -> int:
    """Bytes the host-memory path has sent host to device so far in this process."""
    return int(load_library().qp_host_path_h2d_bytes())


def host_path_packs(n: int) -> bool:
    """True when the host-memory path sends packed upper triangles of P over PCIe."""
    return bool(load_library().qp_host_path_packs(n))


def active_set(z, lb, ub):
    """The active set of a solution (outputs the reference lacks): +1 where
    the coordinate sits on its upper bound, -1 on its lower bound, 0 where it is free.  z is
    produced by clipping (qp_solvers.c:186-187), so exact equality is what "active" means; the
    sign agrees with that of lambda = rho u wherever lambda is not zero.  numpy arrays or torch
    tensors; lb, ub scalars or arrays that broadcast against z."""
    if hasattr(z, "device") and hasattr(z, "new_tensor"):  # torch
        import torch
        lbt = lb if hasattr(lb, "device") else z.new_tensor(lb)
        ubt = ub if hasattr(ub, "device") else z.new_tensor(ub)
        return (z == ubt).to(torch.int8) - (z == lbt).to(torch.int8)
    z = np.asarray(z)
    return (z == np.asarray(ub, dtype=z.dtype)).astype(np.int8) - (z == np.asarray(lb, dtype=z.dtype)).astype(np.int8)


def kernel_for(n: int, aligned16: bool = True) -> str:
    return load_library().qp_kernel_for(n, int(aligned16)).decode()


def probe_fp32_tflops(device: int = -1, repeats: int = 5) -> float:
    return float(load_library().qp_probe_fp32_tflops(device, repeats))


def probe_hbm_gbs(device: int = -1, nbytes: int = 1 << 30, repeats: int = 5) -> float:
    return float(load_library().qp_probe_hbm_gbs(device, nbytes, repeats))


# ---- roofline bookkeeping --------------------------------------------------

def algorithmic_bytes(n: int, warm: bool = False, with_z: bool = True,
                      bound_arrays: bool = True) -> int:
    """HBM bytes one problem must move: in P, q (+ lb, ub when they
    are arrays; the reference's scalar box moves nothing) [+ z0, u0 warm start],
    out x, lambda (+ z) and status, iters.
    n=16: 1352 with bound arrays and no z (the SURVEY figure); 1288 for the
    benchmark configuration (scalar box, z written)."""
    b = (n * n + n) * 4 + (2 * n * 4 if bound_arrays else 0) + (2 * n * 4 if warm else 0)
    b += 2 * n * 4 + 8
    return b + (4 * n if with_z else 0)


def algorithmic_flops(n: int, mean_iters: float) -> float:
    """Reference-algorithm flops per problem, F = 8/3 n^3 + K (2 n^2 + 24 n)."""
    return 8.0 / 3.0 * n ** 3 + mean_iters * (2 * n * n + 24 * n)
