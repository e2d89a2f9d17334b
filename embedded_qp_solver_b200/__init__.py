"""B200-native batched box-QP engine.

A drop-in for the ADMM hot path of YangLingyuan/Embedded-qp-solver
(``admm()`` in qp_solvers/qp_solvers.c:255-319 on top of matrix/matrix_ops):
hand-written sm_100a kernels behind a C ABI (``include/qp_batch.h``,
``lib/libqpb200.so``); this package is the ctypes mirror of that ABI.

The CUDA library is loaded on first use; if it is missing the call raises --
there is no CPU path in this package.
"""
from .api import (P_DENSE, P_UPPER, PRECISION_F32, PRECISION_F64, SCHEDULE_AUTO, SCHEDULE_PASSES, SCHEDULE_SINGLE, pack_upper, packed_stride, probe_copy_seconds, schedule_stats, ABI_SYMBOLS, KERNEL_AUTO, KERNEL_BPQ, KERNEL_CPQ, KERNEL_GEN_REG, KERNEL_GEN_SMEM, KERNEL_GPQ, KERNEL_TPQ, LIB_PATH, QP_MAX_DIM, QP_MAX_DIM_BOX, STATUS_CONVERGED, STATUS_MAX_ITER,
                  STATUS_NONFINITE, STATUS_SINGULAR, Matrix, Options, QpError, QuadraticForm,
                  Workspace, active_set, admm, admm_many, algorithmic_bytes, algorithmic_flops, eval_batch, gradient_batch, gradient_descent_with_line_search, host_path_h2d_bytes, host_path_packs,
                  kernel_for, load_library, newton_batch, newton_method_with_line_search,
                  probe_fp32_tflops, probe_hbm_gbs, read_problem_file, solve_batch, solve_general_batch,
                  write_problem_file)
from . import problems  # noqa: F401

__all__ = [
    "P_DENSE", "P_UPPER", "PRECISION_F32", "PRECISION_F64", "SCHEDULE_AUTO", "SCHEDULE_PASSES", "SCHEDULE_SINGLE", "pack_upper", "packed_stride", "probe_copy_seconds", "schedule_stats",
    "ABI_SYMBOLS", "KERNEL_AUTO", "KERNEL_BPQ", "KERNEL_CPQ", "KERNEL_GEN_REG", "KERNEL_GEN_SMEM", "KERNEL_GPQ", "KERNEL_TPQ", "LIB_PATH", "QP_MAX_DIM", "QP_MAX_DIM_BOX", "STATUS_CONVERGED", "STATUS_MAX_ITER",
    "STATUS_NONFINITE", "STATUS_SINGULAR", "active_set", "Matrix", "Options", "QpError", "QuadraticForm",
    "Workspace", "admm", "admm_many", "algorithmic_bytes", "algorithmic_flops", "eval_batch", "gradient_batch", "gradient_descent_with_line_search", "host_path_h2d_bytes", "host_path_packs",
    "kernel_for", "load_library", "newton_batch", "newton_method_with_line_search", "probe_fp32_tflops", "probe_hbm_gbs", "read_problem_file", "solve_batch", "solve_general_batch", "write_problem_file",
    "problems",
]
