"""CPU, only where the unmodified reference was compiled (oracle/_ref, built from
/root/reference by oracle/Makefile): pins the restatement to the reference on
fresh seeds, and pins the struct layouts the C ABI mirrors."""
import numpy as np
import pytest

import oracle
from embedded_qp_solver_b200 import problems

SHAPES = [n for n in (4, 8, 16, 40, 64, 96, 128) if oracle.ref_available(n)]
pytestmark = pytest.mark.skipif(not SHAPES, reason="oracle/_ref not built (no /root/reference)")


@pytest.mark.parametrize("n", SHAPES)
def test_restatement_bit_identical_to_reference(orc, n):
    ref = oracle.Reference(n)
    cnt = 24 if n <= 16 else 6
    for gen in (lambda: problems.wellcond(cnt, n, seed=11, beta=0.5),
                lambda: problems.wellcond(cnt, n, seed=12, beta=1e12),
                lambda: problems.refdist(cnt, n, seed=13, box=1e4)):
        P, q, lb, ub = (a.astype(np.float64) for a in gen())
        x_ref, _ = ref.admm_many(P, q, lb, ub)
        o = orc.admm(P, q, lb, ub)
        assert np.array_equal(x_ref, o["x"])


@pytest.mark.parametrize("n", SHAPES[:3])
def test_per_coordinate_box_through_unmodified_loop(orc, n):
    """The shim config.h makes qp_solvers.c:277-280 install a per-coordinate box."""
    ref = oracle.Reference(n)
    P, q, _, _ = (a.astype(np.float64) for a in problems.wellcond(8, n, seed=21))
    rng = np.random.default_rng(5)
    lb = -rng.uniform(0.1, 1.0, (8, n))
    ub = rng.uniform(0.1, 1.0, (8, n))
    x_ref, _ = ref.admm_many(P, q, lb, ub)
    o = orc.admm(P, q, lb, ub)
    assert np.array_equal(x_ref, o["x"])


def test_reference_struct_layout():
    """struct _matrix is 16 bytes with elements at 8; struct _quadratic_form is
    24 bytes with q at 8 and r at 16; rows live in bits 31:16 (matrix_type.h:14-23,
    qp.h:8-13) -- what include/qp_batch.h's qp_ref_* structs mirror."""
    import ctypes as C
    from embedded_qp_solver_b200 import api
    ref = oracle.Reference(SHAPES[0])
    assert ref.layout() == [16, 8, 24, 8, 16, (7 << 16) | 3]
    assert C.sizeof(api.qp_ref_matrix) == 16 and api.qp_ref_matrix.elements.offset == 8
    assert C.sizeof(api.qp_ref_quadratic_form) == 24
    assert api.qp_ref_quadratic_form.q.offset == 8 and api.qp_ref_quadratic_form.r.offset == 16


def test_reference_invert_and_eval_match_restatement(orc):
    n = SHAPES[0]
    ref = oracle.Reference(n)
    rng = np.random.default_rng(9)
    B = rng.uniform(-1e3, 1e3, (n, n))
    A = B.T @ B / (1e3 * n)
    Ai, _ = orc.invert(A)
    assert np.array_equal(ref.invert(A), Ai)
    q = rng.standard_normal(n)
    x = rng.standard_normal(n)
    assert ref.eval(A, q, 0.5, x) == orc.eval(A, q, 0.5, x)


@pytest.mark.parametrize("n", [4, 16, 40])
def test_gradient_restatement_is_bit_identical_to_the_reference(orc, n):
    """gradient_descent_with_line_search (qp_solvers.c:65-101) of the unmodified reference
    vs the plain-C restatement, started where the reference's tests start it (random x0)."""
    import oracle
    from embedded_qp_solver_b200 import problems
    ref = oracle.Reference(n)
    P, q, _, _ = problems.wellcond(5, n, seed=9)
    rng = np.random.default_rng(n)
    for c in range(5):
        x0 = rng.uniform(-10, 10, n)
        xr = ref.gradient(P[c].astype(np.float64), q[c].astype(np.float64), x0, 10000)
        xo, steps = orc.gradient(P[c].astype(np.float64), q[c].astype(np.float64), x0, 10000)
        assert np.array_equal(xr, xo) and 0 < steps < 10000
