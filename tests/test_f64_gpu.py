"""GPU: the binary64 box ADMM behind the qf entry points (qp_options.precision = QP_PRECISION_F64,
csrc/qp_f64.cuh) against the oracle in the reference's own arithmetic.

The reference computes in double (/root/reference/matrix/matrix_type/matrix_type.h:20-23).  Its
own test problems (matrix_ops.c:699-734, config.h:16-20) are ill-conditioned and stop after
3 ... 10 000 iterations on a reltol = 1e-2 test: a binary32 run reproduces the reference's
iteration count on about two problems in three there.  In binary64 the
GPU reproduces it on all of them but the odd one whose stop test is decided in the last bits."""
import numpy as np
import pytest

from embedded_qp_solver_b200 import problems

pytestmark = pytest.mark.gpu


def qfs_of(qp, P, q):
    return [qp.QuadraticForm(P[i], q[i], 0.0) for i in range(len(P))]


@pytest.mark.parametrize("n", [4, 16, 33, 40, 64])
def test_reference_distribution_iteration_counts(qp, orc, n):
    count = 200 if n <= 40 else 64
    P, q, _, _ = problems.refdist(count, n, seed=300 + n, dtype=np.float64)
    box = 1e4
    o = orc.admm(P, q, -box, box, max_iter=10000)
    x, st, it = qp.admm_many(qfs_of(qp, P, q), 10000, -box, box,
                             options=qp.Options(precision=qp.PRECISION_F64))
    agree = it == o["iters"]
    assert agree.mean() >= 0.98, f"n={n}: {agree.sum()} of {count} iteration counts agree"
    assert np.array_equal(st[agree], o["status"][agree])
    err = np.max(np.abs(x[agree] - o["x"][agree]), axis=1) / (1 + np.max(np.abs(o["x"][agree]), axis=1))
    assert err.max() <= 1e-9, err.max()
    # the binary32 product path on the same problems, for the record: clearly fewer agree
    x32, st32, it32 = qp.admm_many(qfs_of(qp, P, q), 10000, -box, box)
    assert (it32 == o["iters"]).mean() <= agree.mean()
    assert o["iters"].max() > 200    # the distribution is the heavy-tailed one


@pytest.mark.parametrize("n", [8, 16, 48])
def test_well_conditioned_and_active_box(qp, orc, n):
    count = 300
    P, q, _, _ = problems.wellcond(count, n, seed=11 + n, beta=0.5, dtype=np.float64)
    o = orc.admm(P, q, -0.5, 0.5)
    x, st, it = qp.admm_many(qfs_of(qp, P, q), 10000, -0.5, 0.5,
                             options=qp.Options(precision=qp.PRECISION_F64))
    assert np.array_equal(it, o["iters"]) and np.array_equal(st, o["status"])
    assert np.max(np.abs(x - o["x"])) <= 1e-12
    # options travel: rho, alpha, tolerances, the iteration cap
    o2 = orc.admm(P, q, -0.5, 0.5, rho=1.7, alpha=1.5, abstol=1e-6, reltol=1e-5, max_iter=12)
    x2, st2, it2 = qp.admm_many(qfs_of(qp, P, q), 12, -0.5, 0.5,
                                options=qp.Options(precision=qp.PRECISION_F64, rho=1.7, alpha=1.5,
                                                   abstol=1e-6, reltol=1e-5))
    assert np.array_equal(it2, o2["iters"]) and np.array_equal(st2, o2["status"])
    assert np.max(np.abs(x2 - o2["x"])) <= 1e-10
    assert (st2 == 1).sum() == (o2["status"] == 1).sum()   # the cap of 12 is hit by the same problems


def test_float_entry_points_refuse_binary64(qp):
    from embedded_qp_solver_b200 import api
    P, q, _, _ = problems.wellcond(4, 8, seed=1)
    with pytest.raises(api.QpError) as e:
        qp.solve_batch(P, q, -0.5, 0.5, options=qp.Options(precision=qp.PRECISION_F64))
    assert e.value.code == api.QP_ERR_UNSUPPORTED
    # not positive definite: status 2, as in binary32
    Pb = np.eye(8)[None].repeat(2, 0); Pb[1, 3, 3] = -5.0
    x, st, it = qp.admm_many(qfs_of(qp, Pb, np.ones((2, 8))), 100, -1.0, 1.0,
                             options=qp.Options(precision=qp.PRECISION_F64))
    assert st.tolist() == [0, 2] and it[1] == 0
