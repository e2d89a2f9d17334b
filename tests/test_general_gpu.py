"""GPU: qp_solve_general_batch (rows l <= A x <= u, csrc/qp_gen.cuh).

* A = I: the kernel must reproduce the reference's admm() -- same golden vectors, same bars
  as the box kernels (iteration counts identical, x within 1e-4).
* general A: no reference behaviour exists (parity unpinned); the kernel is compared with the
  float64 numpy restatement oracle/general_np.py, with the KKT conditions at a tight
  tolerance, and with an independent scipy solve."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden, nondegenerate
from embedded_qp_solver_b200 import problems
from oracle import general_np

pytestmark = pytest.mark.gpu
X_TOL = 1e-4


def rel_err(x, x_ref):
    return np.max(np.abs(x - x_ref), axis=1) / (1 + np.max(np.abs(x_ref), axis=1))


def random_rows(count, n, m, seed, scale=0.5):
    rng = np.random.default_rng(seed)
    A = (rng.normal(size=(count, m, n)) / np.sqrt(n)).astype(np.float32)
    l = (-scale * rng.uniform(0.2, 1.0, (count, m))).astype(np.float32)
    u = (scale * rng.uniform(0.2, 1.0, (count, m))).astype(np.float32)
    return A, l, u


@pytest.mark.parametrize("name", ["wellcond_b05_n16.npz", "wellcond_b05_n40.npz",
                                  "wellcond_free_n8.npz", "mpc_n40.npz", "wellcond_b05_n64.npz"])
def test_identity_rows_reproduce_the_reference_golden_vectors(qp, name):
    g = load_golden(os.path.join(os.path.dirname(golden_files()[0]), name))
    count, n = g["q"].shape
    eye = np.ascontiguousarray(np.broadcast_to(np.eye(n, dtype=np.float32), (count, n, n)))
    r = qp.solve_general_batch(g["P"], g["q"], eye, g["lb"], g["ub"])
    ok = nondegenerate(g["margins"]) & (g["status"] == 0)
    assert ok.sum() >= 0.5 * count
    assert np.array_equal(r["iters"][ok], g["iters"][ok])
    assert np.array_equal(r["status"][ok], g["status"][ok])
    assert rel_err(r["x"][ok].astype(np.float64), g["x_ref"][ok]).max() <= X_TOL
    assert rel_err(r["z"][ok].astype(np.float64), g["z"][ok]).max() <= X_TOL
    # and it is the same answer the tuned box kernels give
    b = qp.solve_batch(g["P"], g["q"], g["lb"], g["ub"])
    assert np.array_equal(b["iters"][ok], r["iters"][ok])
    assert rel_err(b["x"][ok].astype(np.float64), r["x"][ok].astype(np.float64)).max() <= X_TOL


@pytest.mark.parametrize("n,m", [(4, 8), (16, 32), (5, 3), (40, 80), (64, 128), (33, 70),
                                 (8, 16), (9, 16), (24, 48), (20, 64), (32, 65), (1, 1)])
def test_general_rows_against_the_numpy_restatement(qp, n, m):
    # counts that leave the last warp's groups ragged (4, 8, 16 or 32 lanes per problem)
    count = 43 if n <= 16 else 11
    P, q, _, _ = problems.wellcond(count, n, seed=2000 + n)
    A, l, u = random_rows(count, n, m, seed=n)
    o = general_np.admm_general_batch(P, q, A, l, u)
    tP, tq, tA, tl, tu = (torch.from_numpy(v).cuda() for v in (P, q, A, l, u))
    r = qp.solve_general_batch(tP, tq, tA, tl, tu)
    torch.cuda.synchronize()
    r = {k: v.cpu().numpy() for k, v in r.items()}
    ok = nondegenerate(o["margins"]) & (o["status"] == 0)
    assert ok.sum() >= 0.7 * count
    assert np.array_equal(r["iters"][ok], o["iters"][ok]) and np.all(r["status"][ok] == 0)
    assert rel_err(r["x"][ok].astype(np.float64), o["x"][ok]).max() <= X_TOL
    assert rel_err(r["z"][ok].astype(np.float64), o["z"][ok]).max() <= X_TOL
    assert rel_err(r["y"][ok].astype(np.float64), o["u"][ok]).max() <= 10 * X_TOL
    # host-pointer path: same bits
    h = qp.solve_general_batch(P, q, A, l, u)
    assert np.array_equal(h["x"], r["x"]) and np.array_equal(h["iters"], r["iters"])


def test_general_rows_kkt_and_scipy_at_tight_tolerance(qp):
    from scipy.optimize import LinearConstraint, minimize
    count, n, m = 6, 8, 14
    P, q, _, _ = problems.wellcond(count, n, seed=77)
    A, l, u = random_rows(count, n, m, seed=78, scale=0.3)
    l[:, :3] = -np.inf  # one-sided rows
    r = qp.solve_general_batch(P, q, A, l, u,
                               options=qp.Options(abstol=1e-7, reltol=1e-7, max_iter=20000))
    assert np.all(r["status"] == 0)
    for c in range(count):
        P64, q64, A64 = P[c].astype(np.float64), q[c].astype(np.float64), A[c].astype(np.float64)
        x, y = r["x"][c].astype(np.float64), r["y"][c].astype(np.float64)
        assert np.max(np.abs(P64 @ x + q64 + A64.T @ y)) <= 2e-5 * (1 + np.max(np.abs(q64)))
        ax = A64 @ x
        assert np.all(ax >= l[c] - 1e-4) and np.all(ax <= u[c] + 1e-4)
        sp = minimize(lambda v: 0.5 * v @ P64 @ v + q64 @ v, np.zeros(n),
                      jac=lambda v: P64 @ v + q64, hess=lambda v: P64, method="trust-constr",
                      constraints=[LinearConstraint(A64, l[c], u[c])],
                      options=dict(gtol=1e-10, xtol=1e-12, maxiter=2000))
        assert np.max(np.abs(x - sp.x)) <= 1e-3 * (1 + np.max(np.abs(sp.x)))


def test_general_rows_warm_start_relaxation_and_status(qp):
    count, n, m = 30, 16, 32
    P, q, _, _ = problems.wellcond(count, n, seed=91)
    A, l, u = random_rows(count, n, m, seed=92)
    first = qp.solve_general_batch(P, q, A, l, u)
    q2 = problems.perturb_q(q, seed=5)
    z0, u0 = first["z"], first["y"]  # rho = 1
    o = general_np.admm_general_batch(P, q2, A, l, u, z0=z0.astype(np.float64),
                                      u0=u0.astype(np.float64), alpha=1.6)
    r = qp.solve_general_batch(P, q2, A, l, u, z0=z0, u0=u0, options=qp.Options(alpha=1.6))
    ok = nondegenerate(o["margins"]) & (o["status"] == 0)
    assert np.array_equal(r["iters"][ok], o["iters"][ok])
    assert rel_err(r["x"][ok].astype(np.float64), o["x"][ok]).max() <= X_TOL
    cold = general_np.admm_general_batch(P, q2, A, l, u, alpha=1.6)
    assert o["iters"].mean() < cold["iters"].mean()
    # P + rho A'A not positive definite -> status 2, neighbours unaffected
    Pb = P.copy()
    Pb[4] = -40 * np.eye(n, dtype=np.float32)
    rb = qp.solve_general_batch(Pb, q, A, l, u)
    assert rb["status"][4] == qp.STATUS_SINGULAR and rb["iters"][4] == 0
    assert np.array_equal(np.delete(rb["iters"], 4), np.delete(first["iters"], 4))
    with pytest.raises(qp.QpError):
        qp.solve_general_batch(np.zeros((1, 4, 4), np.float32), np.zeros((1, 4), np.float32),
                               np.zeros((1, 129, 4), np.float32), np.zeros((1, 129), np.float32),
                               np.zeros((1, 129), np.float32))
