"""GPU: batched Newton (qp_newton_batch) against the oracle's restatement of
newton_method_with_line_search (qp_solvers.c:103-144), which is bit-identical to the
unmodified reference (tests/test_oracle_golden.py), and against the golden x_newton
vectors the reference itself produced.

Bars: the number of Newton steps identical wherever the stop test ||g|| < 0.1 is not
within 1e-3 (relative) of flipping; x within 1e-4 of the FP64 reference."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden
from embedded_qp_solver_b200 import problems

pytestmark = pytest.mark.gpu
X_TOL = 1e-4


def rel_err(x, x_ref):
    return np.max(np.abs(x - x_ref), axis=1) / (1 + np.max(np.abs(x_ref), axis=1))


def oracle_newton(orc, P, q, x0, iterations):
    xs, steps = [], []
    for c in range(P.shape[0]):
        x, s = orc.newton(P[c].astype(np.float64), q[c].astype(np.float64),
                          x0[c].astype(np.float64), max_iter=iterations)
        xs.append(x), steps.append(s)
    return np.array(xs), np.array(steps)


def grad_margin(P, q, x0, steps):
    """How far the stop test (0.1 > ||g||) was from flipping: the reference's Newton
    contracts g by exactly 1 - alpha = 0.1 per step (see qp_newton.cuh), so
    ||g_k|| = 0.1^k ||g_0||."""
    g0 = np.linalg.norm(np.einsum("cij,cj->ci", P.astype(np.float64), x0.astype(np.float64)) + q,
                        axis=1)
    gk = g0 * 0.1 ** steps
    gprev = g0 * 0.1 ** np.maximum(steps - 1, 0)
    return np.minimum(np.abs(gk / 0.1 - 1), np.abs(gprev / 0.1 - 1))


@pytest.mark.parametrize("n", [4, 5, 16, 33, 40, 64])
def test_newton_against_oracle(qp, orc, n):
    count = 64 if n <= 16 else 24
    P, q, _, _ = problems.wellcond(count, n, seed=1200 + n)
    rng = np.random.default_rng(n)
    x0 = rng.uniform(-1e3, 1e3, (count, n)).astype(np.float32)  # X_RAND_ENTRY range, config.h
    xo, so = oracle_newton(orc, P, q, x0, 10)
    r = qp.newton_batch(torch.from_numpy(P).cuda(), torch.from_numpy(q).cuda(),
                        torch.from_numpy(x0).cuda(), iterations=10, want_f=True)
    torch.cuda.synchronize()
    x, steps, st = r["x"].cpu().numpy(), r["steps"].cpu().numpy(), r["status"].cpu().numpy()
    ok = grad_margin(P, q, x0, so) > 1e-3
    assert ok.sum() >= 0.8 * count
    assert np.array_equal(steps[ok], so[ok])
    assert np.all(st[ok] == np.where(so[ok] < 10, 0, 1))
    assert rel_err(x[ok].astype(np.float64), xo[ok]).max() <= X_TOL
    f_ref = np.array([orc.eval(P[c], q[c], 0.0, xo[c]) for c in range(count)])
    f = r["f"].cpu().numpy()
    assert np.max(np.abs(f[ok] - f_ref[ok]) / (1 + np.abs(f_ref[ok]))) <= 1e-3
    # host-pointer path gives the same bits as the device-pointer path
    h = qp.newton_batch(P, q, x0, iterations=10)
    assert np.array_equal(h["x"], x) and np.array_equal(h["steps"], steps)


@pytest.mark.parametrize("path", [p for p in golden_files() if "wellcond_free" in p],
                         ids=os.path.basename)
def test_newton_golden_reference_vectors(qp, path):
    """x_newton in the golden files came from the UNMODIFIED reference's
    newton_method_with_line_search started at x0 = 0 with HESS_ITERATIONS = 10
    (tests/golden/make_golden.py)."""
    g = load_golden(path)
    r = qp.newton_batch(g["P"], g["q"], None, iterations=10)
    P64, q64 = g["P"].astype(np.float64), g["q"].astype(np.float64)
    ok = grad_margin(P64, q64, np.zeros_like(q64), r["steps"]) > 1e-3
    assert ok.sum() >= 0.8 * len(ok)
    assert rel_err(r["x"][ok].astype(np.float64), g["x_newton"][ok]).max() <= X_TOL


def test_newton_agrees_with_admm_when_the_box_is_inactive(qp):
    """The reference's own cross-check (it prints both objectives, test/test.c:100-102):
    with the box inactive, ADMM run to a tight tolerance and Newton run to its cap reach
    the same minimiser -P^-1 q."""
    P, q, _, _ = problems.wellcond(200, 16, seed=1300)
    tP, tq = torch.from_numpy(P).cuda(), torch.from_numpy(q).cuda()
    a = qp.solve_batch(tP, tq, -1e12, 1e12,
                       options=qp.Options(abstol=1e-7, reltol=1e-7, max_iter=5000))
    x0 = np.full_like(q, 50.0)
    nw = qp.newton_batch(tP, tq, torch.from_numpy(x0).cuda(), iterations=10)
    torch.cuda.synchronize()
    xs = np.linalg.solve(P.astype(np.float64), -q.astype(np.float64)[..., None])[..., 0]
    # Newton stops at ||g|| < 0.1 (MIN_GRAD_NEWTON): its distance to the minimiser is
    # bounded by 0.1 / lambda_min(P) = 0.1 here (P = G'G/n + I)
    assert np.max(np.abs(nw["x"].cpu().numpy() - xs)) <= 0.1
    assert np.max(np.abs(a["x"].cpu().numpy() - xs)) <= 1e-4


def test_newton_status_and_api_misuse(qp):
    P, q, _, _ = problems.wellcond(10, 8, seed=5)
    P[2] = -P[2]
    x0 = np.ones_like(q)
    r = qp.newton_batch(P, q, x0, iterations=10)
    assert r["status"][2] == qp.STATUS_SINGULAR and r["steps"][2] == 0
    assert np.array_equal(r["x"][2], x0[2])
    assert np.all(np.delete(r["status"], 2) == 0)
    r0 = qp.newton_batch(P, q, x0, iterations=0)      # no steps allowed: x = x0, cap status
    assert np.array_equal(r0["x"], x0) and np.all(np.delete(r0["status"], 2) == 1)
    with pytest.raises(qp.QpError):
        qp.newton_batch(np.zeros((1, 65, 65), np.float32), np.zeros((1, 65), np.float32))


@pytest.mark.parametrize("n", [4, 16, 40, 64])
def test_gradient_descent_against_oracle(qp, orc, n):
    """qp_gradient_batch vs the oracle's restatement of gradient_descent_with_line_search
    (qp_solvers.c:65-101; bit-identical to the unmodified reference,
    tests/test_oracle_vs_ref.py).  Every step takes a discrete decision (how often alpha is
    shrunk), so binary32 and binary64 follow the same path only while no decision is a near
    tie: the bar is the same step count on most problems, x within 1e-3 there, and the
    reference's own stop criterion ||P x + q|| < 0.1 everywhere."""
    count = 48 if n <= 16 else 16
    P, q, _, _ = problems.wellcond(count, n, seed=1400 + n)
    rng = np.random.default_rng(n)
    x0 = rng.uniform(-10, 10, (count, n)).astype(np.float32)
    xs, ss = [], []
    for c in range(count):
        x, s = orc.gradient(P[c].astype(np.float64), q[c].astype(np.float64),
                            x0[c].astype(np.float64), max_iter=10000)
        xs.append(x), ss.append(s)
    xo, so = np.array(xs), np.array(ss)
    r = qp.gradient_batch(torch.from_numpy(P).cuda(), torch.from_numpy(q).cuda(),
                          torch.from_numpy(x0).cuda(), iterations=10000)
    torch.cuda.synchronize()
    x, steps, st = r["x"].cpu().numpy(), r["steps"].cpu().numpy(), r["status"].cpu().numpy()
    assert np.all(st == 0)
    same = steps == so
    assert same.mean() >= 0.7 and np.max(np.abs(steps - so)) <= 2
    assert rel_err(x[same].astype(np.float64), xo[same]).max() <= 1e-3
    g = np.einsum("cij,cj->ci", P.astype(np.float64), x.astype(np.float64)) + q
    assert np.linalg.norm(g, axis=1).max() < 0.1 * (1 + 1e-3)
    # the cap is honoured and reported
    capped = qp.gradient_batch(P, q, x0, iterations=3)
    assert np.all(capped["steps"] == 3) and np.all(capped["status"] == 1)
