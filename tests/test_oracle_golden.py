"""CPU: the restatement (oracle/admm_oracle.c) against the golden vectors that
the UNMODIFIED reference produced (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden


@pytest.mark.parametrize("path", golden_files(), ids=os.path.basename)
def test_oracle_reproduces_reference_bit_for_bit(orc, path):
    g = load_golden(path)
    o = orc.admm(g["P"].astype(np.float64), g["q"].astype(np.float64),
                 g["lb"].astype(np.float64), g["ub"].astype(np.float64))
    assert np.array_equal(o["x"], g["x_ref"])          # x returned by the reference's admm()
    assert np.array_equal(o["iters"], g["iters"])
    assert np.array_equal(o["status"], g["status"])
    assert np.array_equal(o["z"], g["z"])
    assert np.array_equal(o["u"], g["u"])


@pytest.mark.parametrize("path", golden_files(), ids=os.path.basename)
def test_oracle_objective_matches_reference_eval(orc, path):
    g = load_golden(path)
    for c in range(min(4, g["P"].shape[0])):
        f = orc.eval(g["P"][c].astype(np.float64), g["q"][c].astype(np.float64), 0.0,
                     g["x_ref"][c])
        assert f == g["f_ref"][c]


@pytest.mark.parametrize("n", [4, 16, 40])
def test_oracle_newton_matches_reference_newton(orc, n):
    g = load_golden(os.path.join(os.path.dirname(golden_files()[0]), f"wellcond_free_n{n}.npz"))
    for c in range(4):
        x, _ = orc.newton(g["P"][c].astype(np.float64), g["q"][c].astype(np.float64),
                          np.zeros(n), 10)
        assert np.array_equal(x, g["x_newton"][c])
        # inactive box: ADMM and Newton agree to the ADMM tolerance (the reference's
        # own three-way cross-check, test/test.c:89-106)
        assert np.max(np.abs(x - g["x_ref"][c])) <= 5e-2 * (1 + np.max(np.abs(x)))


def test_scalar_prod_known_answer(orc):
    """The reference's only deterministic known-answer test
    (test/test.c:59-87): d . (a b) == sum_{i<N} i^2, closed form at :77."""
    for n in (4, 16, 48):
        d = np.arange(n, dtype=np.float64)
        a = np.zeros((n, n))
        a[:, 0] = d
        b = np.zeros(n)
        b[0] = 1
        c = a @ b
        should_be = n * (n + 1) * (2 * n + 1) // 6 - n * n
        # eval with P = 0 computes q . x with the reference's dot-product order
        assert orc.eval(np.zeros((n, n)), d, 0.0, c) == should_be


def test_inversion_property(orc):
    """inversion_test (test/test.c:37-57): A * invert(A) ~ I within 1e-6."""
    rng = np.random.default_rng(3)
    for n in (4, 16, 48):
        B = rng.uniform(-1e3, 1e3, (n, n))
        A = B.T @ B / (1e3 * n)
        Ai, singular = orc.invert(A)
        assert singular == 0
        assert np.max(np.abs(A @ Ai - np.eye(n))) < 1e-6


def test_oracle_edge_cases(orc):
    n = 4
    P = np.eye(n)[None]
    q = np.ones((1, n))
    # iteration cap
    o = orc.admm(P, q, -10.0, 10.0, max_iter=1)
    assert o["status"][0] == 1 and o["iters"][0] == 1
    # singular P + rho I: the reference prints "singular matrix" (matrix_ops.c:511)
    o = orc.admm(-np.eye(n)[None], q, -10.0, 10.0)
    assert o["status"][0] == 2
    # fully clipped solution sits on the bound, dual sign encodes which bound
    o = orc.admm(P, 100 * q, -1.0, 1.0)
    assert np.all(o["z"][0] == -1.0) and np.all(o["lam"][0] < 0)
    # f32 twin runs the same trajectory on an easy problem
    o32 = orc.admm(P, q, -10.0, 10.0, dtype=np.float32)
    o64 = orc.admm(P, q, -10.0, 10.0)
    assert o32["iters"][0] == o64["iters"][0]
    assert np.allclose(o32["x"], o64["x"], atol=1e-6)
