"""CPU: the general-rows restatement (oracle/general_np.py) reduces to the reference's
box ADMM for A = I -- checked against the oracle that is bit-identical to the unmodified
reference, and against the reference's own golden vectors."""
import os

import numpy as np

from conftest import golden_files, load_golden
from oracle import general_np


def test_general_restatement_is_the_reference_admm_for_identity_rows(orc):
    g = load_golden(os.path.join(os.path.dirname(golden_files()[0]), "wellcond_b05_n16.npz"))
    P, q = g["P"][:12].astype(np.float64), g["q"][:12].astype(np.float64)
    lb, ub = g["lb"][:12].astype(np.float64), g["ub"][:12].astype(np.float64)
    eye = np.broadcast_to(np.eye(16), (12, 16, 16))
    o = general_np.admm_general_batch(P, q, eye, lb, ub)
    assert np.array_equal(o["iters"], g["iters"][:12]) and np.array_equal(o["status"], g["status"][:12])
    # same algorithm, numpy's summation order: equal to rounding, not bit for bit
    assert np.max(np.abs(o["x"] - g["x_ref"][:12])) <= 1e-12
    assert np.max(np.abs(o["z"] - g["z"][:12])) <= 1e-12 and np.max(np.abs(o["u"] - g["u"][:12])) <= 1e-12


def test_general_restatement_satisfies_kkt_on_polytope_rows():
    rng = np.random.default_rng(11)
    n, m = 6, 12
    G = rng.normal(size=(n, n))
    P = G.T @ G / n + np.eye(n)
    q = rng.normal(size=n)
    A = rng.normal(size=(m, n))
    l, u = -0.3 * np.ones(m), 0.4 * np.ones(m)
    o = general_np.admm_general(P, q, A, l, u, abstol=1e-10, reltol=1e-10, max_iter=200000)
    x, y = o["x"], o["u"]  # rho = 1: multipliers = u
    assert o["status"] == 0
    assert np.max(np.abs(P @ x + q + A.T @ y)) <= 1e-7
    ax = A @ x
    assert np.all(ax >= l - 1e-7) and np.all(ax <= u + 1e-7)
    inside = (ax > l + 1e-6) & (ax < u - 1e-6)
    assert np.max(np.abs(y[inside]), initial=0.0) <= 1e-7
    assert np.all(y[np.abs(ax - u) < 1e-6] >= -1e-9) and np.all(y[np.abs(ax - l) < 1e-6] <= 1e-9)
