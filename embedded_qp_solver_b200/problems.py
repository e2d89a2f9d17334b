"""Seeded synthetic box-QP generators (host side, numpy).

Distributions follow SURVEY.md 8(d):

* ``wellcond``  -- the benchmark distribution "strictly-convex random":
  ``G ~ N(0,1)^{n x n}``, ``P = G^T G / n + I`` (exactly symmetric),
  ``q ~ N(0,1)^n``, box ``[-beta, +beta]``.
* ``refdist``   -- the reference's own test distribution
  (``matirx_random_pos_def``, /root/reference/matrix/matrix_ops/matrix_ops.c:699-734
  with the ranges of /root/reference/config.h:16-20): ``B, q ~ U(-1e3, 1e3)``,
  ``P = B^T B / (1e3 n)``.  Ill-conditioned: a stress set, not a benchmark.
* ``mpc``       -- condensed linear-MPC QPs, n = horizon x inputs.
* ``perturb_q`` -- the warm-start scenario (config 5): same P, q perturbed.

Every generator is a pure function of (seed, count, n): a shard
``[lo, hi)`` of the batch is reproduced by generating with ``first=lo``.
Problems are drawn per index from ``numpy.random.Philox(key=seed, counter=index)``
so shards on different ranks agree with the whole batch.
"""
from __future__ import annotations

import numpy as np


def _rng(seed: int, index: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, index]))


def wellcond(count: int, n: int, seed: int = 2026, beta: float = 0.5, first: int = 0,
             dtype=np.float32):
    P = np.empty((count, n, n), np.float64)
    q = np.empty((count, n), np.float64)
    for c in range(count):
        g = _rng(seed, first + c)
        G = g.standard_normal((n, n))
        P[c] = G.T @ G / n + np.eye(n)
        q[c] = g.standard_normal(n)
    P = 0.5 * (P + P.transpose(0, 2, 1))
    lb = np.full((count, n), -beta)
    ub = np.full((count, n), beta)
    return _cast(P, q, lb, ub, dtype)


def refdist(count: int, n: int, seed: int = 2026, box: float = 1e12, first: int = 0,
            dtype=np.float32):
    P = np.empty((count, n, n), np.float64)
    q = np.empty((count, n), np.float64)
    for c in range(count):
        g = _rng(seed, first + c)
        B = g.uniform(-1e3, 1e3, (n, n))
        P[c] = B.T @ B / (1e3 * n)
        q[c] = g.uniform(-1e3, 1e3, n)
    P = 0.5 * (P + P.transpose(0, 2, 1))
    lb = np.full((count, n), -box)
    ub = np.full((count, n), box)
    return _cast(P, q, lb, ub, dtype)


def mpc(count: int, horizon: int = 10, n_in: int = 4, n_state: int = 8, seed: int = 2026,
        first: int = 0, du_max: float = 0.2, dtype=np.float32):
    """Condensed linear MPC in the input-increment parameterisation.

    x+ = A x + B u with u_k = u_{k-1} + du_k; decision vector = stacked du
    (n = horizon * n_in), cost sum_k x_k^T Q x_k + du_k^T Rw du_k, constraints
    |du| <= du_max: a pure box, so m = 2 n (BASELINE config 4).
    """
    n = horizon * n_in
    P = np.empty((count, n, n), np.float64)
    q = np.empty((count, n), np.float64)
    for c in range(count):
        g = _rng(seed, first + c)
        A = g.standard_normal((n_state, n_state))
        A *= 0.9 / max(1e-9, np.max(np.abs(np.linalg.eigvals(A))))
        B = g.standard_normal((n_state, n_in)) / np.sqrt(n_state)
        x0 = g.standard_normal(n_state)
        # augmented state [x; u_prev], input du
        Aa = np.block([[A, B], [np.zeros((n_in, n_state)), np.eye(n_in)]])
        Ba = np.vstack([B, np.eye(n_in)])
        na = n_state + n_in
        Gam = np.zeros((horizon * na, n))
        Phi = np.zeros((horizon * na, na))
        Ak = np.eye(na)
        for k in range(horizon):
            Ak = Aa @ Ak
            Phi[k * na:(k + 1) * na] = Ak
        for k in range(horizon):
            for j in range(k + 1):
                Gam[k * na:(k + 1) * na, j * n_in:(j + 1) * n_in] = \
                    np.linalg.matrix_power(Aa, k - j) @ Ba
        Qa = np.zeros((na, na))
        Qa[:n_state, :n_state] = np.eye(n_state)
        Qbar = np.kron(np.eye(horizon), Qa)
        Rw = 0.1 * np.eye(n)
        xa0 = np.concatenate([x0, np.zeros(n_in)])
        P[c] = Gam.T @ Qbar @ Gam + Rw
        q[c] = Gam.T @ Qbar @ (Phi @ xa0)
    P = 0.5 * (P + P.transpose(0, 2, 1))
    lb = np.full((count, n), -du_max)
    ub = np.full((count, n), du_max)
    return _cast(P, q, lb, ub, dtype)


def perturb_q(q: np.ndarray, seed: int = 7, rel: float = 0.01, first: int = 0):
    """q' = q + rel * N(0,1): the "perturbed previous solution" warm-start case."""
    out = np.array(q, dtype=np.float64)
    for c in range(out.shape[0]):
        out[c] += rel * _rng(seed, first + c).standard_normal(out.shape[1])
    return out.astype(q.dtype)


def _cast(P, q, lb, ub, dtype):
    return (np.ascontiguousarray(P, dtype), np.ascontiguousarray(q, dtype),
            np.ascontiguousarray(lb, dtype), np.ascontiguousarray(ub, dtype))


def shard_bounds(count: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of a batch for `rank`; remainder to the first
    ranks."""
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi
