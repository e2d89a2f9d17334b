#!/usr/bin/env python
"""Offline study behind the straggler classifier of qp_bpq.cuh (PASS_FIRST): a float64 numpy
restatement of the box ADMM on the 4096 condensed-MPC problems (BASELINE config 4), the
residuals recorded at small iteration counts, and how well simple rules evaluated when a
window of W iterations closes rank the problems that will run longest.  No GPU.

    python tools/mpc_straggler_predictor.py > profiles/r2_mpc_straggler_predictor.txt
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embedded_qp_solver_b200 import problems  # noqa: E402

P, q, lb, ub = problems.mpc(4096, seed=2026)
P = P.astype(np.float64); q = q.astype(np.float64); beta = float(ub[0, 0])
N = len(P); n = 40; rho = 1.0; abstol = 1e-4; reltol = 1e-2   # qp_solvers.c:17-18, :281
R = np.linalg.inv(P + rho * np.eye(n))
z = np.zeros((N, n)); u = np.zeros((N, n))
iters = np.zeros(N, int); done = np.zeros(N, bool)
hist = {}
for k in range(1, 3000):
    x = np.einsum('bij,bj->bi', R, rho * (z - u) - q)
    zn = np.clip(x + u, -beta, beta)
    u = u + x - zn
    rr = ((x - zn) ** 2).sum(1); ss = rho * rho * ((zn - z) ** 2).sum(1)
    ep = (np.sqrt(n) * abstol + reltol * np.sqrt(np.maximum((x ** 2).sum(1), (zn ** 2).sum(1)))) ** 2
    ed = (np.sqrt(n) * abstol + reltol * rho * np.sqrt((u ** 2).sum(1))) ** 2
    z = zn
    conv = (rr < ep) & (ss < ed)
    newly = conv & ~done
    iters[newly] = k; done |= conv
    if k <= 33:
        hist[k] = (rr.copy(), ss.copy(), ep.copy(), ed.copy())
    if done.all():
        break
top = np.argsort(-iters)[:12]
print('iterations: mean', iters.mean(), 'max', iters.max(), 'twelve longest', iters[top].tolist())
print('\n--- rank (0 = first) that each rule gives the eight longest problems among those still running at W')
for W in (3, 4, 6, 8, 12, 16, 24):
    rr0, ss0, _, _ = hist[W - 1]; rr1, ss1, ep, ed = hist[W]
    with np.errstate(all='ignore'):
        s = {}
        s['decay of ||x-z||^2'] = np.nan_to_num(rr1 / rr0, nan=0)
        s['decay of ||z-z_old||^2'] = np.nan_to_num(ss1 / ss0, nan=0)
        lp = np.log(np.maximum(rr1 / ep, 1.0)); ld = np.log(np.maximum(ss1 / ed, 1.0))
        dp = np.log(np.maximum(rr0 / rr1, 1.0000001)); dd = np.log(np.maximum(ss0 / ss1, 1.0000001))
        s['iterations to go, primal'] = lp / dp; s['iterations to go, dual'] = ld / dd
        s['||x-z||^2 / eps_pri^2'] = rr1 / ep; s['rho^2 ||dz||^2 / eps_dual^2'] = ss1 / ed
    ev = iters > W
    for name, sc in s.items():
        sc = np.where(ev, sc, -1)
        order = np.argsort(-sc)
        rank = np.empty(N, int); rank[order] = np.arange(N)
        print(f'W={W:2d} {name:30s}', rank[top[:8]].tolist(), 'of', int(ev.sum()))
print('\n--- the rule in the kernel: flagged when ||x-z||^2 > eps_pri^2 (previous/current)^T, float32')
for W in (2, 3, 4, 6, 8, 12, 16, 24):
    rr0, ss0, _, _ = hist[W - 1]; rr1, ss1, ep, ed = hist[W]
    ev = iters > W; rem = iters - W
    for T in (32, 64, 128):
        with np.errstate(all='ignore'):
            d = (np.float32(rr0) / np.float32(rr1)).astype(np.float32)
            for _ in range(int(np.log2(T))):
                d = d * d
            flag = (np.float32(rr1) > np.float32(ep) * d) & ev
        print(f'W={W:2d} T={T:3d} still running {int(ev.sum()):4d} flagged {int(flag.sum()):4d} '
              f'of the eight longest flagged {int(flag[top[:8]].sum())} longest one missed needs {int(rem[ev & ~flag].max())} more '
              f'mean to go: flagged {rem[flag].mean():.1f} others {rem[ev & ~flag].mean():.1f}')
