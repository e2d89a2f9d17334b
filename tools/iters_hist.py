#!/usr/bin/env python
"""Iteration-count statistics of the condensed-MPC set (BASELINE config 4) from a GPU solve."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp
from embedded_qp_solver_b200 import problems
Pm, qm, lbm, ubm = problems.mpc(4096, seed=2026)
dev = torch.device("cuda", 0)
P = torch.from_numpy(Pm).to(dev); q = torch.from_numpy(qm).to(dev)
beta = float(ubm[0, 0])
out = qp.solve_batch(P, q, -beta, beta, options=qp.Options(schedule=qp.SCHEDULE_SINGLE))
it = out["iters"].float().cpu()
qs = [0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.95, 0.99, 0.999]
res = {"mean": float(it.mean()), "max": int(it.max()), "min": int(it.min()),
       "quantiles": {str(x): float(torch.quantile(it, x)) for x in qs}}
for w in (8, 16, 24, 32, 48, 64, 128, 256, 512, 1024):
    res[f"frac_gt_{w}"] = float((it > w).float().mean())
    res[f"mean_excess_over_{w}"] = float((it - w).clamp(min=0).mean())
it3 = it[: 4095].view(-1, 3)
res["mean_of_max3"] = float(it3.max(dim=1).values.mean())
res["top10"] = sorted(it.int().tolist())[-10:]
print(json.dumps(res))
