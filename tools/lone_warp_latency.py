#!/usr/bin/env python
"""Time per ADMM iteration of ONE warp that has an SM to itself (the critical path of a straggler):
the slowest problem of the condensed-MPC set, three copies, single launch and pass schedule."""
import json, os, sys, time
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
i = int(out["iters"].argmax()); K = int(out["iters"][i])
for copies in (1, 3, 3 * 148):
    Pc = P[i:i + 1].repeat(copies, 1, 1).contiguous(); qc = q[i:i + 1].repeat(copies, 1).contiguous()
    for name, sch in (("single", qp.SCHEDULE_SINGLE), ("passes", qp.SCHEDULE_PASSES)):
        o = qp.Options(schedule=sch)
        r = qp.solve_batch(Pc, qc, -beta, beta, options=o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            qp.solve_batch(Pc, qc, -beta, beta, out=r, options=o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"copies": copies, "schedule": name, "iters": K, "ms": ms, "us_per_iteration": ms * 1e3 / K}))
