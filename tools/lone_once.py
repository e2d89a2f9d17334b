#!/usr/bin/env python
"""One solve of the slowest condensed-MPC problem (three copies: one warp), for ncu."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp
from embedded_qp_solver_b200 import problems
Pm, qm, lbm, ubm = problems.mpc(4096, seed=2026)
dev = torch.device("cuda", 0)
allP = torch.from_numpy(Pm).to(dev); allq = torch.from_numpy(qm).to(dev)
i = int(qp.solve_batch(allP, allq, -float(ubm[0, 0]), float(ubm[0, 0]),
                       options=qp.Options(schedule=qp.SCHEDULE_SINGLE))["iters"].argmax())
P = torch.from_numpy(Pm[i:i + 1]).to(dev).repeat(3, 1, 1).contiguous()
q = torch.from_numpy(qm[i:i + 1]).to(dev).repeat(3, 1).contiguous()
beta = float(ubm[0, 0])
sch = qp.SCHEDULE_PASSES if sys.argv[1] == "passes" else qp.SCHEDULE_SINGLE
for _ in range(2):
    out = qp.solve_batch(P, q, -beta, beta, options=qp.Options(schedule=sch))
torch.cuda.synchronize()
print(int(out["iters"][0]))
