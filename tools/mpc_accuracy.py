"""x error of every kernel family against the FP64 oracle on the condensed-MPC set (n = 40)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
import oracle  # noqa: E402
from embedded_qp_solver_b200 import problems  # noqa: E402


def main():
    oracle.build(ref=False)
    orc = oracle.Oracle()
    P, q, lb, ub = problems.mpc(4096, seed=2026)
    beta = float(ub[0, 0])
    o = orc.admm(P.astype(np.float64), q.astype(np.float64), -beta, beta)
    o32 = orc.admm(P, q, -beta, beta, dtype=np.float32)
    tP, tq = torch.from_numpy(P).cuda(), torch.from_numpy(q).cuda()

    def rel(x):
        return np.max(np.abs(x - o["x"]), axis=1) / (1 + np.max(np.abs(o["x"]), axis=1))

    e32 = rel(o32["x"].astype(np.float64))
    s32 = o32["iters"] == o["iters"]
    print(json.dumps({"family": "oracle_f32", "same_iters": float(s32.mean()),
                      "max": float(e32[s32].max()), "over_1e-4": int((e32[s32] > 1e-4).sum())}))
    for name, k in (("bpq", qp.KERNEL_BPQ), ("gpq", qp.KERNEL_GPQ), ("cpq", qp.KERNEL_CPQ)):
        r = qp.solve_batch(tP, tq, -beta, beta, options=qp.Options(kernel=k))
        torch.cuda.synchronize()
        x = r["x"].cpu().numpy().astype(np.float64)
        it = r["iters"].cpu().numpy()
        same = it == o["iters"]
        e = rel(x)
        worst = np.argsort(-np.where(same, e, 0))[:5]
        print(json.dumps({"family": name, "same_iters": float(same.mean()), "max": float(e[same].max()),
                          "p99": float(np.quantile(e[same], 0.99)), "over_1e-4": int((e[same] > 1e-4).sum()),
                          "worst": [(int(i), float(e[i]), int(it[i])) for i in worst]}))


if __name__ == "__main__":
    main()
