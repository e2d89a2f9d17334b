#!/usr/bin/env python
"""Small solves of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
a few hundred problems each, including a PASSES solve and the host-memory path.
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
from embedded_qp_solver_b200 import problems  # noqa: E402


def main():
    fams = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tpq", "bpq", "gpq", "cpq", "passes", "host"]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    shapes = {"tpq": [(16, 300), (4, 300), (8, 100), (12, 100)], "bpq": [(40, 40), (24, 40)],
              "gpq": [(64, 20), (20, 40)], "cpq": [(96, 4)]}
    for fam in fams:
        if fam in shapes:
            for n, count in shapes[fam]:
                P, q, _, _ = problems.wellcond(count, n, seed=1, beta=0.5)
                r = qp.solve_batch(dev(P), dev(q), -0.5, 0.5, options=qp.Options(schedule=qp.SCHEDULE_SINGLE))
                torch.cuda.synchronize()
                print(fam, n, count, "iters", float(r["iters"].float().mean()), flush=True)
        elif fam == "passes":
            for n, count in ((16, 200), (40, 30)):
                P, q, _, _ = problems.refdist(count, n, seed=2, box=1e4)
                r = qp.solve_batch(dev(P), dev(q), -1e4, 1e4,
                                   options=qp.Options(schedule=qp.SCHEDULE_PASSES, max_iter=200))
                torch.cuda.synchronize()
                print("passes", n, count, "iters", float(r["iters"].float().mean()), flush=True)
        elif fam == "host":
            P, q, _, _ = problems.wellcond(3000, 16, seed=3, beta=0.5)
            r = qp.solve_batch(P, q, -0.5, 0.5, options=qp.Options(chunk=1024))
            print("host", 16, 3000, "iters", float(r["iters"].mean()), flush=True)


if __name__ == "__main__":
    main()
