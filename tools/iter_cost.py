#!/usr/bin/env python
"""Cost of one ADMM iteration and of the setup (factorisation / inverse) per kernel family:
fixed iteration counts (abstol = reltol = 0, max_iter = K), device-resident batch, CUDA events.
Prints ms per launch for each K and the least-squares line  ms = setup + K * per_iteration."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
from bench import make_batch  # noqa: E402

SHAPES = {4: 1 << 20, 8: 1 << 20, 16: 1 << 20, 24: 1 << 18, 32: 1 << 18, 40: 1 << 18, 64: 1 << 16}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4,16,40,64")
    ap.add_argument("--ks", default="1,5,10,20,40,80")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kernel", default="auto")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    kern = {"auto": qp.KERNEL_AUTO, "tpq": qp.KERNEL_TPQ, "gpq": qp.KERNEL_GPQ, "bpq": qp.KERNEL_BPQ,
            "cpq": qp.KERNEL_CPQ}[args.kernel]
    ks = [int(k) for k in args.ks.split(",")]
    for n in [int(s) for s in args.shapes.split(",")]:
        count = SHAPES.get(n, 1 << 16)
        P, q = make_batch(count, n, 2026, dev)
        ms = []
        out = None
        for K in ks:
            o = qp.Options(kernel=kern, schedule=qp.SCHEDULE_SINGLE, abstol=0.0, reltol=0.0, max_iter=K)
            for _ in range(2):
                out = qp.solve_batch(P, q, -0.5, 0.5, out=out, options=o)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                qp.solve_batch(P, q, -0.5, 0.5, out=out, options=o)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / args.steps)
        A = np.stack([np.ones(len(ks)), np.array(ks, float)], axis=1)
        (setup, per_it), *_ = np.linalg.lstsq(A, np.array(ms), rcond=None)
        line = {"n": n, "count": count, "kernel": qp.kernel_for(n) if args.kernel == "auto" else args.kernel,
                "ks": ks, "ms": ms, "setup_ms": float(setup), "per_iteration_ms": float(per_it),
                "ns_per_problem_iteration": float(per_it) * 1e6 / count,
                "fma_frac_iteration": (n * n) * count / (per_it * 1e-3) / 36.2e12,
                "fma_frac_setup_n3_over_3": (n ** 3 / 3) * count / (setup * 1e-3) / 36.2e12}
        print(json.dumps(line), flush=True)
        del P, q, out


if __name__ == "__main__":
    main()
