#!/usr/bin/env python
"""Latency of small host batches through qp_solve_batch / qp_admm_qf (pageable numpy arrays), one
JSON line per shape; run it under QPB200_ZERO_COPY_BYTES=0 and =262144 to compare the one-copy
path (copy, launch, copy, wait) with the kernel working on the pinned staging buffers directly
(launch, wait).  --save FILE keeps the results for a bit-for-bit comparison of the two runs."""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
from bench import make_batch  # noqa: E402


def best(fn, reps):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t), float(np.median(t))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="16,40,4")
    ap.add_argument("--save", default="")
    args = ap.parse_args()
    keep = {}
    for n in [int(s) for s in args.shapes.split(",")]:
        Pt, qt = make_batch(256, n, 2026, torch.device("cpu"))
        P, q = Pt.numpy(), qt.numpy()
        qf = qp.QuadraticForm(P[0].astype(np.float64), q[0].astype(np.float64))
        qp.admm(None, 10000, qf, lb=-0.5, ub=0.5)
        t_qf = best(lambda: qp.admm(None, 10000, qf, lb=-0.5, ub=0.5), 300)
        rows = []
        for c in (1, 2, 4, 8, 16, 32, 64, 128, 256):
            out = {k: np.empty((c, n), np.float32) for k in ("x", "z", "lam")}
            out.update({k: np.empty(c, np.int32) for k in ("status", "iters")})
            qp.solve_batch(P[:c], q[:c], -0.5, 0.5, out=out)
            b, m = best(lambda: qp.solve_batch(P[:c], q[:c], -0.5, 0.5, out=out), 300)
            rows.append({"count": c, "best_us": b * 1e6, "median_us": m * 1e6})
            keep[f"n{n}_c{c}_x"] = out["x"].copy(); keep[f"n{n}_c{c}_it"] = out["iters"].copy()
            keep[f"n{n}_c{c}_lam"] = out["lam"].copy()
        print(json.dumps({"n": n, "QPB200_ZERO_COPY_BYTES": os.environ.get("QPB200_ZERO_COPY_BYTES"),
                          "qp_admm_qf_one_call_us": {"best": t_qf[0] * 1e6, "median": t_qf[1] * 1e6},
                          "rows": rows}), flush=True)
    if args.save:
        np.savez(args.save, **keep)


if __name__ == "__main__":
    main()
