#!/usr/bin/env python
"""Which kernel family for a handful of problems?  Wall clock of qp_solve_batch on pageable numpy
arrays, counts 1 ... 64, kernel AUTO / TPQ / GPQ / BPQ forced (one JSON line per shape)."""
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
    ap.add_argument("--shapes", default="16,8,4,12")
    args = ap.parse_args()
    for n in [int(s) for s in args.shapes.split(",")]:
        Pt, qt = make_batch(64, n, 2026, torch.device("cpu"))
        P, q = Pt.numpy(), qt.numpy()
        rows = []
        for c in (1, 2, 4, 8, 16, 32, 64):
            row = {"count": c}
            for name, k in (("auto", qp.KERNEL_AUTO), ("tpq", qp.KERNEL_TPQ), ("gpq", qp.KERNEL_GPQ), ("bpq", qp.KERNEL_BPQ)):
                o = qp.Options(kernel=k)
                try:
                    qp.solve_batch(P[:c], q[:c], -0.5, 0.5, options=o)
                except qp.QpError:
                    continue
                b, m = best(lambda: qp.solve_batch(P[:c], q[:c], -0.5, 0.5, options=o), 200)
                row[name] = [round(b * 1e6, 1), round(m * 1e6, 1)]
            rows.append(row)
        print(json.dumps({"n": n, "us_best_median": rows}), flush=True)


if __name__ == "__main__":
    main()
