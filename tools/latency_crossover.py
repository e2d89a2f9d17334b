#!/usr/bin/env python
"""Per-call latency of the drop-in and the batch size at which the GPU passes the CPU.

What it replaces: one admm() call per problem, ~25 us at n = 16 on one core
(/root/reference/test/test.c:89-106 prints its time).  Measured here, all through the C ABI
with HOST memory (a C host's arrays: pageable numpy, and pinned for comparison):
  * qp_admm_qf        one struct _quadratic_form (double), count = 1
  * qp_solve_batch    count = 1 ... 2^20, wall clock per call, best of `reps`
  * the unmodified reference (oracle/_ref) on one core and on all cores (one forked process
    per core) for the same counts
and the smallest count at which the GPU call is faster than 1 core / all cores.
One JSON line per shape."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
import oracle  # noqa: E402
from bench import make_batch  # noqa: E402

_DATA = None


def _work(span):
    a, b = span
    n, P, q = _DATA
    x, secs = oracle.Reference(n).admm_many(P[a:b], q[a:b], -0.5, 0.5, 10000)
    return secs


def best(fn, reps):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t), float(np.median(t))


def main():
    global _DATA
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="16,40")
    ap.add_argument("--max-log2", type=int, default=20)
    args = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    for n in [int(s) for s in args.shapes.split(",")]:
        top = 1 << min(args.max_log2, 20 if n <= 16 else 17)
        Pt, qt = make_batch(top, n, 2026, torch.device("cpu"))
        P, q = Pt.numpy(), qt.numpy()
        P64, q64 = P.astype(np.float64), q.astype(np.float64)
        hP, hq = Pt.pin_memory(), qt.pin_memory()
        # --- one struct through the drop-in entry (double in, double out)
        qf = qp.QuadraticForm(P64[0], q64[0])
        qp.admm(None, 10000, qf, lb=-0.5, ub=0.5)
        t_qf = best(lambda: qp.admm(None, 10000, qf, lb=-0.5, ub=0.5), 200)
        # --- the reference, one core
        ref = oracle.Reference(n) if oracle.ref_available(n) else None
        rows = []
        counts = [1 << k for k in range(0, top.bit_length())]
        _DATA = (n, P64, q64)
        pool = mp.get_context("fork").Pool(cores) if ref else None
        for c in counts:
            reps = 50 if c <= 4096 else (10 if c <= 65536 else 3)
            out = {k: np.empty((c, n), np.float32) for k in ("x", "z", "lam")}
            out.update({k: np.empty(c, np.int32) for k in ("status", "iters")})
            qp.solve_batch(P[:c], q[:c], -0.5, 0.5, out=out)
            g_page = best(lambda: qp.solve_batch(P[:c], q[:c], -0.5, 0.5, out=out), reps)
            pout = {k: torch.empty((c, n)).pin_memory() for k in ("x", "z", "lam")}
            pout.update({k: torch.empty(c, dtype=torch.int32).pin_memory() for k in ("status", "iters")})
            qp.solve_batch(hP[:c], hq[:c], -0.5, 0.5, out=pout)
            g_pin = best(lambda: qp.solve_batch(hP[:c], hq[:c], -0.5, 0.5, out=pout), reps)
            row = {"count": c, "gpu_pageable_s": g_page[0], "gpu_pinned_s": g_pin[0],
                   "gpu_pageable_median_s": g_page[1]}
            if ref:
                cc = min(c, 1 << 14)  # one core: extrapolate beyond 16 K problems
                row["cpu_1core_s"] = ref.admm_many(P64[:cc], q64[:cc], -0.5, 0.5, 10000)[1] * c / cc
                if c >= cores:
                    edges = np.linspace(0, c, cores + 1).astype(int)
                    jobs = [(int(a), int(b)) for a, b in zip(edges, edges[1:]) if b > a]
                    t0 = time.perf_counter()
                    pool.map(_work, jobs, chunksize=1)
                    row["cpu_allcores_s"] = time.perf_counter() - t0
                else:
                    row["cpu_allcores_s"] = row["cpu_1core_s"] / max(1, c)  # one problem per core
            rows.append(row)
        if pool:
            pool.close()
            pool.join()
        cross1 = next((r["count"] for r in rows if ref and r["gpu_pageable_s"] < r["cpu_1core_s"]), None)
        crossn = next((r["count"] for r in rows if ref and r["gpu_pageable_s"] < r["cpu_allcores_s"]), None)
        print(json.dumps({"n": n, "host_cores": cores, "qp_admm_qf_one_call_s": {"best": t_qf[0], "median": t_qf[1]},
                          "reference_one_call_s": rows[0].get("cpu_1core_s"),
                          "crossover_count_vs_1_core": cross1, "crossover_count_vs_all_cores": crossn,
                          "rows": rows}), flush=True)


if __name__ == "__main__":
    main()
