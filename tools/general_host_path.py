#!/usr/bin/env python
"""qp_solve_general_batch with HOST pointers (one staging allocation, synchronous copies): QP/s at
n = 16, m = 32 with pageable and with pinned arrays, beside the device-resident rate."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
from bench import make_batch  # noqa: E402

n, m, count = 16, 32, 1 << 16
Pt, qt = make_batch(count, n, 2026, torch.device("cpu"))
rng = np.random.default_rng(7)
A = (rng.normal(size=(count, m, n)) / np.sqrt(n)).astype(np.float32)
u = (0.5 * rng.uniform(0.2, 1.0, (count, m))).astype(np.float32)
l = -u
P, q = Pt.numpy(), qt.numpy()


def best(fn, reps=5):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t)


page = best(lambda: qp.solve_general_batch(P, q, A, l, u))
pin = [torch.from_numpy(a).pin_memory() for a in (P, q, A, l, u)]
pout = {"x": torch.empty((count, n)).pin_memory(), "z": torch.empty((count, m)).pin_memory(),
        "y": torch.empty((count, m)).pin_memory(), "status": torch.empty(count, dtype=torch.int32).pin_memory(),
        "iters": torch.empty(count, dtype=torch.int32).pin_memory()}
pinned_in = best(lambda: qp.solve_general_batch(*pin))
pinned = best(lambda: qp.solve_general_batch(*pin, out=pout))
dev = [t.cuda() for t in pin]
def dv():
    qp.solve_general_batch(*dev); torch.cuda.synchronize()
device = best(dv)
bytes_in = (n * n + n + m * n + 2 * m) * 4
print(json.dumps({"shape": "n=16,m=32 general rows", "count": count, "input_bytes_per_problem": bytes_in,
                  "host_pageable_qp_per_s": count / page, "host_pinned_inputs_pageable_outputs_qp_per_s": count / pinned_in,
                  "host_pinned_qp_per_s": count / pinned,
                  "device_resident_qp_per_s": count / device,
                  "host_pinned_input_gb_per_s": count * bytes_in / pinned / 1e9}))
