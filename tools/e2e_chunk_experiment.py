#!/usr/bin/env python
"""End-to-end rate of the host path (n = 16, 2^20 problems, pinned buffers) against the chunk
size (qp_options.chunk).  One JSON line per chunk size; QPB200_PACK_CACHED / QPB200_HOST_PACK
are read by the library."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
from bench import make_batch  # noqa: E402

n, count = 16, 1 << 20
dev = torch.device("cuda", 0)
P, q = make_batch(count, n, 2026, dev)[:2]
hP, hq = P.cpu().pin_memory(), q.cpu().pin_memory()
hout = {k: torch.empty((count, n)).pin_memory() for k in ("x", "z", "lam")}
hout.update({k: torch.empty(count, dtype=torch.int32).pin_memory() for k in ("status", "iters")})
for chunk in [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "0,4096,8192,16384").split(",")]:
    opts = qp.Options(device=0, chunk=chunk)
    for _ in range(2):
        qp.solve_batch(hP, hq, -0.5, 0.5, options=opts, out=hout)
    w0 = qp.host_path_h2d_bytes()
    t0 = time.perf_counter()
    for _ in range(10):
        qp.solve_batch(hP, hq, -0.5, 0.5, options=opts, out=hout)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(json.dumps({"chunk": chunk, "qp_per_s": count / dt, "ms": dt * 1e3,
                      "wire_bytes": (qp.host_path_h2d_bytes() - w0) // 10,
                      "cached": bool(os.environ.get("QPB200_PACK_CACHED"))}), flush=True)
