#!/usr/bin/env python
"""Device-resident throughput of qp_solve_general_batch (rows l <= A x <= u, m = 2n).

    python tools/bench_general.py [--shapes 4,16,40] [--count 65536] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
from bench import make_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4,16,40")
    ap.add_argument("--count", type=int, default=1 << 16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--max-iter", type=int, default=10000, help="1: the cost of a solve without its iterations")
    ap.add_argument("--kernel", default="auto", choices=["auto", "smem", "reg"],
                    help="auto: registers where n <= 32, m <= 64 (qp_genr.cuh), else shared memory (qp_gen.cuh)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    opt = qp.Options(max_iter=args.max_iter, kernel={"auto": qp.KERNEL_AUTO, "smem": qp.KERNEL_GEN_SMEM, "reg": qp.KERNEL_GEN_REG}[args.kernel])
    for n in [int(s) for s in args.shapes.split(",")]:
        m, count = 2 * n, args.count
        P, q = make_batch(count, n, 2026, dev)
        gen = torch.Generator(device=dev).manual_seed(7)
        A = (torch.randn((count, m, n), generator=gen, device=dev) / n ** 0.5).contiguous()
        u = (0.5 * (0.2 + 0.8 * torch.rand((count, m), generator=gen, device=dev))).contiguous()
        l = (-u).contiguous()
        for _ in range(3):
            out = qp.solve_general_batch(P, q, A, l, u, options=opt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = qp.solve_general_batch(P, q, A, l, u, options=opt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"shape": f"n={n},m={m} general rows", "count": count, "kernel": args.kernel, "max_iter": args.max_iter,
                          "qp_per_s": count / (ms * 1e-3), "ms_per_step": ms,
                          "mean_iters": float(out["iters"].float().mean()),
                          "not_converged": int((out["status"] != 0).sum())}), flush=True)


if __name__ == "__main__":
    main()
