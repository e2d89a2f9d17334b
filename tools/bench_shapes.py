#!/usr/bin/env python
"""Device-resident throughput of every BASELINE.json shape (one JSON line each).

    python tools/bench_shapes.py [--steps K] [--shapes 4,16,40,64]

configs: 1M n=4 | 1M n=16 (headline, also bench.py) | 256K n=40 (condensed-MPC-like
dense box QPs) | 64K n=64 warm-started from a perturbed previous solve.
Same timing rules as bench.py (CUDA events on the launching stream, >= 3 warm-ups,
inputs larger than L2 or noted).
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

SHAPES = {4: 1 << 20, 16: 1 << 20, 40: 1 << 18, 64: 1 << 16}


def make_refdist(count, n, seed, device):
    """matirx_random_pos_def + matrix_random of the reference (matrix_ops.c:683-734) with the
    ranges of config.h:16-20, on the device."""
    gen = torch.Generator(device=device).manual_seed(seed)
    P = torch.empty((count, n, n), device=device)
    q = (torch.rand((count, n), generator=gen, device=device) * 2 - 1) * 1e3
    step = 1 << 17
    for lo in range(0, count, step):
        hi = min(count, lo + step)
        B = (torch.rand((hi - lo, n, n), generator=gen, device=device) * 2 - 1) * 1e3
        Pc = torch.bmm(B.transpose(1, 2), B) / (1e3 * n)
        P[lo:hi] = 0.5 * (Pc + Pc.transpose(1, 2))
    return P.contiguous(), q.contiguous()


def refdist(args, dev, kern, fp32_peak):
    counts = {4: 1 << 19, 8: 1 << 18, 16: 1 << 18, 24: 1 << 16, 32: 1 << 16, 40: 1 << 16}
    for item in filter(None, args.counts.split(",")):
        k, v = item.split(":")
        counts[int(k)] = int(v)
    for n in [int(s) for s in args.shapes.split(",")]:
        count = counts.get(n, 1 << 15)
        P, q = make_refdist(count, n, 2026, dev)
        res = {}
        for name, sch in (("single", qp.SCHEDULE_SINGLE), ("passes", qp.SCHEDULE_PASSES)):
            o = qp.Options(kernel=kern, schedule=sch)
            out = qp.solve_batch(P, q, -1e4, 1e4, options=o)
            torch.cuda.synchronize()
            before = sum(qp.schedule_stats()[k] for k in ("outlived", "evicted_later"))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                qp.solve_batch(P, q, -1e4, 1e4, out=out, options=o)
            e1.record()
            torch.cuda.synchronize()
            after = sum(qp.schedule_stats()[k] for k in ("outlived", "evicted_later"))
            res[name] = dict(ms=e0.elapsed_time(e1) / args.steps,
                             evictions=(after - before) // args.steps,
                             out={k: v.clone() for k, v in out.items()})
        it = res["single"]["out"]["iters"].float()
        same = all(torch.equal(res["single"]["out"][k], res["passes"]["out"][k]) for k in res["single"]["out"])
        hist = torch.histc(torch.log2(it.clamp(min=1)), bins=14, min=0, max=14).int().tolist()
        w = 32 if n <= 16 else 3
        itw = it[: count // w * w].view(-1, w)
        flops = qp.algorithmic_flops(n, float(it.mean()))
        line = {"shape": f"n={n},m={2 * n} refdist box 1e4", "count": count,
                "kernel": qp.kernel_for(n) if args.kernel == "auto" else f"{args.kernel}<{n}>",
                "mean_iters": float(it.mean()), "median_iters": float(it.median()), "max_iters": int(it.max()),
                "hit_cap": int((res["single"]["out"]["status"] == 1).sum()),
                "iters_log2_histogram": hist,
                "lane_efficiency_one_level": float(it.mean() / itw.max(dim=1).values.mean()),
                "single": {"ms_per_step": res["single"]["ms"], "qp_per_s": count / res["single"]["ms"] * 1e3},
                "passes": {"ms_per_step": res["passes"]["ms"], "qp_per_s": count / res["passes"]["ms"] * 1e3,
                           "evictions_per_step": res["passes"]["evictions"],
                           "fp32_frac_reference_flops": flops * count / res["passes"]["ms"] * 1e3 / 1e12 / fp32_peak},
                "speedup": res["single"]["ms"] / res["passes"]["ms"], "bit_identical": bool(same)}
        print(json.dumps(line), flush=True)
        del P, q, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--shapes", default="4,16,40,64")
    ap.add_argument("--beta", type=float, default=0.5)
    ap.add_argument("--cold", action="store_true", help="no warm start at n=64")
    ap.add_argument("--mpc", action="store_true",
                    help="n=40: condensed linear-MPC problems (BASELINE config 4; 4096 distinct, tiled)")
    ap.add_argument("--refdist", action="store_true",
                    help="the reference's own test distribution (matrix_ops.c:699-734: B, q ~ U(+-1e3), "
                         "P = B'B/(1e3 n)), box +-1e4 (config_board.h), max_iter 10000: heavy-tailed "
                         "iteration counts; prints SINGLE and PASSES schedules and an iteration histogram")
    ap.add_argument("--schedule", default="auto", choices=["auto", "single", "passes"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "tpq", "gpq", "bpq", "cpq"])
    ap.add_argument("--max-iter", type=int, default=10000, help="iteration cap (qp_options.max_iter)")
    ap.add_argument("--share", type=int, default=1,
                    help="qp_batch.p_share: every K consecutive problems use the matrix of the first of them")
    ap.add_argument("--counts", default="", help="override batch sizes, e.g. 24:262144,32:262144")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    fp32_peak = qp.probe_fp32_tflops(0, 5)
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    kern = {"auto": qp.KERNEL_AUTO, "tpq": qp.KERNEL_TPQ, "gpq": qp.KERNEL_GPQ, "bpq": qp.KERNEL_BPQ, "cpq": qp.KERNEL_CPQ}[args.kernel]
    sched = {"auto": qp.SCHEDULE_AUTO, "single": qp.SCHEDULE_SINGLE, "passes": qp.SCHEDULE_PASSES}[args.schedule]
    opts = qp.Options(kernel=kern, schedule=sched, max_iter=args.max_iter)
    if args.refdist:
        return refdist(args, dev, kern, fp32_peak)
    counts = dict(SHAPES)
    for item in filter(None, args.counts.split(",")):
        k, v = item.split(":")
        counts[int(k)] = int(v)
    for n in [int(s) for s in args.shapes.split(",")]:
        # shapes outside BASELINE.json: about 1.5 GB of P
        count = counts.get(n) or min(1 << 20, 1 << (((3 << 27) // (n * n)).bit_length() - 1))
        beta = args.beta
        if args.mpc and n == 40:
            from embedded_qp_solver_b200 import problems
            Pm, qm, lbm, ubm = problems.mpc(4096, seed=2026)
            reps = count // 4096
            P = torch.from_numpy(Pm).to(dev).repeat(reps, 1, 1).contiguous()
            q = torch.from_numpy(qm).to(dev).repeat(reps, 1).contiguous()
            beta = float(ubm[0, 0])
        else:
            P, q = make_batch(count, n, 2026, dev)
        warm = n == 64 and not args.cold
        kw = {}
        if args.share > 1:
            P = P[::args.share].contiguous()
            kw["p_share"] = args.share
        if warm:
            first = qp.solve_batch(P, q, -beta, beta, options=opts)
            q = q + 0.01 * torch.randn_like(q)
            kw = dict(z0=first["z"].clone(), u0=first["lam"].clone())
        out = None
        for _ in range(3):
            out = qp.solve_batch(P, q, -beta, beta, out=out, options=opts, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            qp.solve_batch(P, q, -beta, beta, out=out, options=opts, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        k = float(out["iters"].float().mean())
        bytes_per = qp.algorithmic_bytes(n, warm=warm, with_z=True, bound_arrays=False)
        flops_per = qp.algorithmic_flops(n, k)
        rate = count / (ms * 1e-3)
        line = {"shape": f"n={n},m={2 * n}" + (" mpc" if args.mpc and n == 40 else ""), "count": count, "kernel": qp.kernel_for(n) if args.kernel == "auto" else f"{args.kernel}<{n}>",
                "warm_start": warm, "schedule": args.schedule, "p_share": args.share, "qp_per_s": rate, "ms_per_step": ms, "mean_iters": k,
                "not_converged": int((out["status"] != 0).sum()),
                "hbm": {"achieved_gbs": bytes_per * rate / 1e9, "peak": hbm_peak,
                        "frac": bytes_per * rate / 1e9 / hbm_peak, "bytes_per_problem": bytes_per},
                "fp32": {"achieved_tflops": flops_per * rate / 1e12, "peak": fp32_peak,
                         "frac": flops_per * rate / 1e12 / fp32_peak,
                         "flops_per_problem": flops_per},
                "input_mb": (n * n + n) * 4 * count / 1e6}
        print(json.dumps(line), flush=True)
        del P, q, out


if __name__ == "__main__":
    main()
