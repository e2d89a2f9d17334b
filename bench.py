#!/usr/bin/env python
"""Headline benchmark: batched QP solves/sec at n=16, m=32 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU)

A "step" is one pass of the hot path over one batch of `--count` (default 2^20)
synthetic strictly-convex box QPs per GPU (weak scaling: every rank owns its own
shard, no collective on the data path).

One JSON line on rank 0:
  value      whole-job QP/s with the batch resident in HBM, CUDA-event timed on the
             launching stream, max over ranks
  e2e        the same metric through the public C-ABI call with HOST (pinned)
             buffers: H2D of all inputs and D2H of all outputs inside the timed region
  roofline   the binding bound of the dominant (only) kernel: FP32 flops of the
             reference algorithm F = 8/3 n^3 + K(2n^2+24n) against a measured FMA
             peak, or algorithmic HBM bytes against MEASURED_PEAKS.json; both given
  cpu_baseline  the unmodified reference (oracle/_ref) or the oracle port on the
             host cores, one process per core, on a bounded sample of the same batch

  shapes     (N = 1) the other BASELINE.json configurations, device resident, each with QP/s,
             mean iteration count and its own roofline fractions: n=4 (2^20, four input sets
             in rotation so that the 126 MB L2 cannot serve them), n=40 condensed linear MPC
             (2^18) and random (2^18), n=64 warm-started and cold (2^16)
  sustained  the headline step back to back for >= 2 s: QP/s, median SM clock, power, clock
             event reasons (does the burst number survive a sustained FP32 load?)
  e2e.copy_ceiling   the same bytes moved by the copy engines alone (no kernel), all ranks at
             once; e2e.frac_of_ceiling = e2e.value / that
  e2e_device_mask  (N > 1) ONE qp_solve_batch() call on rank 0 over all N GPUs with
             device_mask (the library's own shard + gather), the other ranks idle

`--impl reference` times only that CPU arm (rank 0; other ranks exit 0).
"""
from __future__ import annotations

import argparse
import json
import math
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched QP solves/sec at n=16,m=32"  # the headline shape; --n N renames it
UNIT = "QP/s"


# ------------------------------------------------------------------------------------------
# problem generation (device side, torch as plumbing)
# ------------------------------------------------------------------------------------------
def make_batch(count, n, seed, device):
    """SURVEY.md 8(d) benchmark distribution: P = G'G/n + I, q ~ N(0,1), box +-beta."""
    import torch
    gen = torch.Generator(device=device).manual_seed(seed)
    P = torch.empty((count, n, n), device=device)
    q = torch.randn((count, n), generator=gen, device=device)
    eye = torch.eye(n, device=device)
    step = 1 << 18
    for lo in range(0, count, step):
        hi = min(count, lo + step)
        G = torch.randn((hi - lo, n, n), generator=gen, device=device)
        Pc = torch.baddbmm(eye.expand(hi - lo, n, n), G.transpose(1, 2), G, alpha=1.0 / n)
        P[lo:hi] = 0.5 * (Pc + Pc.transpose(1, 2))
    return P.contiguous(), q.contiguous()


# ------------------------------------------------------------------------------------------
# clocks during the timed region (NVML, same fields as the nvidia-smi recipe)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks",
        0x100: "display_clocks",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def sample(self):
        if not self.ok:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            try:
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        def loop():
            while not self._stop.is_set():
                self.sample()
                time.sleep(0.005)
        self._t = threading.Thread(target=loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()

    def report(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        rep = {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.power:
            rep["power_w_median"] = float(np.median(self.power))
            rep["power_w_max"] = float(np.max(self.power))
        return rep


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# ------------------------------------------------------------------------------------------
_CPU_DATA = None  # (kind, n, P, q, lb, ub): inherited by the forked workers, never pickled


def _cpu_worker(span):
    a, b = span
    kind, n, P, q, lb, ub, fast = _CPU_DATA
    import oracle
    if kind == "reference":
        x, _ = oracle.Reference(n, fast=fast).admm_many(P[a:b], q[a:b], lb[a:b], ub[a:b], 10000)
    else:
        _, x = oracle.Oracle().admm_timed(P[a:b], q[a:b], lb[a:b], ub[a:b], 10000)
    return float(np.abs(x).sum())


class CpuArm:
    """One forked process per core (the reference's static pools are not thread
    safe, kmalloc/kmalloc.c:37-42), each solving a contiguous slice."""

    def __init__(self, n, P, q, beta, cores=None, sample_note="the first problems of the timed batch"):
        global _CPU_DATA
        self.sample_note = sample_note
        import oracle
        oracle.build(ref=False)
        self.n = n
        self.kind = "reference" if oracle.ref_available(n) else "port"
        # the timed arm uses the build made for speed (oracle/Makefile `reffast`), parity tests
        # the one without contraction
        self.fast = False
        self.trial = None
        if self.kind == "reference" and oracle.ref_available(n, fast=True):
            # two builds of the unmodified sources (oracle/Makefile): time both on one core, on
            # a few problems, and give the CPU its faster one
            k = min(256, P.shape[0])
            Pt, qt = np.ascontiguousarray(P[:k], np.float64), np.ascontiguousarray(q[:k], np.float64)
            tt = []
            for fast in (False, True):
                r = oracle.Reference(n, fast=fast)
                r.admm_many(Pt, qt, -beta, beta, 10000)
                tt.append(min(r.admm_many(Pt, qt, -beta, beta, 10000)[1] for _ in range(3)))
            self.fast = tt[1] < tt[0]
            self.trial = {oracle.REF_FLAGS: k / tt[0], oracle.REF_FAST_FLAGS: k / tt[1]}
        self.flags = (oracle.REF_FAST_FLAGS if self.fast else
                      (oracle.REF_FLAGS if self.kind == "reference" else "gcc -O3 (oracle port)"))
        self.cores = cores or len(os.sched_getaffinity(0))
        P = np.ascontiguousarray(P, np.float64)
        q = np.ascontiguousarray(q, np.float64)
        self.count = P.shape[0]
        lb = np.full((self.count, n), -beta)
        ub = np.full((self.count, n), beta)
        _CPU_DATA = (self.kind, n, P, q, lb, ub, self.fast)
        self.pool = mp.get_context("fork").Pool(self.cores)

    def step(self):
        edges = np.linspace(0, self.count, self.cores + 1).astype(int)
        jobs = [(int(a), int(b)) for a, b in zip(edges, edges[1:]) if b > a]
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, jobs, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self, value):
        return {"value": value, "unit": UNIT, "cores": self.cores, "kind": self.kind,
                "compiler": self.flags, "one_core_qp_per_s_by_build": self.trial,
                "sample": f"{self.count} problems of the benchmark distribution (n={self.n}; {self.sample_note}), "
                          f"one forked process per core, {self.kind} admm() in float64"}



# ------------------------------------------------------------------------------------------
# device-resident timing of one configuration, with its own roofline fractions
# ------------------------------------------------------------------------------------------
def executed_flops(kernel, n, k, warm):
    """Flops the kernels really execute per problem: thread-per-QP factors
    P + rho I as L D L' (n^3/3) and solves twice per iteration; the lane-group kernels run a
    block Gauss-Jordan (2 n^3 multiplied out, n^3 in product form, which is what a warm
    start or n >= 48 selects) and a mat-vec per iteration."""
    it = k * (2.0 * n * n + 24.0 * n)
    if kernel.startswith("tpq"):
        return n ** 3 / 3.0 + it
    if kernel.startswith("cpq"):
        return 2.0 * n ** 3 + it
    product_form = n >= 48 or (warm and n >= 24)
    return (1.0 if product_form else 2.0) * n ** 3 + it


def fractions(qp, n, count, ms, k, warm, hbm_peak, fp32_peak, issue_rec=None, sm_mhz=None, sms=148):
    kernel = qp.kernel_for(n)
    rate = count / (ms * 1e-3)
    bytes_per = qp.algorithmic_bytes(n, warm=warm, with_z=True, bound_arrays=False)
    ref = qp.algorithmic_flops(n, k)
    exe = executed_flops(kernel, n, k, warm)
    d = {"kernel": kernel, "qp_per_s": rate, "ms_per_step": ms, "mean_iters": k,
         "hbm_frac": bytes_per * rate / 1e9 / hbm_peak, "bytes_per_problem": bytes_per,
         "fp32_reference_frac": ref * rate / 1e12 / fp32_peak, "reference_flops_per_problem": ref,
         "fp32_executed_frac": exe * rate / 1e12 / fp32_peak, "executed_flops_per_problem": exe}
    # issue slots: warp instructions per launch from the committed ncu capture of this kernel at
    # this batch size (profiles/traffic.json), over this run's launch time
    if issue_rec and issue_rec.get("warp_instructions_per_launch") and sm_mhz:
        d["issue_frac"] = issue_rec["warp_instructions_per_launch"] / (ms * 1e-3) / (sms * 4 * sm_mhz * 1e6)
        d["issue_source"] = issue_rec.get("source", "profiles/traffic.json")
    else:
        d["issue_frac"] = None
    d["roofline_frac"] = max(d["hbm_frac"], d["fp32_reference_frac"])
    d["bound"] = "hbm" if d["hbm_frac"] >= d["fp32_reference_frac"] else "fp32"
    return d


def time_steps(torch, fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_shapes(args, qp, torch, device, hbm_peak, fp32_peak, traffic, sm_mhz):
    """BASELINE.json configs 2, 4, 5 (and n = 40 random, n = 64 cold beside them)."""
    from embedded_qp_solver_b200 import problems
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    steps = args.shape_steps
    lines = []

    def entry(name, n, count, ms, out, warm, note):
        k = float(out["iters"].float().mean())
        d = {"shape": name, "n": n, "m": 2 * n, "count": count, "warm_start": warm}
        d.update(fractions(qp, n, count, ms, k, warm, hbm_peak, fp32_peak,
                           traffic.get(note.get("issue_key") or f"{qp.kernel_for(n)}@{count}")
                           if (note.get("issue_ok") or note.get("issue_key")) else None,
                           sm_mhz, sms))
        d["not_converged"] = int((out["status"] != 0).sum())
        d["inputs"] = note["inputs"]
        lines.append(d)

    # config 2: n = 4, 2^20 problems.  One set is 84 MB, less than the L2: rotate four
    n, count = 4, 1 << 20
    sets = [make_batch(count, n, args.seed + 17 * i, device) for i in range(4)]
    out = qp.solve_batch(sets[0][0], sets[0][1], -args.beta, args.beta)
    ms = time_steps(torch, lambda i: qp.solve_batch(sets[i % 4][0], sets[i % 4][1], -args.beta, args.beta, out=out),
                    steps * 4, warmup=4)
    entry("n=4,m=8 random", n, count, ms, out, False,
          {"inputs": "4 input sets of 84 MB in rotation (336 MB > 126 MB L2)", "issue_ok": True})
    del sets, out

    # config 4: n = 40, 2^18 condensed linear-MPC problems (4096 distinct, tiled: 1.7 GB, all
    # addresses distinct) and, beside it, random problems of the same shape
    n, count = 40, 1 << 18
    Pm, qm, _, ubm = problems.mpc(4096, seed=args.seed)
    reps = count // 4096
    P = torch.from_numpy(Pm).to(device).repeat(reps, 1, 1).contiguous()
    q = torch.from_numpy(qm).to(device).repeat(reps, 1).contiguous()
    bm = float(ubm[0, 0])
    for sched, tag in ((qp.SCHEDULE_SINGLE, "single launch"),
                       (qp.SCHEDULE_PASSES, "first window + streaming pass"),
                       (qp.SCHEDULE_AUTO, "default options: AUTO picks the streaming pass from the "
                                          "lane efficiency its first solve measures")):
        o = qp.Options(schedule=sched)
        out = qp.solve_batch(P, q, -bm, bm, options=o)
        ms = time_steps(torch, lambda i: qp.solve_batch(P, q, -bm, bm, out=out, options=o), steps, warmup=4)
        entry(f"n=40,m=80 condensed MPC ({tag})", n, count, ms, out, False,
              {"inputs": "4096 distinct problems tiled 64 times, 1.7 GB > L2"})
        it = out["iters"].float()
        lines[-1]["iters_max"] = int(it.max())
        lines[-1]["iters_median"] = float(it.median())
        lines[-1]["accuracy_note"] = ("x within 2e-4 of the FP64 reference on all problems, 1e-4 on 99.5 % "
                                      "(the reference algorithm itself in binary32 reaches 1.1e-4; "
                                      "profiles/r1_mpc_accuracy.jsonl)")
    del P, q, out
    P, q = make_batch(count, n, args.seed, device)
    out = qp.solve_batch(P, q, -args.beta, args.beta)
    # (default options: AUTO still believes in the MPC set's tail for the first solves)
    for _ in range(3):
        qp.solve_batch(P, q, -args.beta, args.beta, out=out)
        torch.cuda.synchronize()
    st0 = qp.schedule_stats()
    ms = time_steps(torch, lambda i: qp.solve_batch(P, q, -args.beta, args.beta, out=out), steps)
    st1 = qp.schedule_stats()
    entry("n=40,m=80 random", n, count, ms, out, False, {"inputs": "1.7 GB > L2", "issue_ok": True})
    # beside BASELINE.json's shapes: the same q with one matrix per 1024 problems (qp_batch.p_share)
    Ps = P[::1024].contiguous()
    outs = qp.solve_batch(Ps, q, -args.beta, args.beta, p_share=1024)
    mss = time_steps(torch, lambda i: qp.solve_batch(Ps, q, -args.beta, args.beta, out=outs, p_share=1024), steps)
    shared_line = {"shape": "n=40,m=80 random, one P per 1024 problems (qp_batch.p_share; not a BASELINE shape)",
                   "n": n, "count": count, "qp_per_s": count / (mss * 1e-3), "ms_per_step": mss,
                   "mean_iters": float(outs["iters"].float().mean()),
                   "not_converged": int((outs["status"] != 0).sum()),
                   "note": "each matrix is eliminated once (qp_bpq.cuh PASS_SETUP), every problem is then "
                           "solved by the streaming kernel with the inverse of its group"}
    del Ps, outs
    lines[-1]["schedule"] = {"options": "default (AUTO)",
                             "solves_in_passes": (st1["seen"] - st0["seen"]) // count,
                             "solves_single": (st1["probe_problems"] - st0["probe_problems"]) // count}
    del P, q, out

    # config 5: n = 64, 2^16 problems warm-started from the solution of a perturbed problem
    n, count = 64, 1 << 16
    P, q = make_batch(count, n, args.seed, device)
    first = qp.solve_batch(P, q, -args.beta, args.beta)
    out = {k: v.clone() for k, v in first.items()}
    ms = time_steps(torch, lambda i: qp.solve_batch(P, q, -args.beta, args.beta, out=out), steps)
    entry("n=64,m=128 cold", n, count, ms, out, False,
          {"inputs": "1.07 GB > L2", "issue_key": f"{qp.kernel_for(n)}@{count} cold"})
    q2 = q + 0.01 * torch.randn_like(q)
    z0, u0 = first["z"].clone(), first["lam"].clone()
    ms = time_steps(torch, lambda i: qp.solve_batch(P, q2, -args.beta, args.beta, z0=z0, u0=u0, out=out), steps)
    entry("n=64,m=128 warm start (q perturbed by 1 %)", n, count, ms, out, True,
          {"inputs": "1.07 GB > L2", "issue_ok": True})
    del P, q, q2, out, first, z0, u0
    torch.cuda.empty_cache()
    # beside BASELINE.json's shapes: rows l <= A x <= u with a general A (SURVEY.md 8(f) rank 3;
    # no reference behaviour beyond A = I -- parity unpinned), n = 16, m = 32, both kernels
    n, m, count = 16, 32, 1 << 17
    P, q = make_batch(count, n, args.seed, device)
    gen = torch.Generator(device=device).manual_seed(args.seed + 7)
    A = (torch.randn((count, m, n), generator=gen, device=device) / n ** 0.5).contiguous()
    ub = (0.5 * (0.2 + 0.8 * torch.rand((count, m), generator=gen, device=device))).contiguous()
    lb = (-ub).contiguous()
    general = {"shape": "n=16,m=32 general rows l <= A x <= u, random A (not a BASELINE shape; parity unpinned beyond A = I)",
               "n": n, "m": m, "count": count, "inputs": "403 MB > L2"}
    for name, kern in (("registers (qp_genr.cuh, default)", qp.KERNEL_GEN_REG),
                       ("shared memory (qp_gen.cuh)", qp.KERNEL_GEN_SMEM)):
        o = qp.Options(kernel=kern)
        outg = qp.solve_general_batch(P, q, A, lb, ub, options=o)
        msg = time_steps(torch, lambda i: qp.solve_general_batch(P, q, A, lb, ub, options=o), steps)
        k = float(outg["iters"].float().mean())
        # flops of the iteration as stated (x = R t, A x, A'z, A'u: 2 (n^2 + 3 m n) per iteration) and of
        # the setup (A'A: 2 m n^2, inverse: 2 n^3)
        fl = k * 2.0 * (n * n + 3 * m * n) + 2.0 * m * n * n + 2.0 * n ** 3
        general[name] = {"qp_per_s": count / (msg * 1e-3), "ms_per_step": msg, "mean_iters": k,
                         "not_converged": int((outg["status"] != 0).sum()),
                         "fp32_frac_algorithm_flops": fl * count / (msg * 1e-3) / (fp32_peak * 1e12) if fp32_peak else None}
    del P, q, A, lb, ub, outg
    torch.cuda.empty_cache()
    run_shapes.beside = [shared_line, general]
    return lines

# ------------------------------------------------------------------------------------------
def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--count", type=int, default=1 << 20, help="problems per GPU")
    ap.add_argument("--beta", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--cpu-sample", type=int, default=400_000)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="qp_options.chunk of the host path (0 = library's choice)")
    ap.add_argument("--no-shapes", action="store_true", help="skip the other BASELINE shapes (N = 1)")
    ap.add_argument("--shape-steps", type=int, default=10)
    ap.add_argument("--sustained-s", type=float, default=2.5, help="0 = skip the sustained run")
    ap.add_argument("--no-shared", action="store_true", help="skip the shared-P leg beside the line")
    ap.add_argument("--no-device-mask", action="store_true", help="skip e2e_device_mask (N > 1)")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def config_of(args, world):
    return {"workload": f"{args.count} random strictly-convex box QPs per GPU, n={args.n}, "
                        f"m={2 * args.n} (P=G'G/n+I, q~N(0,1), box +-{args.beta}), FP32",
            "n": args.n, "m": 2 * args.n, "batch_per_gpu": args.count,
            "global_batch": args.count * world, "parallelism": f"shard{world}",
            "l2_policy": "inputs (1.2 GB/step at n=16) larger than the 126 MB L2"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    sample = min(args.count, args.cpu_sample)
    P, q = make_batch(sample, args.n, args.seed, torch.device("cpu"))
    arm = CpuArm(args.n, P.numpy(), q.numpy(), args.beta,
                 sample_note=f"make_batch on the host, seed {args.seed}, all distinct")
    for _ in range(max(1, min(args.warmup, 1))):
        arm.step()
    secs = [arm.step() for _ in range(args.steps)]
    arm.close()
    total = sum(secs)
    value = sample * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args, world), "cpu_baseline": arm.describe(value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import embedded_qp_solver_b200 as qp

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        # host-side barrier for the leg in which rank 0 drives every GPU: an NCCL barrier would
        # park a spinning kernel on the GPUs the library is about to use
        cpu_group = dist.new_group(backend="gloo")
    qp.load_library()
    n, count = args.n, args.count

    P, q = make_batch(count, n, args.seed + 1000 * rank, device)
    out = {k: torch.empty((count, n), device=device) for k in ("x", "z", "lam")}
    out.update({k: torch.empty(count, dtype=torch.int32, device=device)
                for k in ("status", "iters")})
    opts = qp.Options()

    def step():
        qp.solve_batch(P, q, -args.beta, args.beta, options=opts, out=out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sampler.sample()  # the GPU is still executing the queued steps here
    torch.cuda.synchronize()
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    barrier()
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = count * world / (ms_per_step * 1e-3)

    iters = out["iters"].float()
    mean_iters = float(iters.mean())
    n_bad = int((out["status"] != 0).sum())

    # ---- sustained: the same step back to back for seconds, clocks and power sampled ----------
    sustained = None
    if args.sustained_s > 0:
        ssamp = ClockSampler(local_rank)
        barrier()
        sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ssamp.start()
        sv0.record()
        ssteps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < args.sustained_s:  # the launch queue is bounded: the
            for _ in range(64):                            # host cannot run far ahead
                step()
            ssteps += 64
        sv1.record()
        torch.cuda.synchronize()
        ssamp.stop()
        sms_ = sv0.elapsed_time(sv1)
        if world > 1:
            t = torch.tensor([sms_], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sms_ = float(t.item())
        sustained = {"seconds": sms_ * 1e-3, "steps": ssteps, "ms_per_step": sms_ / ssteps,
                     "value": count * world * ssteps / (sms_ * 1e-3), "unit": UNIT,
                     "vs_burst": (sms_ / ssteps) / ms_per_step, "clocks": ssamp.report()}

    # ---- end to end through the C ABI with host buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        hP, hq = P.cpu().pin_memory(), q.cpu().pin_memory()
        hout = {k: torch.empty((count, n)).pin_memory() for k in ("x", "z", "lam")}
        hout.update({k: torch.empty(count, dtype=torch.int32).pin_memory()
                     for k in ("status", "iters")})
        hopts = qp.Options(device=local_rank, chunk=args.e2e_chunk)
        for _ in range(max(3, min(args.warmup, 5))):
            qp.solve_batch(hP, hq, -args.beta, args.beta, options=hopts, out=hout)
        ksteps = args.e2e_steps or min(args.steps, 20)
        barrier()
        wire0 = qp.host_path_h2d_bytes()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            qp.solve_batch(hP, hq, -args.beta, args.beta, options=hopts, out=hout)
        torch.cuda.synchronize()
        wire = (qp.host_path_h2d_bytes() - wire0) // ksteps
        dt = (time.perf_counter() - t0) / ksteps
        if world > 1:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        same = bool(torch.equal(hout["iters"], out["iters"].cpu()))
        # what the copy engines and the host's memory system allow for the same bytes with no
        # kernel in between: the very buffers of the timed run, H2D on one stream and D2H on
        # another, five rounds back to back behind a barrier so that all ranks copy AT THE SAME
        # TIME (a probe that allocates its own buffers lets the ranks drift apart and overstates
        # what N GPUs get from one host)
        h2d_b, d2h_b = (n * n + n) * 4 * count, (3 * n + 2) * 4 * count
        sc0, sc1 = torch.cuda.Stream(device), torch.cuda.Stream(device)
        dPf, hPf = P.view(-1), hP.view(-1)

        def copy_rounds(p_floats, rounds=5):
            barrier()
            t0c = time.perf_counter()
            for _ in range(rounds):
                with torch.cuda.stream(sc0):
                    dPf[:p_floats].copy_(hPf[:p_floats], non_blocking=True)
                    q.copy_(hq, non_blocking=True)
                with torch.cuda.stream(sc1):
                    for k_ in hout:
                        hout[k_].copy_(out[k_], non_blocking=True)
            torch.cuda.synchronize()
            dtc = (time.perf_counter() - t0c) / rounds
            if world > 1:
                tt = torch.tensor([dtc], device=device)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dtc = float(tt.item())
            return dtc

        copy_rounds(count * n * n, rounds=1)
        ceil_s = copy_rounds(count * n * n)
        # ... and for the bytes the library really sent (it may pack the upper triangles)
        wire_p_floats = max(0, int(wire) - n * 4 * count) // 4
        ceil_w = ceil_s if wire_p_floats >= count * n * n else copy_rounds(wire_p_floats)
        P.copy_(hP)  # the rounds above left (the same) data in place; make that explicit
        torch.cuda.synchronize()
        # a caller that produces the QP_P_UPPER layout itself (row-major upper triangles, 544
        # instead of 1024 bytes of P per problem at n = 16): same call, 47 % fewer bytes
        upper = None
        if n in (8, 12, 16):
            iu = torch.triu_indices(n, n)
            hU = torch.zeros((count, qp.packed_stride(n))).pin_memory()
            hU[:, :iu.shape[1]] = hP[:, iu[0], iu[1]]
            for _ in range(2):
                qp.solve_batch(hU, hq, -args.beta, args.beta, options=hopts, out=hout, p_layout=qp.P_UPPER)
            barrier()
            tu0 = time.perf_counter()
            for _ in range(ksteps):
                qp.solve_batch(hU, hq, -args.beta, args.beta, options=hopts, out=hout, p_layout=qp.P_UPPER)
            torch.cuda.synchronize()
            udt = (time.perf_counter() - tu0) / ksteps
            if world > 1:
                t = torch.tensor([udt], device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                udt = float(t.item())
            upper = {"value": count * world / udt, "unit": UNIT, "ms_per_step": udt * 1e3,
                     "h2d_bytes_per_step": (qp.packed_stride(n) + n) * 4 * count,
                     "d2h_bytes_per_step": d2h_b,
                     "matches_device_path": bool(torch.equal(hout["iters"], out["iters"].cpu()))}
            del hU
        # the reference's admm() hands back x and nothing else (qp_solvers.c:318): the same call
        # with z and lambda left out of the solution struct (NULL pointers are skipped)
        xonly = None
        xout = {k: hout[k] for k in ("x", "status", "iters")}
        for _ in range(2):
            qp.solve_batch(hP, hq, -args.beta, args.beta, options=hopts, out=xout, want=("x", "status", "iters"))
        barrier()
        tx0 = time.perf_counter()
        for _ in range(ksteps):
            qp.solve_batch(hP, hq, -args.beta, args.beta, options=hopts, out=xout, want=("x", "status", "iters"))
        torch.cuda.synchronize()
        xdt = (time.perf_counter() - tx0) / ksteps
        if world > 1:
            t = torch.tensor([xdt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            xdt = float(t.item())
        xonly = {"value": count * world / xdt, "unit": UNIT, "ms_per_step": xdt * 1e3,
                 "h2d_bytes_per_step": (n * n + n) * 4 * count, "d2h_bytes_per_step": (n + 2) * 4 * count,
                 "outputs": "x, status, iters (what the reference's solver signature returns, plus the status)",
                 "matches_device_path": bool(torch.equal(hout["iters"], out["iters"].cpu()))}
        # pageable caller memory (staged through the workspace's pinned buffers), fewer steps
        pageable = None
        if world == 1:
            pP, pq = hP.numpy().copy(), hq.numpy().copy()
            pout = {k: np.empty((count, n), np.float32) for k in ("x", "z", "lam")}
            pout.update({k: np.empty(count, np.int32) for k in ("status", "iters")})
            qp.solve_batch(pP, pq, -args.beta, args.beta, options=hopts, out=pout)
            tp0 = time.perf_counter()
            for _ in range(3):
                qp.solve_batch(pP, pq, -args.beta, args.beta, options=hopts, out=pout)
            pdt = (time.perf_counter() - tp0) / 3
            pc = qp.probe_copy_seconds(h2d_b, d2h_b, device=local_rank, pinned=False, repeats=2)
            pageable = {"value": count / pdt, "unit": UNIT, "ms_per_step": pdt * 1e3,
                        "driver_pageable_copy": count / pc,
                        "note": "driver_pageable_copy = the same bytes through cudaMemcpyAsync on pageable "
                                "memory with no kernel; the library stages through its own pinned buffers "
                                "with worker threads instead, which is why it is faster than that",
                        "matches_device_path": bool(np.array_equal(pout["iters"], out["iters"].cpu().numpy()))}
            del pP, pq, pout
        # the caller hands over full row-major matrices; the library may send only their
        # upper triangles over PCIe (qp_host_path_packs)
        packs = qp.host_path_packs(n)
        e2e = {"value": count * world / dt, "unit": UNIT,
               "h2d_bytes_per_step": (n * n + n) * 4 * count,
               "d2h_bytes_per_step": (3 * n + 2) * 4 * count,
               "h2d_bytes_on_the_wire_per_step": wire,
               "host_packs_upper_triangle": bool(packs),
               "steps": ksteps, "ms_per_step": dt * 1e3, "matches_device_path": same,
               "host_memory": "pinned",
               "copy_ceiling": count * world / ceil_w, "frac_of_ceiling": ceil_w / dt,
               "copy_ceiling_dense": count * world / ceil_s,
               "copy_ceiling_note": "the H2D bytes the library put on the wire and the D2H bytes of one step, "
                                    "from and to the same pinned buffers, two streams, no kernel, five rounds "
                                    "with all ranks copying at the same time (max over ranks); "
                                    "copy_ceiling_dense = the same for the caller's dense bytes (what a plain "
                                    "copy of the row-major matrices is bound by)",
               "pageable": pageable, "caller_packed_upper_layout": upper, "x_status_iters_only": xonly}
        del hP, hq, hout

    # ---- beside the metric: one model, many right-hand sides (qp_batch.p_share) -----------------
    # The same 2^20 problems' q with ONE matrix per 1024 problems: what an MPC host does with the
    # reference (the same qf->p, a new qf->q per state).  Not BASELINE.json's workload -- every
    # problem there has its own P -- so it is reported beside the line, never as `value` or `e2e`.
    shared = None
    if world == 1 and not args.no_e2e and not args.no_shared:
        share = 1024
        Ps = P[::share].contiguous()
        so = qp.solve_batch(Ps, q, -args.beta, args.beta, p_share=share)
        sms_ = time_steps(torch, lambda i: qp.solve_batch(Ps, q, -args.beta, args.beta, out=so, p_share=share),
                          min(args.steps, 50))
        hPs, hq2 = Ps.cpu().pin_memory(), q.cpu().pin_memory()
        hso = {k: torch.empty((count, n)).pin_memory() for k in ("x", "z", "lam")}
        hso.update({k: torch.empty(count, dtype=torch.int32).pin_memory() for k in ("status", "iters")})
        res = {}
        for tag, want in (("all_outputs", ("x", "z", "lam", "status", "iters")), ("x_status_iters", ("x", "status", "iters"))):
            o_ = {k: hso[k] for k in want}
            for _ in range(3):
                qp.solve_batch(hPs, hq2, -args.beta, args.beta, out=o_, want=want, p_share=share,
                               options=qp.Options(device=local_rank))
            ts0 = time.perf_counter()
            for _ in range(10):
                qp.solve_batch(hPs, hq2, -args.beta, args.beta, out=o_, want=want, p_share=share,
                               options=qp.Options(device=local_rank))
            torch.cuda.synchronize()
            sdt = (time.perf_counter() - ts0) / 10
            res[tag] = {"value": count / sdt, "unit": UNIT, "ms_per_step": sdt * 1e3,
                        "h2d_bytes_per_step": (n * n * (count // share) + n * count) * 4,
                        "d2h_bytes_per_step": sum((n if k in ("x", "z", "lam") else 1) for k in want) * 4 * count}
        shared = {"workload": f"{count} problems, n={n}: the q of the timed batch, one P per {share} problems",
                  "p_share": share, "device_resident": {"value": count / (sms_ * 1e-3), "unit": UNIT, "ms_per_step": sms_},
                  "e2e": res, "mean_iters": float(so["iters"].float().mean()),
                  "not_converged": int((so["status"] != 0).sum())}
        del Ps, so, hPs, hq2, hso

    # ---- the library's own multi-GPU entry: ONE qp_solve_batch call over all N GPUs ----------
    e2e_mask = None
    if world > 1 and not args.no_e2e and not args.no_device_mask:
        barrier()
        dist.barrier(group=cpu_group)
        if rank == 0:
            total = count * world
            mP = torch.empty((total, n, n)).pin_memory()
            mq = torch.empty((total, n)).pin_memory()
            Pc, qc = P.cpu(), q.cpu()
            for r in range(world):  # every shard carries rank 0's problems: the check below is exact
                mP[r * count:(r + 1) * count] = Pc
                mq[r * count:(r + 1) * count] = qc
            mout = {k: torch.empty((total, n)).pin_memory() for k in ("x", "z", "lam")}
            mout.update({k: torch.empty(total, dtype=torch.int32).pin_memory() for k in ("status", "iters")})
            mopts = qp.Options(device_mask=(1 << world) - 1)
            for _ in range(2):
                qp.solve_batch(mP, mq, -args.beta, args.beta, options=mopts, out=mout)
            msteps = max(3, min(args.steps, 5))
            t0 = time.perf_counter()
            for _ in range(msteps):
                qp.solve_batch(mP, mq, -args.beta, args.beta, options=mopts, out=mout)
            mdt = (time.perf_counter() - t0) / msteps
            ref_it = out["iters"].cpu()
            same = all(bool(torch.equal(mout["iters"][r * count:(r + 1) * count], ref_it)) for r in range(world))
            e2e_mask = {"value": total / mdt, "unit": UNIT, "ms_per_step": mdt * 1e3, "steps": msteps,
                        "problems": total, "device_mask": (1 << world) - 1,
                        "h2d_bytes_per_step": (n * n + n) * 4 * total,
                        "d2h_bytes_per_step": (3 * n + 2) * 4 * total,
                        "host_memory": "pinned", "matches_device_path": same,
                        "note": "one process, one call, one host thread per GPU inside the library; "
                                "the other ranks wait at a barrier"}
            del mP, mq, mout
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline --------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        hbm_peak = float(json.load(open(peaks_path))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json (burst copy)"
    fp32_peak = qp.probe_fp32_tflops(local_rank, 5)
    kernel_s = ms_per_step * 1e-3  # one launch per step: the step IS the kernel
    # this launch: scalar box (no lb/ub arrays), cold start, x/z/lambda/status/iters written
    bytes_per = qp.algorithmic_bytes(n, warm=False, with_z=True, bound_arrays=False)
    flops_per = qp.algorithmic_flops(n, mean_iters)
    hbm = {"bound": "hbm", "achieved": bytes_per * count / kernel_s / 1e9, "peak": hbm_peak,
           "unit": "GB/s", "peak_source": hbm_src, "bytes_per_problem": bytes_per}
    hbm["frac"] = hbm["achieved"] / hbm["peak"]
    fp = {"bound": "fp32", "achieved": flops_per * count / kernel_s / 1e12, "peak": fp32_peak,
          "unit": "TFLOP/s", "peak_source": "qp_probe_fp32_tflops (FFMA loop, this run)",
          "flops_per_problem": flops_per, "mean_iters": mean_iters,
          "flops_model": "reference algorithm: 8/3 n^3 + K (2 n^2 + 24 n)"}
    fp["frac"] = fp["achieved"] / fp["peak"] if fp32_peak > 0 else None
    # what the kernels execute: LDL' of P + rho I (n^3/3) and two triangular solves per
    # iteration instead of the reference's LU + explicit inverse (8/3 n^3) and a mat-vec
    exe = n ** 3 / 3.0 + mean_iters * (2.0 * n * n + 24.0 * n)
    fp["executed_flops_per_problem"] = exe
    fp["executed_frac"] = exe * count / kernel_s / 1e12 / fp32_peak if fp32_peak > 0 else None
    fp["note"] = ("frac is a reference-work rate (SURVEY 8d flops model); the kernel needs fewer flops "
                  "than the reference algorithm, so it can exceed 1; executed_frac counts the flops "
                  "the kernel really does, roofline.issue its issue slots")
    binding = fp if (fp["frac"] or 0) >= hbm["frac"] else hbm
    roofline = dict(binding)
    # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture
    # (profiles/traffic.json, keyed by kernel and batch size); null when not captured
    roofline["traffic"] = None
    rec = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(f"{qp.kernel_for(n)}@{count}")
        if rec:
            roofline["traffic"] = rec["dram_bytes_per_launch"]
            roofline["traffic_source"] = rec["source"]
    # instruction-issue view (the binding resource of this kernel): warp
    # instructions per launch from the same ncu capture over this run's launch time, against
    # one issue per cycle per SM sub-partition at the sampled SM clock
    rec = rec if os.path.exists(tpath) else None
    clk = (sampler.report().get("sm_mhz") or 0) * 1e6
    if rec and rec.get("warp_instructions_per_launch") and clk > 0:
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        peak_issue = sms * 4 * clk
        ach = rec["warp_instructions_per_launch"] / kernel_s
        roofline["issue"] = {"achieved": ach / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-inst/s",
                             "frac": ach / peak_issue,
                             "warp_instructions_per_launch": rec["warp_instructions_per_launch"]}
    roofline["kernel"] = qp.kernel_for(n)
    roofline["hbm"] = hbm
    roofline["fp32"] = fp

    # ---- the other BASELINE shapes (N = 1 only: they are single-GPU configurations) ----------
    shapes = None
    if world == 1 and not args.no_shapes:
        traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
        del out
        shapes = run_shapes(args, qp, torch, device, hbm_peak, fp32_peak, traffic,
                            sampler.report().get("sm_mhz"))
        headline = {"shape": f"n={n},m={2 * n} random (headline)", "n": n, "m": 2 * n, "count": count,
                    "warm_start": False}
        headline.update(fractions(qp, n, count, ms_per_step, mean_iters, False, hbm_peak, fp32_peak,
                                  rec, sampler.report().get("sm_mhz"),
                                  torch.cuda.get_device_properties(device).multi_processor_count))
        headline["not_converged"] = n_bad
        headline["inputs"] = "1.2 GB > L2"
        shapes.insert(1, headline)

    # ---- CPU baseline ----------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        sample = min(count, args.cpu_sample)
        arm = CpuArm(n, P[:sample].cpu().numpy(), q[:sample].cpu().numpy(), args.beta)
        arm.step()
        secs = arm.step()
        arm.close()
        cpu = arm.describe(sample / secs)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args, world),
            "clocks": sampler.report(), "e2e": e2e, "gpu_launches": args.steps * world,
            "roofline": roofline, "cpu_baseline": cpu,
            "mean_iters": mean_iters, "not_converged": n_bad,
            "sustained": sustained, "shapes": shapes}
    if e2e_mask is not None:
        line["e2e_device_mask"] = e2e_mask
    if shared is not None:
        line["shared_p"] = shared
    if getattr(run_shapes, "beside", None):
        line["shapes_beside_baseline"] = run_shapes.beside
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global METRIC
    args = parse()
    METRIC = f"batched QP solves/sec at n={args.n},m={2 * args.n}"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
