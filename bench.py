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
        self.samples, self.reasons = [], set()
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
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# ------------------------------------------------------------------------------------------
_CPU_DATA = None  # (kind, n, P, q, lb, ub): inherited by the forked workers, never pickled


def _cpu_worker(span):
    a, b = span
    kind, n, P, q, lb, ub = _CPU_DATA
    import oracle
    if kind == "reference":
        x, _ = oracle.Reference(n).admm_many(P[a:b], q[a:b], lb[a:b], ub[a:b], 10000)
    else:
        _, x = oracle.Oracle().admm_timed(P[a:b], q[a:b], lb[a:b], ub[a:b], 10000)
    return float(np.abs(x).sum())


class CpuArm:
    """One forked process per core (the reference's static pools are not thread
    safe, kmalloc/kmalloc.c:37-42), each solving a contiguous slice."""

    def __init__(self, n, P, q, beta, cores=None):
        global _CPU_DATA
        import oracle
        oracle.build(ref=False)
        self.n = n
        self.kind = "reference" if oracle.ref_available(n) else "port"
        self.cores = cores or len(os.sched_getaffinity(0))
        P = np.ascontiguousarray(P, np.float64)
        q = np.ascontiguousarray(q, np.float64)
        self.count = P.shape[0]
        lb = np.full((self.count, n), -beta)
        ub = np.full((self.count, n), beta)
        _CPU_DATA = (self.kind, n, P, q, lb, ub)
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
                "sample": f"{self.count} problems of the same seeded batch (n={self.n}), "
                          f"one forked process per core, {self.kind} admm() in float64"}


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
    from embedded_qp_solver_b200 import problems
    sample = min(args.count, args.cpu_sample)
    P, q, _, _ = problems.wellcond(min(sample, 4096), args.n, seed=args.seed, beta=args.beta)
    reps = math.ceil(sample / P.shape[0])  # tile a seeded block up to the sample size
    P = np.tile(P, (reps, 1, 1))[:sample]
    q = np.tile(q, (reps, 1))[:sample]
    arm = CpuArm(args.n, P, q, args.beta)
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
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
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

    # ---- end to end through the C ABI with host buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        hP, hq = P.cpu().pin_memory(), q.cpu().pin_memory()
        hout = {k: torch.empty((count, n)).pin_memory() for k in ("x", "z", "lam")}
        hout.update({k: torch.empty(count, dtype=torch.int32).pin_memory()
                     for k in ("status", "iters")})
        hopts = qp.Options(device=local_rank)
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
        # the caller hands over full row-major matrices; the library may send only their
        # upper triangles over PCIe (qp_host_path_packs)
        packs = qp.host_path_packs(n)
        e2e = {"value": count * world / dt, "unit": UNIT,
               "h2d_bytes_per_step": (n * n + n) * 4 * count,
               "d2h_bytes_per_step": (3 * n + 2) * 4 * count,
               "h2d_bytes_on_the_wire_per_step": wire,
               "host_packs_upper_triangle": bool(packs),
               "steps": ksteps, "ms_per_step": dt * 1e3, "matches_device_path": same,
               "host_memory": "pinned"}
        del hP, hq, hout

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
            "mean_iters": mean_iters, "not_converged": n_bad}
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
