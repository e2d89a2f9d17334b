#!/usr/bin/env python
"""Randomised differential check: QP_SCHEDULE_PASSES (list passes / first window + streaming
pass) and shared matrices against QP_SCHEDULE_SINGLE with replicated matrices, bit for bit, over
random dimensions, counts, distributions, options and window policies.

    python tools/fuzz_schedules.py [--cases 200] [--seed 1]
Prints one line per failure and a summary; exit code 1 on any difference."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embedded_qp_solver_b200 as qp  # noqa: E402
from embedded_qp_solver_b200 import problems  # noqa: E402

KEYS = ("x", "z", "lam", "status", "iters")


def cuda(v):
    return torch.from_numpy(np.ascontiguousarray(v)).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    mpcP, mpcq, _, mpcub = problems.mpc(512, seed=5)
    bad = 0
    for case in range(args.cases):
        n = int(rng.choice([4, 8, 12, 16, 24, 32, 40]))
        count = int(rng.choice([1, 2, 3, 31, 33, 100, 777, 4097, 12000]))
        dist = rng.choice(["wellcond", "refdist", "mpc"] if n == 40 else ["wellcond", "refdist"])
        share = int(rng.choice([1, 1, 2, 5, 64]))
        nmat = (count + share - 1) // share
        seed = int(rng.integers(1 << 30))
        if dist == "mpc":
            idx = rng.integers(0, 512, nmat)
            Pm = mpcP[idx]
            q = mpcq[rng.integers(0, 512, count)]
            box = float(mpcub[0, 0])
        elif dist == "refdist":
            Pm = problems.refdist(nmat, n, seed=seed)[0]
            q = problems.refdist(count, n, seed=seed + 1)[1]
            box = 1e4
        else:
            Pm = problems.wellcond(nmat, n, seed=seed)[0]
            q = problems.wellcond(count, n, seed=seed + 1)[1]
            box = 0.5
        P = np.repeat(Pm, share, axis=0)[:count]
        kw = {}
        lo, hi = -box, box
        if rng.random() < 0.4:
            lo = cuda((-box * (0.5 + rng.random((count, n)))).astype(np.float32))
            hi = cuda((box * (0.5 + rng.random((count, n)))).astype(np.float32))
        if rng.random() < 0.3:
            kw["z0"] = cuda((0.1 * box * rng.standard_normal((count, n))).astype(np.float32))
            kw["u0"] = cuda((0.1 * box * rng.standard_normal((count, n))).astype(np.float32))
        if rng.random() < 0.3:
            kw["rho"] = cuda((0.5 + 2 * rng.random(count)).astype(np.float32))
        opt = dict(max_iter=int(rng.choice([7, 30, 200, 1500])), alpha=float(rng.choice([1.0, 1.0, 1.6])))
        w0 = int(rng.choice([2, 4, 6, 24]))
        os.environ["QPB200_PASSES"] = f"{w0},{int(rng.integers(0, 2))}"
        if rng.random() < 0.25:
            os.environ["QPB200_BPQ_STREAM"] = "0"
        else:
            os.environ.pop("QPB200_BPQ_STREAM", None)
        ref = qp.solve_batch(cuda(P), cuda(q), lo, hi, options=qp.Options(schedule=qp.SCHEDULE_SINGLE, **opt), **kw)
        got = qp.solve_batch(cuda(Pm), cuda(q), lo, hi, options=qp.Options(schedule=qp.SCHEDULE_PASSES, **opt),
                             p_share=share, **kw)
        torch.cuda.synchronize()
        diff = [k for k in KEYS if not torch.equal(ref[k].view(torch.int32), got[k].view(torch.int32))]
        if diff:
            bad += 1
            print(f"DIFF case {case}: n={n} count={count} dist={dist} share={share} w0={w0} opt={opt} "
                  f"kw={sorted(kw)} arrays={torch.is_tensor(lo)} stream={os.environ.get('QPB200_BPQ_STREAM')} -> {diff}",
                  flush=True)
    print(f"{args.cases} cases, {bad} with differences")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
