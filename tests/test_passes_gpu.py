"""GPU: the pass scheduler (QP_SCHEDULE_PASSES, include/qp_batch.h) against the plain
one-launch kernels.

The scheduler moves problems between warps and launches; it must never change a result.  A continued
problem repeats the factorisation and carries on from its saved (z, u, iteration count), so
x, z, lambda, status and iters are compared BIT FOR BIT with QP_SCHEDULE_SINGLE, which in turn
is pinned to the oracle and the reference's golden vectors in test_parity_gpu.py.  The
heavy-tailed inputs are the reference's own test distribution
(/root/reference/matrix/matrix_ops/matrix_ops.c:699-734, config.h:16-20).
"""
import os

import numpy as np
import pytest
import torch

from embedded_qp_solver_b200 import problems

pytestmark = pytest.mark.gpu

KEYS = ("x", "z", "lam", "status", "iters")


def dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() if a is not None else None
            for a in arrs]


def run(qp, P, q, lb, ub, schedule, **kw):
    opt = kw.pop("options", None) or qp.Options()
    opt.schedule = schedule
    args = dev(P, q)
    lb_ = dev(lb)[0] if isinstance(lb, np.ndarray) else lb
    ub_ = dev(ub)[0] if isinstance(ub, np.ndarray) else ub
    extra = {k: dev(v)[0] for k, v in kw.items() if isinstance(v, np.ndarray)}
    extra.update({k: v for k, v in kw.items() if not isinstance(v, np.ndarray)})
    r = qp.solve_batch(args[0], args[1], lb_, ub_, options=opt, **extra)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in r.items()}


def evictions(qp):
    s = qp.schedule_stats()
    return s["outlived"] + s["evicted_later"]


def same(a, b, what):
    for k in KEYS:
        assert np.array_equal(a[k].view(np.int32) if a[k].dtype == np.float32 else a[k],
                              b[k].view(np.int32) if b[k].dtype == np.float32 else b[k]), \
            f"{what}: {k} differs between the schedules"


@pytest.fixture
def policy():
    """QPB200_PASSES=w0,grow for the duration of a test."""
    old = os.environ.get("QPB200_PASSES")

    def set_(v):
        if v is None:
            os.environ.pop("QPB200_PASSES", None)
        else:
            os.environ["QPB200_PASSES"] = v
    yield set_
    if old is None:
        os.environ.pop("QPB200_PASSES", None)
    else:
        os.environ["QPB200_PASSES"] = old


@pytest.mark.parametrize("n", [4, 8, 12, 16])
def test_reference_distribution_thread_per_qp(qp, n):
    """3 ... 2000 iterations inside one warp: problems are evicted again and again, nothing changes."""
    count = 4096 + 37
    P, q, _, _ = problems.refdist(count, n, seed=11, box=1e4)
    opt = dict(max_iter=2000)
    before = evictions(qp)
    a = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_PASSES, options=qp.Options(**opt))
    after = evictions(qp)
    b = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_SINGLE, options=qp.Options(**opt))
    same(a, b, f"refdist n={n}")
    assert after - before > count // 4, f"no passes: {after - before} evictions"
    assert a["iters"].max() == 2000 and (a["status"] == 1).any()   # the cap is hit inside a pass
    assert a["iters"].min() < 16


@pytest.mark.parametrize("n", [24, 32, 40])
def test_reference_distribution_lane_groups(qp, n):
    count = 1024 + 5
    P, q, _, _ = problems.refdist(count, n, seed=12, box=1e4)
    before = evictions(qp)
    a = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_PASSES, options=qp.Options(max_iter=1500))
    after = evictions(qp)
    b = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_SINGLE, options=qp.Options(max_iter=1500))
    same(a, b, f"refdist n={n}")
    assert after - before > 0


def test_condensed_mpc_set(qp):
    """BASELINE config 4 (n = 40): 21 ... 2783 iterations."""
    P, q, lb, ub = problems.mpc(2048, seed=2026)
    beta = float(ub[0, 0])
    a = run(qp, P, q, -beta, beta, qp.SCHEDULE_PASSES)
    b = run(qp, P, q, -beta, beta, qp.SCHEDULE_SINGLE)
    same(a, b, "mpc n=40")


@pytest.mark.parametrize("n,kernel", [(16, "KERNEL_TPQ"), (4, "KERNEL_TPQ"), (40, "KERNEL_BPQ"),
                                      (16, "KERNEL_BPQ"), (8, "KERNEL_BPQ"), (24, "KERNEL_BPQ"),
                                      (32, "KERNEL_BPQ")])
def test_every_option_with_aggressive_eviction(qp, policy, n, kernel):
    """A window of 4 iterations that never grows: even well-conditioned problems change warps
    two or three times, the others a dozen times.  Bounds arrays, warm start,
    over-relaxation and per-problem rho (rho != 1: the saved u, not lambda, is carried)."""
    count = 777
    Pw, qw, _, _ = problems.wellcond(400, n, seed=5, beta=0.5)
    Pr, qr, _, _ = problems.refdist(count - 400, n, seed=6)   # these need hundreds of iterations
    P, q = np.concatenate([Pw, Pr]), np.concatenate([qw, qr])
    rng = np.random.default_rng(3)
    lb = (-0.5 - rng.random((count, n))).astype(np.float32)
    ub = (0.3 + rng.random((count, n))).astype(np.float32)
    z0 = (0.1 * rng.standard_normal((count, n))).astype(np.float32)
    u0 = (0.1 * rng.standard_normal((count, n))).astype(np.float32)
    rho = (0.5 + 2.0 * rng.random(count)).astype(np.float32)
    k = getattr(qp, kernel)
    cases = {
        "plain": dict(lb=-0.5, ub=0.5),
        "bounds": dict(lb=lb, ub=ub),
        "warm": dict(lb=-0.5, ub=0.5, z0=z0, u0=u0),
        "bounds+warm+rho": dict(lb=lb, ub=ub, z0=z0, u0=u0, rho=rho),
        "relax": dict(lb=lb, ub=ub, options=qp.Options(alpha=1.6, kernel=k)),
        "cap": dict(lb=-0.5, ub=0.5, options=qp.Options(max_iter=27, kernel=k)),
    }
    for name, kw in cases.items():
        lo, hi = kw.pop("lb"), kw.pop("ub")
        o = kw.pop("options", None) or qp.Options(kernel=k, max_iter=300)
        if name != "cap":
            o.max_iter = 300
        policy(None)
        ref = run(qp, P, q, lo, hi, qp.SCHEDULE_SINGLE,
                  options=qp.Options(**{**o.__dict__}), **kw)
        policy("4,0")
        before = evictions(qp)
        got = run(qp, P, q, lo, hi, qp.SCHEDULE_PASSES, options=qp.Options(**{**o.__dict__}), **kw)
        after = evictions(qp)
        same(got, ref, f"{kernel} n={n} {name}")
        if name == "plain":
            # (lane groups with the full inverse are evicted once: the streaming pass runs them out)
            least = count // 2 if kernel == "KERNEL_BPQ" else 2 * count
            assert after - before > least, f"aggressive policy did not evict: {after - before}"
        if kernel == "KERNEL_BPQ":
            # the list passes the streaming kernel replaced are still there (product form, and
            # QPB200_BPQ_STREAM=0): same bits
            os.environ["QPB200_BPQ_STREAM"] = "0"
            try:
                got = run(qp, P, q, lo, hi, qp.SCHEDULE_PASSES, options=qp.Options(**{**o.__dict__}), **kw)
            finally:
                os.environ.pop("QPB200_BPQ_STREAM", None)
            same(got, ref, f"{kernel} n={n} {name}, list passes")


def test_packed_upper_triangles_with_passes(qp, policy):
    count, n = 1500, 16
    P, q, _, _ = problems.refdist(count, n, seed=21, box=1e4)
    Pu = qp.pack_upper(P)
    opt = dict(max_iter=500)
    a = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_SINGLE, options=qp.Options(**opt))
    b = run(qp, Pu, q, -1e4, 1e4, qp.SCHEDULE_PASSES, options=qp.Options(**opt), p_layout=qp.P_UPPER)
    same(a, b, "packed + passes")


def test_host_path_with_passes(qp):
    """Chunks of the host-memory path each run their passes on the device's arena."""
    count, n = 9000, 16
    P, q, _, _ = problems.refdist(count, n, seed=31, box=1e4)
    o1 = qp.Options(max_iter=400, chunk=2048, schedule=qp.SCHEDULE_PASSES)
    o2 = qp.Options(max_iter=400, chunk=2048, schedule=qp.SCHEDULE_SINGLE)
    a = qp.solve_batch(P, q, -1e4, 1e4, options=o1)
    b = qp.solve_batch(P, q, -1e4, 1e4, options=o2)
    same(a, b, "host path")


def test_auto_follows_the_tail(qp):
    """QP_SCHEDULE_AUTO: small batches and small iteration caps are plain launches that count
    nothing; a large batch is a plain launch that adds up what its warps did and what they
    waited for, and the solve after one whose warps mostly waited runs in passes (and goes
    back when nothing outlives the first window any more)."""
    n = 16
    P, q, _, _ = problems.refdist(4096, n, seed=41, box=1e4)
    Pw, qw, _, _ = problems.wellcond(4096, n, seed=42, beta=0.5)
    ref = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_SINGLE, options=qp.Options(max_iter=300))
    # whatever earlier tests left behind: two well-conditioned solves put AUTO into single launches
    for _ in range(2):
        run(qp, Pw, qw, -0.5, 0.5, qp.SCHEDULE_AUTO, options=qp.Options(max_iter=300))
    s0 = qp.schedule_stats()
    run(qp, P[:512], q[:512], -1e4, 1e4, qp.SCHEDULE_AUTO, options=qp.Options(max_iter=300))
    run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_AUTO, options=qp.Options(max_iter=50))
    assert qp.schedule_stats() == s0     # small batch / few iterations allowed: nothing counted
    # heavy tail: the first solve probes, the second one runs in passes
    a = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_AUTO, options=qp.Options(max_iter=300))
    same(a, ref, "auto, probing")
    s1 = qp.schedule_stats()
    assert s1["probe_problems"] - s0["probe_problems"] == 4096 and s1["seen"] == s0["seen"]
    assert s1["probe_iterations"] - s0["probe_iterations"] == int(ref["iters"].sum())
    worked = (s1["probe_iterations"] - s0["probe_iterations"]) / (s1["probe_waited_for"] - s0["probe_waited_for"])
    assert worked < 0.5
    for k in range(2):
        a = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_AUTO, options=qp.Options(max_iter=300))
        same(a, ref, f"auto, passes #{k}")
    s2 = qp.schedule_stats()
    assert s2["seen"] - s1["seen"] == 2 * 4096 and s2["outlived"] - s1["outlived"] > 2 * 1024
    assert s2["evicted_later"] > s1["evicted_later"] and s2["probe_problems"] == s1["probe_problems"]
    # the tail disappears: one more solve in passes (nothing evicted), then single launches again
    for k in range(3):
        run(qp, Pw, qw, -0.5, 0.5, qp.SCHEDULE_AUTO, options=qp.Options(max_iter=300))
    s3 = qp.schedule_stats()
    assert s3["outlived"] == s2["outlived"] and s3["seen"] - s2["seen"] == 4096
    assert s3["probe_problems"] - s2["probe_problems"] == 2 * 4096


def test_two_streams_share_the_arena(qp):
    """Launches on different streams are chained through the arena's event."""
    n, count = 16, 4096
    P, q, _, _ = problems.refdist(count, n, seed=51, box=1e4)
    tP, tq = dev(P, q)
    ref = run(qp, P, q, -1e4, 1e4, qp.SCHEDULE_SINGLE, options=qp.Options(max_iter=300))
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = []
    for _ in range(4):
        for st in streams:
            with torch.cuda.stream(st):
                outs.append(qp.solve_batch(tP, tq, -1e4, 1e4,
                                           options=qp.Options(max_iter=300, schedule=qp.SCHEDULE_PASSES)))
    torch.cuda.synchronize()
    for r in outs:
        same({k: v.cpu().numpy() for k, v in r.items()}, ref, "streams")


def test_auto_streams_the_condensed_mpc_set(qp):
    """QP_SCHEDULE_AUTO on lane groups: the condensed-MPC set (n = 40, mean 36 iterations, 42 % of
    the lane-iterations of a single launch wait for the slowest of three problems) goes to the
    first window + streaming pass after one probing solve, well-conditioned problems bring it
    back to single launches; nothing changes a bit either way."""
    P, q, lb, ub = problems.mpc(4096, seed=2026)
    beta = float(ub[0, 0])
    Pw, qw, _, _ = problems.wellcond(4096, 40, seed=43, beta=0.5)
    ref = run(qp, P, q, -beta, beta, qp.SCHEDULE_SINGLE)
    refw = run(qp, Pw, qw, -0.5, 0.5, qp.SCHEDULE_SINGLE)
    for _ in range(3):   # whatever earlier tests left behind: back to single launches
        run(qp, Pw, qw, -0.5, 0.5, qp.SCHEDULE_AUTO)
    s0 = qp.schedule_stats()
    a = run(qp, P, q, -beta, beta, qp.SCHEDULE_AUTO)        # probes
    same(a, ref, "auto, probing")
    s1 = qp.schedule_stats()
    assert s1["probe_problems"] - s0["probe_problems"] == 4096 and s1["seen"] == s0["seen"]
    worked = (s1["probe_iterations"] - s0["probe_iterations"]) / (s1["probe_waited_for"] - s0["probe_waited_for"])
    assert 0.5 < worked < 0.75
    for k in range(2):
        a = run(qp, P, q, -beta, beta, qp.SCHEDULE_AUTO)    # first window + streaming pass
        same(a, ref, f"auto, streaming #{k}")
    s2 = qp.schedule_stats()
    assert s2["seen"] - s1["seen"] == 2 * 4096 and s2["probe_problems"] == s1["probe_problems"]
    assert s2["outlived"] - s1["outlived"] > 2 * 3000        # a window of 6: nearly everything streams
    b = run(qp, Pw, qw, -0.5, 0.5, qp.SCHEDULE_AUTO)        # still streaming; notices the short problems
    same(b, refw, "auto, well-conditioned while streaming")
    b = run(qp, Pw, qw, -0.5, 0.5, qp.SCHEDULE_AUTO)        # single launch again
    same(b, refw, "auto, back to single launches")
    s3 = qp.schedule_stats()
    assert s3["seen"] - s2["seen"] == 4096 and s3["probe_problems"] - s2["probe_problems"] == 4096
