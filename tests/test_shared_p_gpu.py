"""GPU: qp_batch.p_share -- problems that share their matrix P (one model, many right-hand
sides: an MPC host calls the reference's admm(), /root/reference/qp_solvers/qp_solvers.c:255,
with the same qf->p and a new qf->q per state; struct _quadratic_form keeps them apart,
qp_solvers/qp.h:8-13).

Sharing changes where P comes from and nothing else: every result must be bit-identical to the
same batch with P replicated per problem, in every kernel family, through the device path, the
chunked host path (chunks that start in the middle of a group), the one-copy path of small
batches, with packed upper triangles, warm starts, per-problem rho and the pass schedules."""
import numpy as np
import pytest
import torch

from embedded_qp_solver_b200 import problems

pytestmark = pytest.mark.gpu

KEYS = ("x", "z", "lam", "status", "iters")


def make(count, n, share, seed, dist="wellcond"):
    nmat = (count + share - 1) // share
    gen = problems.wellcond if dist == "wellcond" else problems.refdist
    Pm = gen(nmat, n, seed=seed)[0]
    q = gen(count, n, seed=seed + 1)[1]
    P = np.repeat(Pm, share, axis=0)[:count].copy()
    return Pm, P, q


def same(a, b, what):
    for k in KEYS:
        av = a[k].cpu().numpy() if torch.is_tensor(a[k]) else np.asarray(a[k])
        bv = b[k].cpu().numpy() if torch.is_tensor(b[k]) else np.asarray(b[k])
        assert np.array_equal(av.view(np.int32), bv.view(np.int32)), f"{what}: {k}"


def cuda(*a):
    return [torch.from_numpy(np.ascontiguousarray(v)).cuda() if v is not None else None for v in a]


@pytest.mark.parametrize("n,kernel", [(4, "KERNEL_TPQ"), (8, "KERNEL_TPQ"), (16, "KERNEL_TPQ"),
                                      (16, "KERNEL_BPQ"), (24, "KERNEL_BPQ"), (40, "KERNEL_BPQ"),
                                      (20, "KERNEL_GPQ"), (64, "KERNEL_GPQ"), (96, "KERNEL_CPQ")])
def test_device_path_every_family(qp, n, kernel):
    count = 1000 + 13 if n <= 40 else 133
    rng = np.random.default_rng(n)
    for share in (2, 7, 32, 100, count, count + 5):
        Pm, P, q = make(count, n, share, seed=100 + n + share)
        lb = (-0.5 - rng.random((count, n))).astype(np.float32)
        ub = (0.3 + rng.random((count, n))).astype(np.float32)
        z0 = (0.1 * rng.standard_normal((count, n))).astype(np.float32)
        rho = (0.5 + rng.random(count)).astype(np.float32)
        o = qp.Options(kernel=getattr(qp, kernel))
        tPm, tP, tq, tlb, tub, tz0, trho = cuda(Pm, P, q, lb, ub, z0, rho)
        for kw in (dict(), dict(lb=tlb, ub=tub), dict(lb=tlb, ub=tub, z0=tz0, u0=tz0, rho=trho)):
            lo, hi = kw.pop("lb", -0.5), kw.pop("ub", 0.5)
            ref = qp.solve_batch(tP, tq, lo, hi, options=o, **kw)
            got = qp.solve_batch(tPm, tq, lo, hi, options=o, p_share=share, **kw)
            torch.cuda.synchronize()
            same(got, ref, f"{kernel} n={n} share={share} {sorted(kw)}")


@pytest.mark.parametrize("n", [4, 16, 40, 64])
def test_host_path_chunks_that_split_groups(qp, n):
    count = 20_011 if n <= 16 else 3_001
    for share in (3, 1000, 4096):
        Pm, P, q = make(count, n, share, seed=7 * n + share)
        ref = qp.solve_batch(*cuda(P, q), -0.5, 0.5)
        torch.cuda.synchronize()
        for chunk in (0, 1024, 4096):                 # 1024 and 4096 are not multiples of 3 or 1000
            got = qp.solve_batch(Pm, q, -0.5, 0.5, options=qp.Options(chunk=chunk), p_share=share)
            same(got, ref, f"host n={n} share={share} chunk={chunk}")
        tPm, tq = torch.from_numpy(Pm).pin_memory(), torch.from_numpy(q).pin_memory()
        got = qp.solve_batch(tPm, tq, -0.5, 0.5, options=qp.Options(chunk=1024), p_share=share)
        same(got, ref, f"host pinned n={n} share={share}")


@pytest.mark.parametrize("n", [5, 16, 40])
def test_small_batches_one_copy_path(qp, n):
    for count, share in ((1, 4), (5, 2), (33, 32), (100, 7), (100, 1000)):
        Pm, P, q = make(count, n, share, seed=n + count + share)
        ref = qp.solve_batch(*cuda(P, q), -0.5, 0.5)
        torch.cuda.synchronize()
        got = qp.solve_batch(Pm, q, -0.5, 0.5, p_share=share)
        same(got, ref, f"small n={n} count={count} share={share}")


def test_packed_upper_triangles_shared(qp):
    n, count, share = 16, 5000, 64
    Pm, P, q = make(count, n, share, seed=3)
    ref = qp.solve_batch(*cuda(P, q), -0.5, 0.5)
    torch.cuda.synchronize()
    Um = qp.pack_upper(Pm)
    got = qp.solve_batch(*cuda(Um, q), -0.5, 0.5, p_layout=qp.P_UPPER, p_share=share)
    torch.cuda.synchronize()
    same(got, ref, "packed + shared, device")
    got = qp.solve_batch(Um, q, -0.5, 0.5, p_layout=qp.P_UPPER, p_share=share, options=qp.Options(chunk=1024))
    same(got, ref, "packed + shared, host")


@pytest.mark.parametrize("n", [16, 40])
def test_pass_schedules_with_shared_matrices(qp, n):
    """The reference's heavy-tailed distribution: list passes (n = 16), first window + streaming
    pass (n = 40)."""
    count, share = 2048 + 7, 16
    Pm, P, q = make(count, n, share, seed=40 + n, dist="refdist")
    o = dict(max_iter=1500)
    ref = qp.solve_batch(*cuda(P, q), -1e4, 1e4, options=qp.Options(schedule=qp.SCHEDULE_SINGLE, **o))
    for sch in (qp.SCHEDULE_SINGLE, qp.SCHEDULE_PASSES):
        got = qp.solve_batch(*cuda(Pm, q), -1e4, 1e4, options=qp.Options(schedule=sch, **o), p_share=share)
        torch.cuda.synchronize()
        same(got, ref, f"n={n} schedule={sch}")
    assert int(ref["iters"].max()) > 200


def test_reference_structs_with_a_shared_p(qp):
    """qp_admm_qf_batch() notices runs of structs that point at the same P."""
    n, count, share = 16, 300, 25
    Pm, P, q = make(count, n, share, seed=77)
    ref = qp.solve_batch(*cuda(P, q), -0.5, 0.5)
    torch.cuda.synchronize()
    mats = [qp.Matrix(Pm[i].astype(np.float64)) for i in range(len(Pm))]
    qfs = [qp.QuadraticForm(mats[i // share], q[i].astype(np.float64), 0.0) for i in range(count)]
    x, status, iters = qp.admm_many(qfs, 10000, -0.5, 0.5)
    assert np.array_equal(iters, ref["iters"].cpu().numpy())
    assert np.array_equal(np.asarray(x, np.float32), ref["x"].cpu().numpy())


def test_other_entry_points_refuse_shared_matrices(qp):
    Pm, P, q = make(10, 8, 5, seed=1)
    from embedded_qp_solver_b200 import api
    import ctypes as C
    lib = api.load_library()
    pb = api.qp_batch()
    pb.n, pb.count, pb.mem, pb.p_share = 8, 10, api.MEM_HOST, 5
    pb.P, pb.q = Pm.ctypes.data, q.ctypes.data
    x = np.zeros((10, 8), np.float32); f = np.zeros(10, np.float32)
    assert lib.qp_eval_batch(C.byref(pb), x.ctypes.data, C.c_float(0.0), f.ctypes.data,
                             None) == api.QP_ERR_UNSUPPORTED
