"""GPU: qp_solve_general_batch (rows l <= A x <= u; csrc/qp_genr.cuh for n <= 32, m <= 64 --
operands in registers -- and csrc/qp_gen.cuh -- operands in shared memory -- for the rest).

* A = I: the kernel must reproduce the reference's admm() -- same golden vectors, same bars
  as the box kernels (iteration counts identical, x within 1e-4).
* general A: no reference behaviour exists (parity unpinned); the kernel is compared with the
  float64 numpy restatement oracle/general_np.py, with the KKT conditions at a tight
  tolerance, and with an independent scipy solve."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden, nondegenerate
from embedded_qp_solver_b200 import problems
from oracle import general_np

pytestmark = pytest.mark.gpu
X_TOL = 1e-4


def rel_err(x, x_ref):
    return np.max(np.abs(x - x_ref), axis=1) / (1 + np.max(np.abs(x_ref), axis=1))


def random_rows(count, n, m, seed, scale=0.5):
    rng = np.random.default_rng(seed)
    A = (rng.normal(size=(count, m, n)) / np.sqrt(n)).astype(np.float32)
    l = (-scale * rng.uniform(0.2, 1.0, (count, m))).astype(np.float32)
    u = (scale * rng.uniform(0.2, 1.0, (count, m))).astype(np.float32)
    return A, l, u


@pytest.mark.parametrize("kernel", ["auto", "smem"])
@pytest.mark.parametrize("name", ["wellcond_b05_n16.npz", "wellcond_b05_n40.npz",
                                  "wellcond_free_n8.npz", "mpc_n40.npz", "wellcond_b05_n64.npz"])
def test_identity_rows_reproduce_the_reference_golden_vectors(qp, name, kernel):
    g = load_golden(os.path.join(os.path.dirname(golden_files()[0]), name))
    count, n = g["q"].shape
    if kernel == "smem" and n > 32:
        pytest.skip("auto is the shared-memory kernel at this size")
    eye = np.ascontiguousarray(np.broadcast_to(np.eye(n, dtype=np.float32), (count, n, n)))
    r = qp.solve_general_batch(g["P"], g["q"], eye, g["lb"], g["ub"], options=qp.Options(
        kernel=qp.KERNEL_GEN_SMEM if kernel == "smem" else qp.KERNEL_AUTO))
    ok = nondegenerate(g["margins"]) & (g["status"] == 0)
    assert ok.sum() >= 0.5 * count
    assert np.array_equal(r["iters"][ok], g["iters"][ok])
    assert np.array_equal(r["status"][ok], g["status"][ok])
    assert rel_err(r["x"][ok].astype(np.float64), g["x_ref"][ok]).max() <= X_TOL
    assert rel_err(r["z"][ok].astype(np.float64), g["z"][ok]).max() <= X_TOL
    # and it is the same answer the tuned box kernels give
    b = qp.solve_batch(g["P"], g["q"], g["lb"], g["ub"])
    assert np.array_equal(b["iters"][ok], r["iters"][ok])
    assert rel_err(b["x"][ok].astype(np.float64), r["x"][ok].astype(np.float64)).max() <= X_TOL


@pytest.mark.parametrize("n,m", [(4, 8), (16, 32), (5, 3), (40, 80), (64, 128), (33, 70),
                                 (8, 16), (9, 16), (24, 48), (20, 64), (32, 65), (1, 1)])
def test_general_rows_against_the_numpy_restatement(qp, n, m):
    # counts that leave the last warp's groups ragged (4, 8, 16 or 32 lanes per problem)
    count = 43 if n <= 16 else 11
    P, q, _, _ = problems.wellcond(count, n, seed=2000 + n)
    A, l, u = random_rows(count, n, m, seed=n)
    o = general_np.admm_general_batch(P, q, A, l, u)
    tP, tq, tA, tl, tu = (torch.from_numpy(v).cuda() for v in (P, q, A, l, u))
    r = qp.solve_general_batch(tP, tq, tA, tl, tu)
    torch.cuda.synchronize()
    r = {k: v.cpu().numpy() for k, v in r.items()}
    ok = nondegenerate(o["margins"]) & (o["status"] == 0)
    assert ok.sum() >= 0.7 * count
    assert np.array_equal(r["iters"][ok], o["iters"][ok]) and np.all(r["status"][ok] == 0)
    assert rel_err(r["x"][ok].astype(np.float64), o["x"][ok]).max() <= X_TOL
    assert rel_err(r["z"][ok].astype(np.float64), o["z"][ok]).max() <= X_TOL
    assert rel_err(r["y"][ok].astype(np.float64), o["u"][ok]).max() <= 10 * X_TOL
    # host-pointer path: same bits
    h = qp.solve_general_batch(P, q, A, l, u)
    assert np.array_equal(h["x"], r["x"]) and np.array_equal(h["iters"], r["iters"])


def test_general_rows_kkt_and_scipy_at_tight_tolerance(qp):
    from scipy.optimize import LinearConstraint, minimize
    count, n, m = 6, 8, 14
    P, q, _, _ = problems.wellcond(count, n, seed=77)
    A, l, u = random_rows(count, n, m, seed=78, scale=0.3)
    l[:, :3] = -np.inf  # one-sided rows
    r = qp.solve_general_batch(P, q, A, l, u,
                               options=qp.Options(abstol=1e-7, reltol=1e-7, max_iter=20000))
    assert np.all(r["status"] == 0)
    for c in range(count):
        P64, q64, A64 = P[c].astype(np.float64), q[c].astype(np.float64), A[c].astype(np.float64)
        x, y = r["x"][c].astype(np.float64), r["y"][c].astype(np.float64)
        assert np.max(np.abs(P64 @ x + q64 + A64.T @ y)) <= 2e-5 * (1 + np.max(np.abs(q64)))
        ax = A64 @ x
        assert np.all(ax >= l[c] - 1e-4) and np.all(ax <= u[c] + 1e-4)
        sp = minimize(lambda v: 0.5 * v @ P64 @ v + q64 @ v, np.zeros(n),
                      jac=lambda v: P64 @ v + q64, hess=lambda v: P64, method="trust-constr",
                      constraints=[LinearConstraint(A64, l[c], u[c])],
                      options=dict(gtol=1e-10, xtol=1e-12, maxiter=2000))
        assert np.max(np.abs(x - sp.x)) <= 1e-3 * (1 + np.max(np.abs(sp.x)))


@pytest.mark.parametrize("n,m", [(3, 5), (8, 16), (16, 32), (16, 17), (13, 27), (17, 20), (24, 48),
                                 (32, 64), (31, 64), (20, 33), (1, 1)])
def test_register_kernel_against_the_shared_memory_kernel_and_the_restatement(qp, n, m):
    """The two kernels run the same iteration with different summation orders in the mat-vecs
    (the elimination has the same bits): same iteration counts wherever the stop test is not a
    coin toss, same x; ragged counts leave partly filled warps; an unaligned A takes the scalar
    staging path of the register kernel."""
    count = 203 if n <= 16 else 37
    P, q, _, _ = problems.wellcond(count, n, seed=3000 + 7 * n + m)
    A, l, u = random_rows(count, n, m, seed=100 * n + m)
    l[::5, 0] = -np.inf
    u[1::7, -1] = np.inf
    o = general_np.admm_general_batch(P, q, A, l, u)
    ok = nondegenerate(o["margins"]) & (o["status"] == 0)
    assert ok.sum() >= 0.7 * count
    tP, tq, tl, tu = (torch.from_numpy(v).cuda() for v in (P, q, l, u))
    holder = torch.empty(A.size + 1, dtype=torch.float32, device="cuda")
    tA_unaligned = holder[1:].view(count, m, n)
    tA_unaligned.copy_(torch.from_numpy(A))
    tA = torch.from_numpy(A).cuda()
    res = {}
    for name, kern, a in (("reg", qp.KERNEL_GEN_REG, tA), ("reg_unaligned", qp.KERNEL_GEN_REG, tA_unaligned),
                          ("smem", qp.KERNEL_GEN_SMEM, tA), ("auto", qp.KERNEL_AUTO, tA)):
        r = qp.solve_general_batch(tP, tq, a, tl, tu, options=qp.Options(kernel=kern))
        torch.cuda.synchronize()
        res[name] = {k: v.cpu().numpy() for k, v in r.items()}
    for name, r in res.items():
        assert np.array_equal(r["iters"][ok], o["iters"][ok]), name
        assert np.all(r["status"][ok] == 0), name
        assert rel_err(r["x"][ok].astype(np.float64), o["x"][ok]).max() <= X_TOL, name
        assert rel_err(r["z"][ok].astype(np.float64), o["z"][ok]).max() <= X_TOL, name
        assert rel_err(r["y"][ok].astype(np.float64), o["u"][ok]).max() <= 10 * X_TOL, name
    reg, smem = res["reg"], res["smem"]
    for k in ("x", "z", "y", "iters", "status"):  # the same kernel, the same bits
        assert np.array_equal(reg[k], res["auto"][k]), k
        assert np.array_equal(reg[k], res["reg_unaligned"][k]), k
    assert np.array_equal(reg["status"], smem["status"])
    assert np.abs(reg["iters"] - smem["iters"]).max() <= 1
    same = reg["iters"] == smem["iters"]
    assert same.mean() >= 0.97
    assert rel_err(reg["x"][same].astype(np.float64), smem["x"][same].astype(np.float64)).max() <= 2e-5


def test_general_rows_kernel_choice(qp):
    count = 5
    for n, m in ((40, 80), (33, 10), (8, 65)):
        P, q, _, _ = problems.wellcond(count, n, seed=1)
        A, l, u = random_rows(count, n, m, seed=2)
        assert np.all(qp.solve_general_batch(P, q, A, l, u)["status"] == 0)  # auto: shared memory
        with pytest.raises(qp.QpError):
            qp.solve_general_batch(P, q, A, l, u, options=qp.Options(kernel=qp.KERNEL_GEN_REG))
    P, q, lb, ub = problems.wellcond(count, 16, seed=1)
    A, l, u = random_rows(count, 16, 32, seed=2)
    with pytest.raises(qp.QpError):  # a box family cannot take rows
        qp.solve_general_batch(P, q, A, l, u, options=qp.Options(kernel=qp.KERNEL_TPQ))
    with pytest.raises(qp.QpError):  # and the box entry point has no general-rows kernel
        qp.solve_batch(P, q, lb, ub, options=qp.Options(kernel=qp.KERNEL_GEN_REG))


def test_register_kernel_full_batch_properties(qp):
    """2^16 problems at n = 16, m = 32 (what tools/bench_general.py times): every problem converges,
    rows within their bounds to the primal tolerance, and a shuffled copy of the batch returns the
    shuffled results bit for bit (no dependence on the position inside a warp or on the CTA)."""
    count, n, m = 1 << 16, 16, 32
    gen = torch.Generator(device="cuda").manual_seed(11)
    G = torch.randn((count, n, n), generator=gen, device="cuda")
    P = (G.transpose(1, 2) @ G / n + torch.eye(n, device="cuda")).contiguous()
    q = torch.randn((count, n), generator=gen, device="cuda")
    A = (torch.randn((count, m, n), generator=gen, device="cuda") / n ** 0.5).contiguous()
    u = (0.5 * (0.2 + 0.8 * torch.rand((count, m), generator=gen, device="cuda"))).contiguous()
    l = (-u).contiguous()
    r = qp.solve_general_batch(P, q, A, l, u)
    assert int((r["status"] != 0).sum()) == 0
    ax = torch.einsum("bmn,bn->bm", A.double(), r["x"].double())
    viol = torch.clamp(torch.maximum(l.double() - ax, ax - u.double()), min=0).max(dim=1).values
    assert float(viol.max()) <= 2e-2 * 2.0  # reltol x the scale of A x, generous: a stop-test bound
    perm = torch.randperm(count, generator=gen, device="cuda")
    rp = qp.solve_general_batch(P[perm].contiguous(), q[perm].contiguous(), A[perm].contiguous(),
                                l[perm].contiguous(), u[perm].contiguous())
    for k in ("x", "z", "y", "iters", "status"):
        assert torch.equal(rp[k], r[k][perm]), k
    # host pointers: chunks alternate between two device buffers and two streams (four chunks
    # here, the last one ragged); pageable and pinned arrays, the device path's bits
    part = 30_001
    host = [t[:part].cpu() for t in (P, q, A, l, u)]
    for pinned in (False, True):
        args = [t.pin_memory() if pinned else t.numpy() for t in host]
        out = None
        if pinned:  # results written in place by the copy engine; x and iters preallocated only
            out = {"x": torch.empty((part, n)).pin_memory(),
                   "iters": torch.empty(part, dtype=torch.int32).pin_memory()}
        for _ in range(2):
            h = qp.solve_general_batch(*args, out=out)
            assert out is None or h is out
            for k in ("x", "z", "y", "iters", "status"):
                assert np.array_equal(np.asarray(h[k]), r[k][:part].cpu().numpy()), (k, pinned)
    with pytest.raises(ValueError):
        qp.solve_general_batch(*[t.numpy() for t in host], out={"x": np.empty((3, n), np.float32)})


@pytest.mark.parametrize("kernel", ["auto", "smem"])
def test_general_rows_warm_start_relaxation_and_status(qp, kernel):
    count, n, m = 30, 16, 32
    P, q, _, _ = problems.wellcond(count, n, seed=91)
    A, l, u = random_rows(count, n, m, seed=92)
    kopt = qp.KERNEL_GEN_SMEM if kernel == "smem" else qp.KERNEL_AUTO
    first = qp.solve_general_batch(P, q, A, l, u, options=qp.Options(kernel=kopt))
    q2 = problems.perturb_q(q, seed=5)
    z0, u0 = first["z"], first["y"]  # rho = 1
    o = general_np.admm_general_batch(P, q2, A, l, u, z0=z0.astype(np.float64),
                                      u0=u0.astype(np.float64), alpha=1.6)
    r = qp.solve_general_batch(P, q2, A, l, u, z0=z0, u0=u0, options=qp.Options(alpha=1.6, kernel=kopt))
    ok = nondegenerate(o["margins"]) & (o["status"] == 0)
    assert np.array_equal(r["iters"][ok], o["iters"][ok])
    assert rel_err(r["x"][ok].astype(np.float64), o["x"][ok]).max() <= X_TOL
    cold = general_np.admm_general_batch(P, q2, A, l, u, alpha=1.6)
    assert o["iters"].mean() < cold["iters"].mean()
    # P + rho A'A not positive definite -> status 2, neighbours unaffected
    Pb = P.copy()
    Pb[4] = -40 * np.eye(n, dtype=np.float32)
    rb = qp.solve_general_batch(Pb, q, A, l, u, options=qp.Options(kernel=kopt))
    assert rb["status"][4] == qp.STATUS_SINGULAR and rb["iters"][4] == 0
    assert np.array_equal(np.delete(rb["iters"], 4), np.delete(first["iters"], 4))
    with pytest.raises(qp.QpError):
        qp.solve_general_batch(np.zeros((1, 4, 4), np.float32), np.zeros((1, 4), np.float32),
                               np.zeros((1, 129, 4), np.float32), np.zeros((1, 129), np.float32),
                               np.zeros((1, 129), np.float32))
