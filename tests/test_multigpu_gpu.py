"""GPU: the library's own host-memory path -- sharding over several GPUs (qp_options.device_mask),
workspace lifetime, packed input layout.

What it replaces: the per-problem loop over the solver function pointer at
/root/reference/test/test.c:97 (one admm() call per problem on one core).  SURVEY.md 8(e): the
batch is cut into contiguous shards, one host thread and one copy/solve pipeline per GPU,
results land in the caller's arrays at the shard's offset; no collective.
The multi-device tests skip on a box with one GPU (`gpurun --gpus 2` runs them).
"""
import threading

import numpy as np
import pytest
import torch

from embedded_qp_solver_b200 import problems

pytestmark = pytest.mark.gpu

KEYS = ("x", "z", "lam", "status", "iters")


def same(a, b, what):
    for k in KEYS:
        assert np.array_equal(np.asarray(a[k]).view(np.int32), np.asarray(b[k]).view(np.int32)), \
            f"{what}: {k} differs"


def ngpu(qp):
    return qp.load_library().qp_device_count()


def pinned_like(a):
    t = torch.from_numpy(a).pin_memory()
    return t


@pytest.mark.parametrize("count", [1, 3, 1000, 40_003])
@pytest.mark.parametrize("pinned", [False, True])
def test_device_mask_matches_single_gpu(qp, count, pinned):
    """One host batch over GPUs {0,1} and over all GPUs: bit-identical to GPU 0 alone.  Counts
    below the number of devices, not divisible by it, and large enough for several chunks."""
    nd = ngpu(qp)
    if nd < 2:
        pytest.skip("needs at least two GPUs")
    n = 16
    P, q, _, _ = problems.wellcond(count, n, seed=77, beta=0.5)
    lb = np.full((count, n), -0.5, np.float32)
    ub = np.full((count, n), 0.4, np.float32)
    args = [pinned_like(a) if pinned else a for a in (P, q, lb, ub)]
    ref = qp.solve_batch(P, q, lb, ub, options=qp.Options(device=0))
    for mask in (0x3, (1 << nd) - 1):
        out = None
        if pinned:
            out = {k: torch.empty((count, n)).pin_memory() for k in ("x", "z", "lam")}
            out.update({k: torch.empty(count, dtype=torch.int32).pin_memory() for k in ("status", "iters")})
        got = qp.solve_batch(*args, options=qp.Options(device_mask=mask), out=out)
        got = {k: (v.numpy() if hasattr(v, "numpy") and not isinstance(v, np.ndarray) else v) for k, v in got.items()}
        same(got, ref, f"mask {mask:#x} count {count} pinned {pinned}")


def test_device_mask_other_shapes_and_options(qp):
    nd = ngpu(qp)
    if nd < 2:
        pytest.skip("needs at least two GPUs")
    mask = (1 << nd) - 1
    for n, count in ((4, 5001), (40, 2001), (64, 301), (96, 33)):
        P, q, _, _ = problems.wellcond(count, n, seed=78, beta=0.5)
        rng = np.random.default_rng(n)
        z0 = (0.1 * rng.standard_normal((count, n))).astype(np.float32)
        u0 = (0.1 * rng.standard_normal((count, n))).astype(np.float32)
        rho = (0.5 + rng.random(count)).astype(np.float32)
        ref = qp.solve_batch(P, q, -0.5, 0.5, z0=z0, u0=u0, rho=rho, options=qp.Options(device=0))
        got = qp.solve_batch(P, q, -0.5, 0.5, z0=z0, u0=u0, rho=rho,
                             options=qp.Options(device_mask=mask, chunk=256))
        same(got, ref, f"n={n}")


def test_explicit_workspace_over_all_gpus(qp):
    nd = ngpu(qp)
    if nd < 2:
        pytest.skip("needs at least two GPUs")
    n = 16
    ws = qp.Workspace(n, 100_000, device_mask=(1 << nd) - 1)
    try:
        for count in (100_000, 7, 31_111):
            P, q, _, _ = problems.wellcond(count, n, seed=count, beta=0.5)
            ref = qp.solve_batch(P, q, -0.5, 0.5, options=qp.Options(device=0))
            got = qp.solve_batch(P, q, -0.5, 0.5, workspace=ws)
            same(got, ref, f"workspace count {count}")
    finally:
        ws.close()


def test_two_host_threads_different_shapes(qp):
    """Two host threads call qp_solve_batch() at once with different n, alternating batch
    sizes: each keeps its own default workspace (round 1 freed one under the other)."""
    errors = []
    sets = {}
    for n in (16, 40):
        for count in (3000, 45_000 if n == 16 else 6000):
            P, q, _, _ = problems.wellcond(count, n, seed=5 * n + count % 7, beta=0.5)
            sets[(n, count)] = (P, q, qp.solve_batch(P, q, -0.5, 0.5))

    def worker(n):
        try:
            for rep in range(6):
                for (nn, count), (P, q, ref) in sets.items():
                    if nn != n:
                        continue
                    got = qp.solve_batch(P, q, -0.5, 0.5)
                    same(got, ref, f"thread n={n} count={count} rep={rep}")
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=worker, args=(n,)) for n in (16, 40)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_same_shape_from_two_threads(qp):
    n, count = 16, 20_000
    P, q, _, _ = problems.wellcond(count, n, seed=123, beta=0.5)
    ref = qp.solve_batch(P, q, -0.5, 0.5)
    errors = []

    def worker():
        try:
            for _ in range(5):
                same(qp.solve_batch(P, q, -0.5, 0.5), ref, "same shape")
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=worker) for _ in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("n", [8, 12, 16])
def test_packed_upper_layout(qp, n):
    """QP_P_UPPER: the caller hands over the upper triangles only; device and host path give
    the bits of the dense layout through the same kernel family."""
    count = 5000 + n
    P, q, _, _ = problems.wellcond(count, n, seed=9, beta=0.5)
    Pu = qp.pack_upper(P)
    assert Pu.shape == (count, qp.packed_stride(n))
    ref = qp.solve_batch(P, q, -0.5, 0.5, options=qp.Options(kernel=qp.KERNEL_TPQ))
    host = qp.solve_batch(Pu, q, -0.5, 0.5, p_layout=qp.P_UPPER, options=qp.Options(chunk=1024))
    same(host, ref, "host packed")
    d = qp.solve_batch(torch.from_numpy(Pu).cuda(), torch.from_numpy(q).cuda(), -0.5, 0.5,
                       p_layout=qp.P_UPPER)
    torch.cuda.synchronize()
    same({k: v.cpu().numpy() for k, v in d.items()}, ref, "device packed")
    with pytest.raises(qp.QpError):   # only the thread-per-QP kernels read this layout
        qp.solve_batch(Pu, q, -0.5, 0.5, p_layout=qp.P_UPPER, options=qp.Options(kernel=qp.KERNEL_GPQ))


def test_packed_layout_rejected_for_other_dimensions(qp):
    P = np.zeros((4, qp.packed_stride(40)), np.float32)
    q = np.zeros((4, 40), np.float32)
    with pytest.raises(qp.QpError) as e:
        qp.solve_batch(P, q, p_layout=qp.P_UPPER)
    assert e.value.code == -2


def test_upper_only_matrix_needs_the_packed_layout(qp):
    """which triangle a kernel reads is unspecified for QP_P_DENSE.  A caller who
    holds only the upper triangle passes it packed, and then every bit agrees with the full
    symmetric matrix."""
    n, count = 16, 999
    P, q, _, _ = problems.wellcond(count, n, seed=4, beta=0.5)
    ref = qp.solve_batch(P, q, -0.5, 0.5)
    upper_only = np.triu(P)                       # lower triangle zeroed
    got = qp.solve_batch(qp.pack_upper(upper_only), q, -0.5, 0.5, p_layout=qp.P_UPPER)
    same(got, ref, "upper only")


def test_error_in_the_middle_leaves_the_workspace_usable(qp):
    """A launch that cannot run (forced kernel family that does not take the shape) fails the
    call; the next call on the same default workspace works and returns the right answer."""
    n, count = 16, 10_000
    P, q, _, _ = problems.wellcond(count, n, seed=8, beta=0.5)
    ref = qp.solve_batch(P, q, -0.5, 0.5)
    with pytest.raises(qp.QpError):
        qp.solve_batch(P[:, :12, :12].copy(), q[:, :12].copy(), -0.5, 0.5,
                       options=qp.Options(kernel=qp.KERNEL_BPQ))   # bpq takes n % 8 == 0 only
    same(qp.solve_batch(P, q, -0.5, 0.5), ref, "after the failures")


@pytest.mark.parametrize("n", [16, 40])
def test_shared_matrices_over_several_gpus(qp, n):
    """qp_batch.p_share with device_mask: the shards start in the middle of a group of problems
    that share a matrix (count and shard sizes are not multiples of the group size)."""
    nd = ngpu(qp)
    if nd < 2:
        pytest.skip("needs at least two GPUs")
    count, share = 30_011, 1000
    nmat = (count + share - 1) // share
    Pm = problems.wellcond(nmat, n, seed=5, beta=0.5)[0]
    q = problems.wellcond(count, n, seed=6, beta=0.5)[1]
    P = np.repeat(Pm, share, axis=0)[:count]
    ref = qp.solve_batch(P, q, -0.5, 0.5, options=qp.Options(device=0))
    for mask in (0x3, (1 << nd) - 1):
        got = qp.solve_batch(Pm, q, -0.5, 0.5, options=qp.Options(device_mask=mask, chunk=4096), p_share=share)
        same(got, ref, f"shared, n {n}, mask {mask:#x}")
