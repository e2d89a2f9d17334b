"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the container that has /root/reference:

    make -C oracle all && python tests/golden/make_golden.py

Inputs are drawn in float32 and widened to float64 for the reference, so the
FP32 CUDA path and the FP64 reference see the same numbers.  Per set we keep
  P,q,lb,ub (float32)          the problems
  x_ref (float64)              x returned by the reference's admm()
                               (/root/reference/qp_solvers/qp_solvers.c:255-319)
  z,u,iters,status,margins     from the restatement oracle/admm_oracle.c, kept
                               only after its x matched x_ref bit for bit
  f_ref (float64)              quadratic_form_eval(x_ref) (qp.c:9-27), the number
                               the reference's tests print (test/test.c:100-102)
  x_newton (float64)           reference Newton result, inactive-box sets only
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from embedded_qp_solver_b200 import problems  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

SETS = [
    # name, n, count, generator
    ("wellcond_b05", lambda n, c: problems.wellcond(c, n, seed=2026, beta=0.5)),
    ("wellcond_free", lambda n, c: problems.wellcond(c, n, seed=2027, beta=1e12)),
    ("refdist_1e4", lambda n, c: problems.refdist(c, n, seed=2028, box=1e4)),
]
COUNTS = {4: 64, 8: 32, 16: 64, 32: 16, 40: 16, 48: 8, 64: 8}


def main():
    orc = oracle.Oracle()
    for n, count in COUNTS.items():
        ref = oracle.Reference(n)
        for name, gen in SETS:
            P, q, lb, ub = gen(n, count)
            P64, q64, lb64, ub64 = (a.astype(np.float64) for a in (P, q, lb, ub))
            x_ref, _ = ref.admm_many(P64, q64, lb64, ub64, max_iter=oracle.REF_MAX_ITER)
            o = orc.admm(P64, q64, lb64, ub64)
            assert np.array_equal(x_ref, o["x"]), (n, name, "restatement != reference")
            f_ref = np.array([ref.eval(P64[c], q64[c], 0.0, x_ref[c]) for c in range(count)])
            extra = {}
            if name == "wellcond_free":
                extra["x_newton"] = np.stack(
                    [ref.newton(P64[c], q64[c], np.zeros(n), 10) for c in range(count)])
            path = os.path.join(OUT, f"{name}_n{n}.npz")
            np.savez_compressed(path, P=P, q=q, lb=lb, ub=ub, x_ref=x_ref, z=o["z"], u=o["u"],
                                iters=o["iters"], status=o["status"], margins=o["margins"],
                                f_ref=f_ref, **extra)
            print(f"{path}: {count} problems, iters {o['iters'].min()}..{o['iters'].max()}, "
                  f"status {np.bincount(o['status'], minlength=4).tolist()}")
    # the mpc shape (BASELINE config 4): n = 40
    ref = oracle.Reference(40)
    P, q, lb, ub = problems.mpc(16, seed=2029)
    P64, q64, lb64, ub64 = (a.astype(np.float64) for a in (P, q, lb, ub))
    x_ref, _ = ref.admm_many(P64, q64, lb64, ub64)
    o = orc.admm(P64, q64, lb64, ub64)
    assert np.array_equal(x_ref, o["x"])
    f_ref = np.array([ref.eval(P64[c], q64[c], 0.0, x_ref[c]) for c in range(16)])
    np.savez_compressed(os.path.join(OUT, "mpc_n40.npz"), P=P, q=q, lb=lb, ub=ub, x_ref=x_ref,
                        z=o["z"], u=o["u"], iters=o["iters"], status=o["status"],
                        margins=o["margins"], f_ref=f_ref)
    print("mpc_n40: iters", o["iters"].min(), o["iters"].max(),
          "status", np.bincount(o["status"], minlength=4).tolist())


if __name__ == "__main__":
    main()
