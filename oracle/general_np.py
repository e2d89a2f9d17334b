"""oracle/general_np.py -- TEST INFRASTRUCTURE, not product code.

numpy (float64) restatement of the ADMM the reference runs for the box
(/root/reference/qp_solvers/qp_solvers.c:255-319), written for the splitting A x = z
so that rows  l <= A x <= u  can be handled.  Every update cites
the reference line it generalises; with A = I it is that line.  The reference itself has no
general-A path: PARITY UNPINNED for A != I.  For A = I this restatement is checked against
oracle.Oracle().admm (bit-identical to the unmodified reference) in tests/test_general_np.py.
"""
import numpy as np


def admm_general(P, q, A, l, u, rho=1.0, alpha=1.0, abstol=1e-4, reltol=1e-2, max_iter=10000,
                 z0=None, u0=None):
    """One problem.  Returns dict(x, z, u, iters, status, margins)."""
    P, q, A, l, u = (np.asarray(v, np.float64) for v in (P, q, A, l, u))
    m, n = A.shape
    z = np.zeros(m) if z0 is None else np.array(z0, np.float64)   # :263
    ud = np.zeros(m) if u0 is None else np.array(u0, np.float64)  # :264
    x = np.zeros(n)
    R = np.linalg.inv(P + rho * A.T @ A)                           # :285-292 (A = I: (P + rho I)^-1)
    status, iters, margins = 1, max_iter, (np.inf, np.inf)
    for i in range(max_iter):                                      # :284
        z_old = z.copy()                                           # :293-294
        x = R @ (rho * A.T @ (z - ud) - q)                         # admm_update_x :146-159
        ax = A @ x
        zh = alpha * ax + (1 - alpha) * z_old                      # admm_update_x_hat :161-174
        z = np.minimum(np.maximum(zh + ud, l), u)                  # admm_update_z :176-190
        ud = ud + (zh - z)                                         # admm_update_u :192-203
        r = np.linalg.norm(ax - z)                                 # admm_r_norm :205-216
        s = np.linalg.norm(-rho * A.T @ (z - z_old))               # admm_s_norm :218-231
        eps_pri = np.sqrt(m) * abstol + reltol * max(np.linalg.norm(ax), np.linalg.norm(z))   # :233-243
        eps_dual = np.sqrt(n) * abstol + reltol * rho * np.linalg.norm(A.T @ ud)              # :245-253
        margins = ((eps_pri - r) / eps_pri, (eps_dual - s) / eps_dual)
        if r < eps_pri and s < eps_dual:                           # :308-309
            status, iters = 0, i + 1
            break
    return dict(x=x, z=z, u=ud, iters=iters, status=status, margins=np.array(margins))


def admm_general_batch(P, q, A, l, u, **kw):
    z0, u0 = kw.pop("z0", None), kw.pop("u0", None)
    outs = [admm_general(P[c], q[c], A[c], l[c], u[c],
                         z0=None if z0 is None else z0[c], u0=None if u0 is None else u0[c], **kw)
            for c in range(P.shape[0])]
    return {k: np.stack([o[k] for o in outs]) for k in outs[0]}
