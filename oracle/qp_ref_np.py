"""oracle/qp_ref_np.py -- TEST INFRASTRUCTURE, not product code.

numpy restatement of the reference's secondary oracle test/qp_ref.py, whose arithmetic
lives in the third-party package ``qpsolvers`` (PyPI; no version pinned anywhere in the
reference, not vendored, not installed here, no network): PARITY UNPINNED for that
package.  What is restated is what the script's call means: ``solve_qp(P, q, G=0, h=0)``
(qp_ref.py:35) has an empty feasible-set restriction, i.e. it returns the unconstrained
minimiser of x'Px/2 + q'x, x = -P^-1 q; the parser follows read_qp (qp_ref.py:8-30)
line by line and eval_qp is qp_ref.py:5-6.

    python oracle/qp_ref_np.py FILE        prints what test/qp_ref.py prints
"""
import struct
import sys

import numpy as np


def eval_qp(P, q, x):
    """qp_ref.py:5-6"""
    return .5 * x.T @ P @ x + q @ x


def read_qp(filename):
    """qp_ref.py:8-30: the whole file as doubles, n first, then P row-major, then q.
    (Several records back to back are returned as a list; the reference reads one.)"""
    with open(filename, "rb") as fp:
        data = fp.read()
    vals = struct.unpack("d" * (len(data) // 8), data)
    out, pos = [], 0
    while pos < len(vals):
        n = int(vals[pos])
        P = np.array(vals[pos + 1:pos + 1 + n * n]).reshape(n, n)
        q = np.array(vals[pos + 1 + n * n:pos + 1 + n * n + n])
        out.append((n, P, q))
        pos += 1 + n * n + n
    return out


def solve_unconstrained(P, q):
    """What solve_qp(P, q, G=0, h=0) returns for a strictly convex P."""
    return np.linalg.solve(P, -q)


def main(argv):
    for n, P, q in read_qp(argv[1]):
        x = solve_unconstrained(P, q)
        print("reference implementation gave:")
        print("f(x_min) = %e" % eval_qp(P, q, x))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
