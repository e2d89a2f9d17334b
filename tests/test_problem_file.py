"""The reference's problem file [double n][P][q] (test/test.c:108-126, qp_ref.py:8-30):
writer/reader of the C ABI pinned byte-for-byte against a file the unmodified reference
wrote (tests/golden/problem_file_n48.bin, captured by tests/golden/make_problem_file.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)


def golden():
    with open(os.path.join(GOLD, "problem_file_n48.bin"), "rb") as f:
        data = f.read()
    z = np.load(os.path.join(GOLD, "problem_file_n48.npz"))
    return data, z["P"], z["q"]


def test_writer_reproduces_the_reference_bytes(qp, tmp_path):
    data, P, q = golden()
    path = tmp_path / "ours.bin"
    qp.write_problem_file(path, P, q)
    assert path.read_bytes() == data
    assert len(data) == 8 * (1 + 48 * 48 + 48)


def test_reader_and_oracle_parser_agree_on_the_reference_file(qp):
    _, P, q = golden()
    path = os.path.join(GOLD, "problem_file_n48.bin")
    P2, q2 = qp.read_problem_file(path)
    assert P2.shape == (1, 48, 48) and np.array_equal(P2[0], P) and np.array_equal(q2[0], q)
    from oracle import qp_ref_np
    (n, P3, q3), = qp_ref_np.read_qp(path)
    assert n == 48 and np.array_equal(P3, P) and np.array_equal(q3, q)


def test_multi_record_files_and_errors(qp, tmp_path):
    rng = np.random.default_rng(3)
    P, q = rng.normal(size=(5, 7, 7)), rng.normal(size=(5, 7))
    path = tmp_path / "batch.bin"
    qp.write_problem_file(path, P[:2], q[:2])
    qp.write_problem_file(path, P[2:], q[2:], append=True)
    P2, q2 = qp.read_problem_file(path)
    assert np.array_equal(P2, P) and np.array_equal(q2, q)
    P3, q3 = qp.read_problem_file(path, first=3, count=2)
    assert np.array_equal(P3, P[3:]) and np.array_equal(q3, q[3:])
    with pytest.raises(qp.QpError) as e:
        qp.read_problem_file(tmp_path / "missing.bin")
    assert e.value.code == -5  # QP_ERR_IO
    (tmp_path / "torn.bin").write_bytes(path.read_bytes()[:-8])
    with pytest.raises(qp.QpError) as e:
        qp.read_problem_file(tmp_path / "torn.bin")
    assert e.value.code == -1  # QP_ERR_INVALID: size is not a whole number of records
    with pytest.raises(qp.QpError):
        qp.read_problem_file(path, first=4, count=2)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference sources")
def test_live_capture_from_the_unmodified_reference(qp, tmp_path):
    """Where the reference can be built: make it write a fresh file now and compare."""
    import make_problem_file
    data, P, q = make_problem_file.capture(seed=7)
    path = tmp_path / "ours.bin"
    qp.write_problem_file(path, P, q)
    assert path.read_bytes() == data


def test_oracle_restatement_of_qp_ref_prints_the_reference_format():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "qp_ref_np.py"),
                          os.path.join(GOLD, "problem_file_n48.bin")],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "reference implementation gave:" and out[1].startswith("f(x_min) = ")
    _, P, q = golden()
    x = np.linalg.solve(P, -q)
    assert abs(float(out[1].split("=")[1]) - (0.5 * x @ P @ x + q @ x)) <= 1e-6 * abs(x @ q)


@pytest.mark.gpu
def test_gpu_qp_ref_drop_in_matches_the_restatement(qp, tmp_path):
    """`python -m embedded_qp_solver_b200.qp_ref FILE` (the drop-in for test/test.c:10) vs
    the numpy restatement, on a well-conditioned file and on the reference's own."""
    from embedded_qp_solver_b200 import problems, qp_ref
    from oracle import qp_ref_np
    P, q, _, _ = problems.wellcond(6, 16, seed=31)
    path = tmp_path / "w.bin"
    qp.write_problem_file(path, P, q)
    _, _, x, f = qp_ref.solve_file(str(path))
    for c, (n, Pc, qc) in enumerate(qp_ref_np.read_qp(path)):
        xs = qp_ref_np.solve_unconstrained(Pc, qc)
        assert np.max(np.abs(x[c] - xs)) <= 1e-4 * (1 + np.max(np.abs(xs)))
        assert abs(f[c] - qp_ref_np.eval_qp(Pc, qc, xs)) <= 1e-5 * (1 + abs(f[c]))
    out = subprocess.run([sys.executable, "-m", "embedded_qp_solver_b200.qp_ref", str(path)],
                         capture_output=True, text=True, check=True, cwd=ROOT).stdout.splitlines()
    assert len(out) == 12 and out[0] == "reference implementation gave:"
