"""CPU: the C-ABI shared library loads, exports every symbol the header declares,
and rejects API misuse without touching a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "qp_batch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(qp_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_and_library_agree(qp):
    lib = qp.load_library()
    declared = header_functions()
    assert declared, "no functions parsed from include/qp_batch.h"
    missing = [f for f in declared if not hasattr(lib, f)]
    assert not missing, f"declared in qp_batch.h but not exported: {missing}"
    assert sorted(qp.ABI_SYMBOLS) == declared


def test_default_options_are_the_reference_constants(qp):
    from embedded_qp_solver_b200 import api
    o = api.qp_options()
    qp.load_library().qp_options_default(C.byref(o))
    # rho, alpha: qp_solvers.c:281-282; ABSTOL, RELTOL: :17-18; ADMM_ITERATIONS: config.h:38
    assert (o.rho, o.alpha, o.max_iter) == (1.0, 1.0, 10000)
    assert o.abstol == np.float32(1e-4) and o.reltol == np.float32(1e-2)
    assert o.device == -1 and o.device_mask == 0


def test_struct_layouts_match_header(qp, tmp_path):
    """The ctypes mirror against what a C compiler makes of include/qp_batch.h: size of every
    struct and offset of every field."""
    import subprocess
    from embedded_qp_solver_b200 import api
    structs = {"qp_batch": api.qp_batch, "qp_batch_solution": api.qp_batch_solution,
               "qp_options": api.qp_options, "qp_general_batch": api.qp_general_batch,
               "qp_general_solution": api.qp_general_solution,
               "struct qp_ref_matrix": api.qp_ref_matrix,
               "struct qp_ref_quadratic_form": api.qp_ref_quadratic_form}
    rename = {"lam": "lambda"}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "qp_batch.h"', 'int main(void){']
    for cname, st in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {rename.get(fname, fname)}));')
    lines.append('printf("abi %d\\n", QP_B200_ABI_VERSION); return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)],
                   check=True)
    got = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True,
                                                        check=True).stdout.strip().split("\n"))
    for cname, st in structs.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"
    assert int(got["abi"]) == 2


def test_api_misuse_returns_error_codes(qp):
    from embedded_qp_solver_b200 import api
    lib = qp.load_library()
    pb, sol = api.qp_batch(), api.qp_batch_solution()
    assert lib.qp_solve_batch(None, C.byref(sol), None) == api.QP_ERR_INVALID
    pb.n, pb.count = 0, 1
    assert lib.qp_solve_batch(C.byref(pb), C.byref(sol), None) == api.QP_ERR_INVALID
    pb.n = api.QP_MAX_DIM_BOX + 1
    assert lib.qp_solve_batch(C.byref(pb), C.byref(sol), None) == api.QP_ERR_UNSUPPORTED
    pb.n = 16  # null P with count > 0
    assert lib.qp_solve_batch(C.byref(pb), C.byref(sol), None) == api.QP_ERR_INVALID
    pb.count = 0  # an empty batch is fine, even without a GPU
    assert lib.qp_solve_batch(C.byref(pb), C.byref(sol), None) == api.QP_OK
    P = np.zeros((1, 16, 16), np.float32)
    q = np.zeros((1, 16), np.float32)
    with pytest.raises(qp.QpError):
        qp.solve_batch(P, q, options=qp.Options(rho=0.0))
    with pytest.raises(qp.QpError):
        qp.solve_batch(P, q, options=qp.Options(alpha=2.5))


def test_kernel_selection(qp):
    assert qp.kernel_for(4) == "tpq<4>" and qp.kernel_for(16) == "tpq<16>"
    assert qp.kernel_for(16, aligned16=False) == "gpq<16>"
    assert qp.kernel_for(40) == "bpq<40>" and qp.kernel_for(64) == "gpq<64>"
    assert qp.kernel_for(40, aligned16=False) == "gpq<40>" and qp.kernel_for(33) == "gpq<40>"
    assert qp.kernel_for(24) == "bpq<24>" and qp.kernel_for(32) == "bpq<32>" and qp.kernel_for(48) == "gpq<48>"
    assert qp.kernel_for(13) == "gpq<16>" and qp.kernel_for(65) == "cpq<80>" and qp.kernel_for(129) == "unsupported"


def test_roofline_bookkeeping(qp):
    # 152 / 1352 / 7208 / 17672 bytes without z; +4n with z
    assert [qp.algorithmic_bytes(n, with_z=False) for n in (4, 16, 40, 64)] == \
        [152, 1352, 7208, 17672]
    assert qp.algorithmic_bytes(64, warm=True, with_z=False) == 18184
    # benchmark configuration: scalar box (nothing to load), z written
    assert qp.algorithmic_bytes(16, bound_arrays=False) == 1288
    assert round(qp.algorithmic_flops(16, 12)) == 21675


def test_missing_library_fails_loudly(qp, tmp_path):
    from embedded_qp_solver_b200 import api
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        api.load_library(str(tmp_path / "nope.so"))


def test_header_is_plain_c99_and_matches_the_ctypes_mirror(qp, tmp_path):
    """include/qp_batch.h compiled as C99 by gcc (a C host binds it directly): struct sizes
    and field offsets the compiler sees are the ones the ctypes mirror uses."""
    import shutil
    import subprocess
    from embedded_qp_solver_b200 import api
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "probe.c"
    src.write_text("""
#include <stdio.h>
#include <stddef.h>
#include "qp_batch.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(qp_batch), sizeof(qp_batch_solution),
           sizeof(qp_options), sizeof(qp_general_batch), sizeof(qp_general_solution),
           sizeof(struct qp_ref_matrix), sizeof(struct qp_ref_quadratic_form));
    printf("%zu %zu %zu %zu %zu\\n", offsetof(qp_batch, count), offsetof(qp_batch, z0),
           offsetof(qp_options, stream), offsetof(qp_options, kernel), offsetof(qp_general_batch, A));
    return 0;
}
""")
    exe = tmp_path / "probe"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, "-o", str(exe),
                    str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(v) for v in out[:7]]
    assert sizes == [C.sizeof(api.qp_batch), C.sizeof(api.qp_batch_solution),
                     C.sizeof(api.qp_options), C.sizeof(api.qp_general_batch),
                     C.sizeof(api.qp_general_solution), C.sizeof(api.qp_ref_matrix),
                     C.sizeof(api.qp_ref_quadratic_form)]
    offs = [int(v) for v in out[7:]]
    assert offs == [api.qp_batch.count.offset, api.qp_batch.z0.offset, api.qp_options.stream.offset,
                    api.qp_options.kernel.offset, api.qp_general_batch.A.offset]


@pytest.mark.parametrize("n", [8, 12, 16])
def test_host_packing_of_upper_triangles(n):
    """The packing step of the QP_MEM_HOST path (no GPU involved): record layout, and the
    AVX-512 version (n = 16, when this CPU has it) bit for bit equal to the 128-bit one,
    from even and odd record offsets and for odd counts."""
    import numpy as np
    from embedded_qp_solver_b200.api import load_library, QP_OK
    lib = load_library()
    rng = np.random.default_rng(n)
    count = 37
    P = rng.standard_normal((count, n, n)).astype(np.float32)
    pk = (n * (n + 1) // 2 + 3) // 4 * 4
    iu = np.triu_indices(n)
    want = np.zeros((count, pk), np.float32)
    want[:, : len(iu[0])] = P[:, iu[0], iu[1]]

    def run(first, cnt, wide):
        raw = np.full(count * pk + 32, np.nan, np.float32)
        off = (-raw.ctypes.data // 4) % 16  # 64-byte aligned start of record 0
        out = raw[off: off + count * pk].reshape(count, pk)
        rc = lib.qp_host_pack_upper(P[first:].ctypes.data, out[first:].ctypes.data, cnt, n, wide)
        # nothing outside the records asked for is written
        assert np.isnan(out[:first]).all() and np.isnan(out[first + cnt:]).all()
        assert np.isnan(raw[:off]).all() and np.isnan(raw[off + count * pk:]).all()
        return rc, out[first: first + cnt].copy()

    rc, got = run(0, count, 0)
    assert rc == QP_OK and np.array_equal(got, want)
    rc, got = run(0, count, -1)
    assert rc == QP_OK and np.array_equal(got, want)
    if n == 16:
        for wide in (1, 2):
            for first, cnt in [(0, count), (1, count - 1), (1, 6), (2, 5), (3, 1), (0, 0)]:
                rc, got = run(first, cnt, wide)
                if rc != QP_OK:
                    pytest.skip("no AVX-512 on this CPU")
                assert np.array_equal(got, want[first: first + cnt]), (wide, first, cnt)


def test_active_set_helper_follows_the_sign_of_lambda():
    """qp.active_set (outputs the reference lacks): +1 on the upper bound, -1 on the
    lower one, the convention of lambda = rho u in include/qp_batch.h; host logic only."""
    import numpy as np
    import embedded_qp_solver_b200 as qp
    z = np.array([[0.5, -0.5, 0.1, 0.5], [-0.25, 0.0, 0.25, -0.5]], np.float32)
    assert np.array_equal(qp.active_set(z, -0.5, 0.5), [[1, -1, 0, 1], [0, 0, 0, -1]])
    lb = np.full_like(z, -0.25)
    ub = np.array([0.5, 0.5, 0.25, 0.5], np.float32)
    assert np.array_equal(qp.active_set(z, lb, ub), [[1, 0, 0, 1], [-1, 0, 1, 0]])
    import torch
    assert torch.equal(qp.active_set(torch.from_numpy(z), -0.5, 0.5),
                       torch.tensor([[1, -1, 0, 1], [0, 0, 0, -1]], dtype=torch.int8))


def test_python_mirror_constants_match_the_header():
    """Every #define of include/qp_batch.h that the ctypes mirror restates (kernel families,
    schedules, precisions, statuses, memory kinds, layouts, limits) has the header's value."""
    import os
    import re
    import embedded_qp_solver_b200 as qp
    from embedded_qp_solver_b200 import api
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "include", "qp_batch.h")).read()
    defs = {k: int(v, 0) for k, v in re.findall(r"#define\s+(QP_[A-Z0-9_]+)\s+\(?(-?(?:0x)?[0-9a-fA-F]+)u?\)?\s", src)}
    checked = 0
    for name, value in defs.items():
        short = name[3:]  # QP_KERNEL_TPQ -> KERNEL_TPQ
        for holder in (qp, api):
            if hasattr(holder, short):
                assert getattr(holder, short) == value, name
                checked += 1
                break
            if hasattr(holder, name):
                assert getattr(holder, name) == value, name
                checked += 1
                break
    assert checked >= 15, (checked, sorted(defs))
    for must in ("QP_KERNEL_GEN_SMEM", "QP_KERNEL_GEN_REG", "QP_KERNEL_CPQ", "QP_SCHEDULE_PASSES"):
        assert must in defs
