"""bench.py prints ONE JSON line per run with the keys the driver reads; both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
             "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
             "cpu_baseline", "gpu_launches"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    """`--impl reference`: the reference's own CPU implementation on the host cores, on the
    same metric and config; no GPU involved."""
    l = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--count", "4096",
                  "--cpu-sample", "4096")
    assert BASE_KEYS <= set(l)
    assert l["impl"] == "reference" and l["higher_is_better"] is True and l["vs_baseline"] is None
    assert l["metric"] == "batched QP solves/sec at n=16,m=32" and l["unit"] == "QP/s"
    assert l["value"] > 0 and l["gpu_launches"] == 0 and "workload" in l["config"]
    cb = l["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == l["value"]
    assert l["e2e"] == {"value": l["value"], "unit": "QP/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_our_arm_line():
    l = run_bench("--steps", "5", "--warmup", "3", "--count", "65536", "--cpu-sample", "20000",
                  "--shape-steps", "2", "--sustained-s", "0.5")
    assert BASE_KEYS | {"roofline", "clocks"} <= set(l)
    assert l["n_gpus"] == 1 and l["steps"] == 5 and l["warmup"] >= 3 and l["gpu_launches"] == 5
    assert l["scaling"] == "weak" and l["dtype"] == "f32" and l["data"] == "synthetic"
    r = l["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "fp32") and 0 < r["frac"] < 1.5
    assert r["hbm"]["unit"] == "GB/s" and r["fp32"]["unit"] == "TFLOP/s"
    e = l["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["matches_device_path"] is True
    assert l["cpu_baseline"]["value"] > 0 and l["cpu_baseline"]["cores"] >= 1
    assert "sm_mhz" in l["clocks"] and "reasons" in l["clocks"]
    assert l["not_converged"] == 0
    assert e["copy_ceiling"] > 0 and 0 < e["frac_of_ceiling"] <= 1.3
    assert e["pageable"]["value"] > 0 and e["pageable"]["matches_device_path"] is True
    su = l["sustained"]
    assert su["seconds"] >= 0.5 and su["value"] > 0 and "sm_mhz" in su["clocks"]
    sh = {d["shape"]: d for d in l["shapes"]}
    assert len(sh) >= 6
    for d in sh.values():
        assert d["qp_per_s"] > 0 and d["not_converged"] == 0 and d["mean_iters"] >= 1
        assert {"hbm_frac", "fp32_reference_frac", "fp32_executed_frac", "issue_frac", "roofline_frac"} <= set(d)
    assert any(k.startswith("n=4,") for k in sh) and any("MPC" in k for k in sh)
    assert any("warm" in k for k in sh) and any(k.startswith("n=64") and "cold" in k for k in sh)
    assert "compiler" in l["cpu_baseline"]
