"""The measurement contract of bench.py, checked without a GPU: the reference arm runs on the host cores and prints
one JSON line with the contract's keys, and the committed N=1 line of the round (profiles/) carries every key the
contract names -- roofline, cpu_baseline, e2e with host<->device bytes, launches, clocks -- with sane values."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def test_reference_arm_prints_the_contract_line():
    import oracle
    if oracle.ref_gmg() is None and not os.path.exists(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        pytest.skip("no checker built")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and BASE_KEYS <= set(line)
    assert line["metric"] == "gmg_vcycle_dof_per_s" and line["unit"] == "DoF/s" and line["dtype"] == "f64"
    assert line["value"] > 1e4 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_committed_bench_line_carries_the_contract():
    line = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final_n1.json")))
    assert BASE_KEYS | {"roofline", "clocks"} <= set(line)
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "reference"
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < line["value"]
    assert line["gpu_launches"] > 0 and line["n_gpus"] == 1 and line["warmup"] >= 3
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(line["value"] - 8193 ** 2 * line["steps"] / (line["ms_per_step"] * line["steps"] * 1e-3)) < 1e-6 * line["value"]
    amg = line["amg"]["kernels"]
    assert {"multicolour_gs_sweep", "jacobi_sweep", "residual_norm", "restrict_Rx", "prolong_add_Px"} <= set(amg)
    assert all(0 < k["frac"] < 1.0 for k in amg.values())
