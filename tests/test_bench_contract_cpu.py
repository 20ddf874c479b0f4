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


def test_committed_round2_lines_carry_the_contract_and_the_parity_assertions():
    """round 2: the N=1 line (config 5 AMG object, drop-in leg, 16385^2 single-GPU line) and the 8-GPU line (checksums of
    the slabs and of the AMG row blocks equal to one GPU, the transport named)"""
    l1 = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1_final.json")))
    assert BASE_KEYS | {"roofline", "clocks", "dropin", "c4_single_gpu", "amg"} <= set(l1)
    assert l1["n_gpus"] == 1 and l1["config"]["grid"] == 8193 and l1["value"] > 5e10
    r = l1["roofline"]
    assert r["traffic"] and "r02_ncu_fine_leg_full.txt" in r["traffic_source"] and 0.5 < r["hbm_frac_actual"] < 1.0
    assert 0 < l1["e2e"]["value"] < l1["value"] and set(l1["e2e"]["phases_s"]) == {"set_rhs", "set_u+solve", "get_u"}
    d = l1["dropin"]
    assert d["iterations"] == 1000 and 0.5 < d["fraction_of_value"] < 1.0 and d["final_relres"] < 1e-10
    a = l1["amg"]
    assert a["n"] == 15992001 and a["n_gpus"] == 1
    for v in ("multicolour_gs_fine_l1_jacobi_coarse", "l1_jacobi_all_levels"):
        c = a[v]["cycle"]
        assert c["ms_per_cycle"] < 5.0 and 0.3 < c["reduction_per_cycle"] < 0.8 and len(a[v]["levels"]) == 10
    l8 = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n8_p2p.json")))
    assert l8["n_gpus"] == 8 and l8["config"]["grid"] == 16385 and l8["value"] > 3e11
    assert l8["parity"]["match"] is True and l8["parity"]["u_checksum"] == l8["parity"]["single_gpu_checksum"]
    assert "peer stores" in l8["run"].get("slab_exchange", "peer stores")      # key added after this line was measured
    assert all(v["match"] for v in l8["amg"]["parity"].values())
