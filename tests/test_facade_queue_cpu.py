"""The lazy operator queue of the GMG facade (multigrid_prj_b200/dropin/.../mgb200_gmg_facade.hpp) without a GPU: the facade
headers are compiled against a RECORDING mock of the C ABI (tests/facade_mock/mock_mgb200.c) and the reference's operator
chains are checked for the library calls they turn into -- the driver's iteration must become one mgb_gmg_iterate, and every
other use of the context must first run what is queued, operator by operator, in order."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "facade_mock")
FACADE = os.path.join(ROOT, "multigrid_prj_b200", "dropin", "GeometricMultigrid", "include")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("facade")
    out = str(d / "scenarios")
    subprocess.check_call(["/usr/bin/gcc", "-c", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(MOCK, "mock_mgb200.c"), "-o", str(d / "mock.o")])
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O1", "-w", "-I", FACADE, "-I", os.path.join(ROOT, "include"),
                           os.path.join(MOCK, "scenarios.cpp"), str(d / "mock.o"), "-o", out])
    return out


def calls(exe, what, mode="fast", eager=False):
    env = dict(os.environ, MGB_GMG_MODE=mode)
    env.pop("MGB_FACADE_EAGER", None)
    if eager:
        env["MGB_FACADE_EAGER"] = "1"
    out = subprocess.run([exe, what], capture_output=True, text=True, env=env, timeout=60, check=True).stdout.splitlines()
    body = out[out.index("== begin " + what) + 1:out.index("== end")]
    return [l for l in body if not l.startswith("Achieved residual")], out


def test_driver_iteration_is_one_library_call_in_fast_mode(exe):
    body, out = calls(exe, "driver")
    ops = [l.split()[0] for l in body]
    # the first residual (main.cpp:73) uploads u and runs alone; then per iteration: set_cycle + ONE iterate, nothing else
    assert ops == ["upload", "residual"] + ["set_cycle", "iterate", "norm"] * 3, body
    assert body[3] == "iterate #1" and body[9] == "iterate #3"
    assert sum(l.startswith("Achieved residual on coarse grid: ") for l in out) == 3       # the reference's console line survives


def test_reference_mode_and_eager_mode_run_operator_by_operator(exe):
    for mode, eager in (("reference", False), ("fast", True)):
        body, _ = calls(exe, "driver", mode=mode, eager=eager)
        ops = [l.split()[0] for l in body if not l.startswith("download")]
        per_it = ["smooth", "smooth", "set_cycle", "cycle", "residual", "norm"]
        assert ops == ["upload", "residual"] + per_it * 3, (mode, eager, body)
        assert not any(l.startswith("iterate") for l in body)


@pytest.mark.parametrize("what,sweeps", [("one_sweep", 1), ("three_sweeps", 3)])
def test_other_sweep_counts_are_not_fused(exe, what, sweeps):
    body, _ = calls(exe, what)
    ops = [l.split()[0] for l in body]
    assert ops == ["upload"] + ["smooth"] * sweeps + ["set_cycle", "cycle", "residual"], body


def test_queued_operators_run_before_the_vector_is_read_back(exe):
    body, _ = calls(exe, "sweeps_then_save")
    assert [l.split()[0] for l in body] == ["upload", "smooth", "smooth", "download", "u0"], body
    assert body[-1] == "u0 2"                                   # both queued sweeps were applied before the download
    body, _ = calls(exe, "cycle_then_save")
    assert [l.split()[0] for l in body] == ["upload", "smooth", "smooth", "set_cycle", "cycle", "download", "u0"], body
    assert body[-1] == "u0 102"


def test_another_vector_flushes_the_queue_first(exe):
    body, _ = calls(exe, "other_vector")
    ops = [l.split()[0] for l in body]
    # u's queued chain runs, u is brought home, only then does `other` take the device slot
    i_cycle, i_down, i_up2 = ops.index("cycle"), ops.index("download"), len(ops) - 1 - ops[::-1].index("upload")
    assert i_cycle < i_down < i_up2, body
    assert body[-1] == "u0 102 other0 1", body
    body, _ = calls(exe, "residual_of_other")
    ops = [l.split()[0] for l in body]
    assert "iterate" not in ops and ops.index("cycle") < ops.index("residual"), body
    assert body[-1] == "u0 102"


def test_bicgstab_class_dispatches_to_the_krylov_solver_with_the_reference_console_lines(exe):
    body, out = calls(exe, "bicgstab")
    ops = [l.split()[0] for l in body]
    assert ops[:4] == ["upload", "smooth", "smooth", "Avviamento"], body          # queued sweeps first, then the solver
    assert "krylov method=1 precond=0 maxit=289" in body                          # unpreconditioned, at most `size` steps (17^2)
    assert sum(l.startswith("Norma del residuo: ") for l in body) == 2
    assert body[-2:] == ["Convergenza raggiunta.", "download level=0 which=0"]    # the operator object goes out of scope: u comes home
