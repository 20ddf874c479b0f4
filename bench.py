#!/usr/bin/env python
"""bench.py -- GMG V-cycle throughput on B200 (BASELINE.json metric: "V-cycle DoF/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n N] [--levels L]

One "step" = one driver iteration of the reference (GeometricMultigrid/src/main.cpp:84-90):
2 pre-sweeps on the fine grid, one sawtooth multigrid cycle, one residual norm.
  value  = fine DoF x K / device time of K steps, inputs resident in HBM (CUDA events on the
           library's own stream; max over ranks).
  e2e    = the same metric through the C ABI with HOST buffers: upload f and u0 from pinned host
           memory, K steps each reading its residual norm back (the step's result), download u.
  roofline = the dominant kernel (fine-level red-black GS colour pass) timed alone, live.
  cpu_baseline = the reference's own CPU classes (oracle/_ref, compiled from /root/reference) on the
           box's host cores, on a bounded sample of the same workload.
  amg      = BASELINE configs[4] at every N: AMG on the 16 M DoF synthetic unstructured triangulation, assembled and set up
           on the device, fine levels in row blocks over the N GPUs: level-0 kernels (ms, algorithmic GB/s, fraction of
           peak -- the "smoother/SpMV HBM GB/s vs peak" half of the metric), V(2,2) cycles (ms, GDoF*cycles/s, launches,
           convergence factor), checksum of x against a single-GPU repeat.
Workloads: N=1 -> config C3 (8193^2, L=13, BASELINE configs[2]); N>1 -> config C4 (16385^2, L=14, configs[3]) in row
slabs.  The N=1 line also carries `c4_single_gpu` (the 16385^2 grid on one GPU), so scaling can be taken on ONE grid.
  parity (N>1) = rank 0 repeats the same iterations on ONE GPU and the 64-bit checksum of u over all ranks must
           equal the single-GPU one (bit-identical slabs); a mismatch fails the run.
  value is timed with the residual norm all-reduced inside every iteration (as the reference's loop reads it every
           iteration); `value_deferred_norm` is the variant with one all-reduce at the end.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LENGTH, ALPHA, TEST = 10.0, 1.0, 1
METRIC, UNIT = "gmg_vcycle_dof_per_s", "DoF/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def wait_started(self, timeout=5.0):
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.05)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference(n, levels, cycles, threads=None):
    """times the reference's own classes (oracle/_ref) -- the checker, never the product"""
    import oracle
    r = oracle.ref_gmg()
    kind = "reference"
    ncores = os.cpu_count() or 1
    if r is None:
        o = oracle.gmg()
        b = o.rhs(n, LENGTH, TEST)
        t = time.perf_counter()
        o.solve(n, LENGTH, ALPHA, levels, oracle.GS, b, maxiter=cycles, tol=0.0)
        dt = time.perf_counter() - t
        return n * n * cycles / dt, 1, "port", dt
    threads = threads or ncores
    r.set_threads(threads)
    b = r.rhs(n, LENGTH, TEST)
    t = time.perf_counter()
    r.solve(n, LENGTH, ALPHA, levels, 0, b, maxiter=cycles, tol=0.0)
    dt = time.perf_counter() - t
    return n * n * cycles / dt, threads, kind, dt


def amg_cpu_baseline(A, rhs, sweeps=2):
    """the checker's restatement of the reference's Gauss-Seidel loop (AMG/include/Utilities.hpp:44-58) on the host,
    one core, on the same matrix: the CPU number beside the AMG kernels.  Never fails the bench."""
    try:
        import oracle
        o = oracle.amg()
        n = A.shape[0]
        Ao = oracle.Csr(n, n, A.indptr, A.indices, A.data)
        x = np.zeros(n)
        t0 = time.perf_counter()
        o.gs(Ao, rhs, x, sweeps)
        dt = (time.perf_counter() - t0) / sweeps
        return {"gs_sweep_ms": dt * 1e3, "achieved": (12.0 * A.nnz + 28.0 * n) / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                "sample": f"{sweeps} lexicographic Gauss-Seidel sweeps of oracle/amg_oracle.c on the same {n}-row matrix"}
    except Exception as e:                                   # noqa: BLE001 -- a reported baseline, not the product
        return {"error": repr(e)}


def amg_config5(device, peak, side, levels, rank, world, dist, cycles=10, reps=10, cpu=True):
    """BASELINE configs[4]: AMG on a synthetic unstructured 2D triangulation (side^2 nodes; 4001 -> 15 992 001 DoF), the fine
    levels cut into row blocks over `world` GPUs.  The mesh is generated and assembled ON THE DEVICE (mgb_fem_synthetic), the
    hierarchy is built ON THE DEVICE (mgb_amg_config_device: PMIS + direct interpolation + Galerkin products), every rank
    redundantly and deterministically on its own GPU, keeping its row blocks.  Reported: level-0 kernels and whole
    correction-scheme V(2,2) cycles (CUDA events on the library's stream, max over ranks), algorithmic bytes per SURVEY.md 8d
    (12 nnz + 28 n per sweep / SpMV), the convergence factor per cycle, and the checksum of x against a single-GPU repeat."""
    import ctypes as C
    from multigrid_prj_b200 import Amg, System
    from multigrid_prj_b200 import amg as M
    from multigrid_prj_b200 import gmg as G
    from multigrid_prj_b200.gmg import Timer

    def maxr(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def new_id():
        if dist is None:
            return None
        ids = [G.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def barrier():
        if dist is not None:
            dist.barrier()

    t0 = time.perf_counter()
    sysm = System.synthetic(side, device=device)
    n, nnz = sysm.info()
    t_asm = maxr(time.perf_counter() - t0)
    out = {"workload": f"AMG, P1 FEM Poisson on a synthetic unstructured triangulation ({side}^2 nodes): {n} DoF, {nnz} nnz "
                       f"(BASELINE configs[4]), fine levels in row blocks over {world} GPU(s)",
           "n": n, "nnz": nnz, "n_gpus": world, "peak": peak, "unit": "GB/s",
           "assembly_s": t_asm, "assembly": "mesh generated and assembled on the device (mgb_fem_synthetic), every rank its own copy"}
    tm = Timer()

    def variant(name, **kw):
        t0 = time.perf_counter()
        a = Amg.from_system(sysm, levels=levels, rank=rank, n_ranks=world, nccl_id=new_id(), device=device, **kw)
        a.sync()
        setup_s = maxr(time.perf_counter() - t0)
        lay = [dict(a.info(l), rows=a.rows(l)[1], sharded=a.rows(l)[2]) for l in range(a.levels)]
        rec = {"setup_s": setup_s, "ghost_exchange": None if world == 1 else ("NVLink peer stores" if a.lib.mgb_amg_uses_p2p(a.h) else "NCCL send/recv"),
               "levels": [{"n": x["n"], "nnz": x["nnz_a"], "colours": x["colours"], "sharded": x["sharded"]} for x in lay],
               "kernels": {}}
        st = a.stream()

        def bench(kname, fn, r=reps):
            fn(); a.sync(); barrier(); a.reset_stats()
            tm.start(st)
            for _ in range(r):
                fn()
            tm.stop(st)
            ms = maxr(tm.elapsed_ms() / r)
            s_ = a.stats()
            gbs = sumr(s_["bytes_algorithmic"]) / r / (ms * 1e-3) / 1e9
            rec["kernels"][kname] = {"ms": ms, "achieved": gbs, "frac": gbs / world / peak, "launches": s_["kernel_launches"] / r}

        nrm = C.c_double()
        if lay[0]["colours"] > 0:
            bench("L0 multicolour GS sweep", lambda: a.smooth(0, M.GS_MULTICOLOUR, 1))
        bench("L0 l1-Jacobi sweep", lambda: a.smooth(0, M.L1_JACOBI, 2), r=max(reps // 2, 1))
        rec["kernels"]["L0 l1-Jacobi sweep"]["ms"] /= 2; rec["kernels"]["L0 l1-Jacobi sweep"]["launches"] /= 2
        bench("L0 residual + norm", lambda: a.lib.mgb_amg_residual(a.h, 0, C.byref(nrm)))
        bench("restrict L0->L1", lambda: a.restrict(1))
        bench("prolong-add L1->L0", lambda: a.prolong(0))
        # K correction-scheme V(2,2) cycles from x = 0 (one norm read back per cycle, as mgb_amg_solve reports it)
        zero = np.zeros(n)
        a.set_vector(0, 0, zero); a.solve(tol=0.0, maxit=3)                      # warm-up + graph capture
        a.set_vector(0, 0, zero); a.sync(); barrier(); a.reset_stats()
        tm.start(st)
        hist = a.solve(tol=0.0, maxit=cycles)
        tm.stop(st)
        ms = maxr(tm.elapsed_ms())
        s_ = a.stats()
        alg = sumr(s_["bytes_algorithmic"])
        rec["cycle"] = {"cycles": cycles, "ms_per_cycle": ms / cycles, "dof_cycles_per_s": n * cycles / (ms * 1e-3),
                        "achieved": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / world / peak,
                        "launches_per_cycle": s_["kernel_launches"] / cycles, "graph_launches": s_["graph_launches"],
                        "residual": [float(hist[0]), float(hist[-1])],
                        "reduction_per_cycle": float((hist[-1] / hist[0]) ** (1.0 / cycles)),
                        "x_checksum": f"{a.checksum():#018x}"}
        a.close()
        out[name] = rec
        return rec

    v1 = variant("multicolour_gs_fine_l1_jacobi_coarse")                       # config 5's smoother on the fine level
    v2 = variant("l1_jacobi_all_levels", smoother=M.L1_JACOBI)
    out["parity"] = None
    if world > 1:
        # the same cycles on ONE GPU (rank 0): bit-identical iterates expected (Jacobi-type sweeps and per-colour ghost refresh
        # do not depend on the partition)
        res = {}
        if rank == 0:
            for name, kw in (("multicolour_gs_fine_l1_jacobi_coarse", {}), ("l1_jacobi_all_levels", {"smoother": M.L1_JACOBI})):
                with Amg.from_system(sysm, levels=levels, device=device, **kw) as a1:
                    a1.set_vector(0, 0, np.zeros(n)); a1.solve(tol=0.0, maxit=3)
                    a1.set_vector(0, 0, np.zeros(n)); a1.sync()
                    tm.start(a1.stream())
                    a1.solve(tol=0.0, maxit=cycles)
                    tm.stop(a1.stream())
                    res[name] = {"x_checksum": f"{a1.checksum():#018x}", "ms_per_cycle": tm.elapsed_ms() / cycles}
                    res[name]["match"] = res[name]["x_checksum"] == out[name]["cycle"]["x_checksum"]
        barrier()
        out["parity"] = res if rank == 0 else None
    if cpu and rank == 0:
        try:
            import oracle
            ptr, col, val, rhs = sysm.get()
            Ao = oracle.Csr(n, n, ptr, col, val)
            x = np.zeros(n)
            t0 = time.perf_counter()
            oracle.amg().gs(Ao, rhs, x, 1)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"gs_sweep_ms": dt * 1e3, "achieved": (12.0 * nnz + 28.0 * n) / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                   "sample": f"1 lexicographic Gauss-Seidel sweep of oracle/amg_oracle.c (AMG/include/Utilities.hpp:44-58) on the same {n}-row matrix"}
        except Exception as e:                               # noqa: BLE001 -- a reported baseline, not the product
            out["cpu_baseline"] = {"error": repr(e)}
    sysm.close()
    return out


def dropin_leg(n, levels, value):
    """the reference's UNMODIFIED driver (GeometricMultigrid/src/main.cpp, compiled against the facade headers by
    multigrid_prj_b200/dropin/Makefile) on the same problem, fast mode: DoF/s from its own "Solving elapsed time" line.
    Its loop runs until 1e-11 or 1000 iterations (main.cpp:80-90); at 8193^2 the fp64 floor of the problem is 6e-11, so it
    runs all 1000.  The timed region of the driver includes the first-touch upload of f and u from pageable host vectors."""
    import re
    import shutil
    exe = os.path.join(ROOT, "multigrid_prj_b200", "dropin", "_build", "Multigrid")
    if not os.path.exists(exe):
        return {"error": "drop-in driver not built (needs the reference sources at build time)"}
    d = tempfile.mkdtemp(prefix="mgb_dropin_")
    try:
        t0 = time.perf_counter()
        p = subprocess.run([exe, "-n", str(n), "-a", str(int(ALPHA)), "-w", str(int(LENGTH)), "-ml", str(levels), "-test", str(TEST), "-smt", "0"],
                           cwd=d, capture_output=True, text=True, timeout=600, env=dict(os.environ, MGB_GMG_MODE="fast"))
        wall = time.perf_counter() - t0
        if p.returncode != 0:
            return {"error": f"driver exited with {p.returncode}: {p.stderr[-300:]}"}
        m = re.search(r"Solving elapsed time: ([0-9.eE+-]+) sec", p.stdout)
        hist = open(os.path.join(d, "MGGS4.txt")).read().split()
        iters = int(hist[0]) - 1
        secs = float(m.group(1))
        v = float(n) * n * iters / secs
        return {"value": v, "unit": UNIT, "iterations": iters, "solve_seconds": secs, "wall_seconds": wall, "final_relres": float(hist[-1]),
                "fraction_of_value": v / value if value else None,
                "what": f"multigrid_prj_b200/dropin/_build/Multigrid -n {n} -a {int(ALPHA)} -w {int(LENGTH)} -ml {levels} -test {TEST} -smt 0 with "
                        "MGB_GMG_MODE=fast: the reference's own main.cpp; every `u * GS * GS * MG0; u * RES` of its loop is one "
                        "mgb_gmg_iterate call (lazy operator queue of the facade), the norm is read back every iteration"}
    except Exception as e:                                   # noqa: BLE001 -- a reported leg, never fails the bench
        return {"error": repr(e)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference(args):
    """--impl reference: the reference's CPU implementation on the host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, levels = 1025, 10
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, kind, dt = cpu_reference(n, levels, 1)
        if i >= args.warmup:
            vals.append((v, dt))
    value = n * n * len(vals) / sum(d for _, d in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(d for _, d in vals) / len(vals),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args),
        "sample_config": {"ran": f"GMG 2D Poisson {n}x{n}, L={levels}, test {TEST}, lexicographic GS (the reference's own classes)",
                          "why": "bounded sample of the arm's workload: the reference needs ~100 s and 16 GB per iteration at "
                                 "8193^2; its cost per DoF does not depend on the grid size"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"each step = 1 driver iteration (2 GS + sawtooth cycle + residual) of the "
                                   f"reference's GS solver on {n}^2, L={levels}, test {TEST}; throughput per DoF is "
                                   f"size-independent (SURVEY.md section 6: 0.72/0.67 MDoF*cyc/s at 1025^2/2049^2); "
                                   f"only the residual loops are OpenMP-parallel, lexicographic GS is serial"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_name(args):
    n, L = problem(args)
    return (f"GMG 2D Poisson {n}x{n} ({n * n / 1e6:.0f}M DoF), L={L}, test {TEST}, alpha={ALPHA}, W={LENGTH}, "
            f"u0=0; {args.mode} mode")


def config_of(args):
    """the SAME dictionary on both arms (the driver compares them); run-dependent values live outside `config`"""
    n, L = problem(args)
    return {"workload": workload_name(args), "grid": n, "levels": L,
            "baseline_config": "configs[2] (8193^2 on 1 B200)" if n == 8193 else ("configs[3] (16385^2 in row slabs)" if n == 16385 else "custom"),
            "l2": "inputs larger than L2 (fine arrays of 0.5-2 GB vs 126 MB L2)"}


def timed_steps_single(h, timer, steps):
    h.sync()
    timer.start(h.stream())
    h.run_cycles(steps, want_relres=False)
    timer.stop(h.stream())
    return timer.elapsed_ms()


def ncu_traffic(kernel_regex):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the newest committed `ncu --set full` summary under
    profiles/ that names the kernel (a TRAFFIC_BYTES line written next to the capture); None if there is none"""
    import glob
    import re
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_*full*.txt"))):
        txt = open(f, errors="ignore").read()
        m = re.search(r"^TRAFFIC_BYTES\s+([0-9.eE+]+)", txt, re.M)
        if m and re.search(kernel_regex, txt):
            best = (float(m.group(1)), os.path.relpath(f, ROOT))
    return best


def problem(args):
    if args.n:
        n = args.n
        L = args.levels or max(1, int(np.log2(n - 1)))
        return n, L
    return (8193, 13) if args.gpus == 1 else (16385, 14)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--levels", type=int, default=0)
    ap.add_argument("--mode", default="fast", choices=["fast", "parity-jacobi", "parity-gs"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-amg", action="store_true", help="skip the AMG kernel leg (N=1 only)")
    ap.add_argument("--amg-side", type=int, default=4001, help="nodes per side of the synthetic triangulation (4001: config 5)")
    ap.add_argument("--amg-levels", type=int, default=10)
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the single-GPU repeat + checksum comparison")
    ap.add_argument("--no-c4", action="store_true", help="N=1: skip the 16385^2 single-GPU line")
    ap.add_argument("--no-convergence", action="store_true", help="N=1: skip the time-to-tolerance leg (cycle shapes, Krylov)")
    ap.add_argument("--no-dropin", action="store_true", help="N=1: skip the run of the reference's own driver against the facade")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    from multigrid_prj_b200 import Gmg, GmgConfig
    from multigrid_prj_b200 import gmg as G
    from multigrid_prj_b200.gmg import Timer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun", file=sys.stderr)
        return 2
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")

    n, L = problem(args)
    if args.mode == "fast":
        cfg = GmgConfig.fast(n, L, length=LENGTH, alpha=ALPHA, device=local)
    elif args.mode == "parity-jacobi":
        cfg = GmgConfig(n=n, levels=L, length=LENGTH, alpha=ALPHA, smoother=G.JACOBI, device=local)
    else:
        cfg = GmgConfig(n=n, levels=L, length=LENGTH, alpha=ALPHA, smoother=G.GS_LEX, device=local)
    cfg.rank, cfg.n_ranks = rank, world
    if world > 1:
        ids = [G.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        cfg.nccl_id = ids[0]

    def barrier():
        if dist is not None:
            dist.barrier()

    g = Gmg(cfg)
    transport = None if world == 1 else ("NVLink peer stores (pools exported by CUDA IPC; csrc/p2p.cuh)" if g.lib.mgb_gmg_uses_p2p(g.h)
                                         else "NCCL send/recv")
    g.set_rhs_test(TEST)
    g.set_u(None)
    timer = Timer()
    st = g.stream()

    def timed_steps(h, steps):
        """K driver iterations, device-timed on the library's stream, max over ranks"""
        h.sync(); barrier()
        h.reset_stats()
        timer.start(h.stream())
        h.run_cycles(steps, want_relres=False)
        timer.stop(h.stream())
        t_ms = timer.elapsed_ms()
        h.sync(); barrier()
        if dist is not None:
            import torch
            t = torch.tensor([t_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms

    # ---- device-resident throughput ------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_started()
    g.run_cycles(args.warmup)
    g.run_cycles(4)          # 4 more untimed steps: lets the library capture its CUDA graph for this buffer state
    ms = timed_steps(g, args.steps)          # the norm of every iteration is all-reduced inside the loop (defer_norm = 0)
    stats = g.stats()
    total_cycles = args.warmup + 4 + args.steps
    u_checksum = g.checksum()
    relres = g.run_cycles(0)
    dof = float(n) * float(n)
    value = dof * args.steps / (ms * 1e-3)
    deferred = None
    if world > 1:            # named extra: one all-reduce at the end instead of one per iteration
        from multigrid_prj_b200._lib import check
        check(g.lib.mgb_gmg_set_defer_norm(g.h, 1))
        g.run_cycles(4)
        ms_d = timed_steps(g, args.steps)
        check(g.lib.mgb_gmg_set_defer_norm(g.h, 0))
        deferred = {"value": dof * args.steps / (ms_d * 1e-3), "ms_per_step": ms_d / args.steps,
                    "what": "same steps with mgb_gmg_config.defer_norm = 1: every rank keeps its partial sum, one all-reduce when the norm is read"}

    # ---- parity of the slab decomposition: the same iterations on ONE GPU, checksums must agree ------------------
    parity = {"u_checksum": f"{u_checksum:#018x}", "cycles": total_cycles}
    parity_ok = True
    if world > 1 and not args.no_parity:
        if rank == 0:
            c1 = GmgConfig.fast(n, L, length=LENGTH, alpha=ALPHA, device=local) if args.mode == "fast" else None
            if c1 is not None:
                with Gmg(c1) as g1:
                    g1.set_rhs_test(TEST); g1.set_u(None)
                    g1.run_cycles(args.warmup); g1.run_cycles(4); g1.run_cycles(args.steps, want_relres=False)
                    ref_sum = g1.checksum()
                    t1 = timed_steps_single(g1, timer, args.steps)
                parity.update(single_gpu_checksum=f"{ref_sum:#018x}", match=bool(ref_sum == u_checksum),
                              single_gpu_same_grid={"value": dof * args.steps / (t1 * 1e-3), "ms_per_step": t1 / args.steps,
                                                    "what": f"the same {n}^2 grid on one GPU of this box (rank 0), for scaling on one grid"})
                parity_ok = bool(ref_sum == u_checksum)
        barrier()

    # ---- dominant kernel alone ---------------------------------------------------------------------------
    # fast path on one rank: the fused fine-level launch of the upward leg (prolongation from level 1 + nu red-black
    # sweeps + u += err + residual norm), which is the largest entry of the ncu launch list; otherwise the plain
    # fine-level smoothing launch.
    peak, peak_src = peaks()
    reps = 10
    kind = cfg.smoother if cfg.smoother != G.GS_LEX else G.GS_RB
    rows0 = g.rows(0)[1]
    fused_leg = (world == 1 and kind == G.GS_RB and cfg.rb_fused and cfg.fuse_correction and cfg.fuse_prolong and L >= 2)
    group = cfg.nu if (kind == G.GS_RB and cfg.rb_fused) else 1
    if fused_leg:
        run_k = lambda: g.fine_leg()
        gen2 = os.environ.get("MGB_STREAM_IMPL", "2") != "1" and cfg.nu == 5
        kname = (f"k_rb_stream{'2' if gen2 else ''}<{2 * cfg.nu}, EXACT=0, MODE=1, PIN=1>: prolongation + {cfg.nu} red-black sweeps + "
                 f"correction + residual norm in one launch ({cfg.nu} x 24 + 10 + 24 + 16 B/pt algorithmic)"
                 + ("; rows fed by cp.async.bulk on mbarriers, right-hand-side ring in tensor memory, 3 CTAs per SM" if gen2 else ""))
        moved = 26.0 * n * rows0                      # read res 8, u 8, coarse err 2; write u 8
        tr = ncu_traffic(r"k_rb_stream2<10, 0, 1, 1>" if gen2 else r"k_rb_stream<10, 0, 1, 1>") if n == 8193 else None
        traffic, traffic_src = (tr[0], tr[1] + " (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, one launch at 8193^2; "
                                "read from the committed file, not measured in this run)") if tr else (None, None)
    else:
        run_k = lambda: g.smooth(0, kind, sweeps=group, sol=G.VEC_E, rhs=G.VEC_R)
        kname = {G.GS_RB: (f"k_rb_stream<{2 * group}> ({group} fused red-black sweeps per launch = {group} x 24 B/pt algorithmic)")
                 if cfg.rb_fused else "k_rbgs_colour (one colour pass, 12 B/pt)", G.JACOBI: "k_jacobi (24 B/pt)"}[kind]
        moved = 24.0 * n * rows0
        traffic, traffic_src = None, None
    run_k(); run_k()
    g.sync()
    g.reset_stats()
    timer.start(st)
    for _ in range(reps):
        run_k()
    timer.stop(st)
    kms = timer.elapsed_ms()
    clocks = sampler.stop() if sampler else None
    ks = g.stats()
    launches = reps if fused_leg else ks["kernel_launches"]          # the fused leg adds one tiny reduce launch per call
    alg_bytes = ks["bytes_algorithmic"] / launches
    achieved = alg_bytes / (kms * 1e-3 / launches) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "hbm_bytes_moved_per_launch": moved, "hbm_frac_actual": moved / (kms * 1e-3 / launches) / 1e9 / peak,
                "note": "achieved counts SURVEY 8d algorithmic bytes of every operation the launch performs (no credit for "
                        "fusion); the launch is temporally blocked and moves each array through HBM once, so frac > 1 measures "
                        "what fusion saved and hbm_frac_actual is the fraction of HBM bandwidth the launch really uses "
                        "(it is bound by issue slots and shared-memory bandwidth at 12 warps per SM, see DESIGN.md section 5)",
                "kernel": kname, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kms / launches,
                "step_algorithmic_gbs": stats["bytes_algorithmic"] / (ms * 1e-3) / 1e9,
                "step_frac": stats["bytes_algorithmic"] / (ms * 1e-3) / 1e9 / peak}

    # ---- end to end through the C ABI with host buffers ----------------------------------------------
    # every rank owns the pinned host copy of ITS slab of f and u; the ABI takes the address of the global
    # array, so the slab buffer is passed with the offset of its first row subtracted (only slab rows are touched)
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        import torch
        r0, rows = g.rows(0)
        f_host = torch.empty((rows, n), dtype=torch.float64).pin_memory()
        u_host = torch.empty((rows, n), dtype=torch.float64).pin_memory()
        fh, uh = f_host.numpy(), u_host.numpy()
        from multigrid_prj_b200._lib import check
        check(g.lib.mgb_gmg_get_level(g.h, 0, G.VEC_F, C.c_void_p(fh.ctypes.data - r0 * n * 8)))   # the slab of f, from the device
        fptr = C.c_void_p(fh.ctypes.data - r0 * n * 8)
        uptr = C.c_void_p(uh.ctypes.data - r0 * n * 8)
        from multigrid_prj_b200._lib import check
        hist = np.zeros(args.steps + 1)
        nh = C.c_int()
        g.sync(); barrier()
        t0 = time.perf_counter()
        check(g.lib.mgb_gmg_set_rhs(g.h, fptr))
        t1 = time.perf_counter()
        check(g.lib.mgb_gmg_set_u(g.h, None))            # u0 = 0 (main.cpp:49): NULL = zero-fill on the device, nothing to upload
        check(g.lib.mgb_gmg_solve(g.h, 0.0, args.steps, 1, hist.ctypes.data_as(C.c_void_p), C.byref(nh)))
        t2 = time.perf_counter()
        check(g.lib.mgb_gmg_get_u(g.h, uptr))
        t3 = time.perf_counter()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        slab_bytes = float(rows) * n * 8
        e2e = {"value": dof * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": slab_bytes / args.steps, "d2h_bytes_per_step": slab_bytes / args.steps + 8,
               "seconds": dt, "phases_s": {"set_rhs": t1 - t0, "set_u+solve": t2 - t1, "get_u": t3 - t2}, "pcie_floor_note": "the two slab copies alone take ~2 x slab_bytes / 50 GB/s; the solve is PCIe-bound below ~40 steps",
               "what": f"per rank: mgb_gmg_set_rhs (pinned host slab -> HBM), mgb_gmg_set_u(NULL) (u0 = 0: device fill), mgb_gmg_solve with "
                       f"{args.steps} steps each reading its residual norm back, mgb_gmg_get_u (HBM -> pinned host); "
                       f"wall clock, max over ranks; bytes are per rank", "final_relres": float(hist[nh.value - 1])}

    # ---- CPU baseline beside it (rank 0, N=1 only) --------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cn, cl, cc = 2049, 11, 2
        v, cores, ckind, dt = cpu_reference(cn, cl, cc)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": ckind,
               "sample": f"{cc} driver iterations of the reference's GS solver on {cn}^2, L={cl} ({dt:.1f} s); "
                         f"lexicographic GS is serial, only the residual loops use the {cores} OpenMP threads"}

    # ---- time to tolerance (N=1): DoF/s per iteration says nothing about how many iterations an algorithm needs; the cycle
    # shapes and the Krylov solver the library offers on the same handle, each from u = 0 to a relative residual of 1e-9
    convergence = None
    if rank == 0 and world == 1 and args.mode == "fast" and not args.no_convergence:
        convergence = {"tolerance": 1e-9, "what": "wall time of one solve call from u = 0 (norm read back every iteration / step), same grid"}
        try:
            def timed_solve(label, fn):
                g.set_u(None); g.sync()
                t0 = time.perf_counter()
                hist = fn()
                dt = time.perf_counter() - t0
                convergence[label] = {"iterations": int(hist.size - 1), "seconds": dt, "final_relres": float(hist[-1])}
            timed_solve("sawtooth (2 pre-sweeps + cycle, the reference's shape; fast path)", lambda: g.solve(tol=1e-9, maxiter=40))
            g.set_cycle_type(G.CYCLE_V, 2, 0)
            timed_solve("V(2,5) cycles", lambda: g.solve(tol=1e-9, maxiter=40))
            g.set_cycle_type(G.CYCLE_W, 2, 0)
            timed_solve("W(2,5) cycles", lambda: g.solve(tol=1e-9, maxiter=40))
            g.set_cycle_type(G.CYCLE_SAWTOOTH, 0, 0)
            timed_solve("BiCGSTAB preconditioned by the sawtooth cycle", lambda: g.krylov(G.KRYLOV_BICGSTAB, G.PRECOND_MG, tol=1e-9, maxit=40))
        except Exception as e:                               # noqa: BLE001 -- a reported leg, never fails the bench
            convergence["error"] = repr(e)
        finally:
            try:
                g.set_cycle_type(G.CYCLE_SAWTOOTH, 0, 0)
            except Exception:                                # noqa: BLE001
                pass

    # ---- like for like (N=1): the REFERENCE'S algorithm on the GPU (exact lexicographic GS + injection: bit-identical to the
    # reference, tests/test_gmg_gpu.py) on the cpu_baseline's own configuration, so that one ratio compares the same arithmetic
    same_alg = None
    if rank == 0 and world == 1 and cpu is not None:
        try:
            cn, cl, cc = 2049, 11, 2
            with Gmg(GmgConfig(n=cn, levels=cl, length=LENGTH, alpha=ALPHA, smoother=G.GS_LEX, device=local)) as gp:
                gp.set_rhs_test(TEST); gp.set_u(None)
                gp.run_cycles(1)
                tp = timed_steps_single(gp, timer, cc)
            vp = float(cn) * cn * cc / (tp * 1e-3)
            same_alg = {"value": vp, "unit": UNIT, "ms_per_step": tp / cc, "ratio_to_cpu_baseline": vp / cpu["value"],
                        "what": f"parity mode on one B200: {cc} driver iterations of the reference's own algorithm (exact lexicographic "
                                f"Gauss-Seidel as a skewed wavefront, injection, sawtooth) on {cn}^2, L={cl} -- the same arithmetic, bit "
                                "for bit, as the cpu_baseline sample; the fast path differs by the ordering of the smoother and the "
                                "restriction (DESIGN.md section 3) and reaches the same solution within 1e-8"}
        except Exception as e:                               # noqa: BLE001 -- a reported leg, never fails the bench
            same_alg = {"error": repr(e)}

    # ---- AMG: BASELINE configs[4] (16 M DoF, sharded over the N GPUs): kernels against the HBM peak + whole cycles ------------
    amg = None
    amg_ok = True
    if not args.no_amg:
        g.close()                                  # the GMG arrays are not needed any more
        amg = amg_config5(local, peak, args.amg_side, args.amg_levels, rank, world, dist, cpu=not args.no_cpu)
        if amg.get("parity"):
            amg_ok = all(v["match"] for v in amg["parity"].values())

    g.close()
    # ---- N=1 only: the slab runs' grid (config C4, 16385^2) on this one GPU, so that scaling can be taken on one grid -------
    c4 = None
    if rank == 0 and world == 1 and n == 8193 and not args.no_c4 and args.mode == "fast":
        with Gmg(GmgConfig.fast(16385, 14, length=LENGTH, alpha=ALPHA, device=local)) as g4:
            g4.set_rhs_test(TEST); g4.set_u(None)
            g4.run_cycles(args.warmup); g4.run_cycles(4)
            t4 = timed_steps_single(g4, timer, args.steps)
            c4 = {"value": 16385.0 ** 2 * args.steps / (t4 * 1e-3), "ms_per_step": t4 / args.steps, "u_checksum": f"{g4.checksum():#018x}",
                  "cycles": args.warmup + 4 + args.steps,
                  "what": "GMG 2D Poisson 16385x16385, L=14 (BASELINE configs[3]) on ONE B200: the N=1 point of the slab runs' grid"}

    dropin = None
    if rank == 0 and world == 1 and args.mode == "fast" and not args.no_dropin:
        dropin = dropin_leg(n, L, value)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args),
            "scaling_note": "N=1 runs BASELINE configs[2] (8193^2), N>=2 run configs[3] (16385^2 in row slabs): total work is fixed "
                            "among N>=2 (strong); the N=1 line's c4_single_gpu and every N>1 line's parity.single_gpu_same_grid give "
                            "the 16385^2 grid on one GPU for an efficiency on one grid",
            "run": {"smoother": cfg.smoother, "restriction": cfg.restriction, "final_relres": relres,
                    "norm": "all-reduced inside every iteration" if world > 1 else "single rank",
                    "slab_exchange": transport,
                    "extra_warmup": "4 untimed steps after --warmup during which the CUDA graph of the iteration is captured"},
            "roofline": roofline, "cpu_baseline": cpu, "same_algorithm_on_gpu": same_alg, "convergence": convergence, "e2e": e2e,
            "gpu_launches": int(stats["kernel_launches"]), "clocks": clocks, "parity": parity,
            "value_deferred_norm": deferred, "c4_single_gpu": c4, "dropin": dropin, "amg": amg,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if not parity_ok:
        print("bench.py: PARITY FAILURE: the slab-decomposed u differs from the single-GPU u (checksums above)", file=sys.stderr)
        return 3
    if not amg_ok:
        print("bench.py: PARITY FAILURE: the row-block sharded AMG iterate differs from the single-GPU one", file=sys.stderr)
        return 4
    return 0


if __name__ == "__main__":
    sys.exit(main())
