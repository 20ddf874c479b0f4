"""Row-block sharded AMG solve phase on 1/2/4/8 B200 (BASELINE config 5 shape: synthetic unstructured triangulation):

    python tools/amg_scale.py --side 4001 --levels 10                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        tools/amg_scale.py --side 4001 --levels 10 [--json out.json]

Every rank assembles the same system and hierarchy on its host (deterministic), keeps its row blocks on its GPU
(`mgb_amg_create_sharded`) and times, with CUDA events on the library's stream and the MAX over ranks:
level-0 multicolour Gauss-Seidel sweep (ghosts refreshed after every colour = bit-identical to one GPU, and the
hybrid variant with one exchange per sweep), Jacobi sweep, residual + all-reduced norm, restriction, prolongation and
K correction-scheme V(2,2) cycles.  Throughput = level-0 DoF x cycles / time; GB/s = SURVEY 8d algorithmic bytes
(12 nnz + 28 n per sweep, summed over ranks) / time.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from multigrid_prj_b200 import Amg                      # noqa: E402
from multigrid_prj_b200 import amg as M                 # noqa: E402
from multigrid_prj_b200 import gmg as G                 # noqa: E402
from multigrid_prj_b200.gmg import Timer                # noqa: E402
from amg_bench import synthetic_system                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=2001)
    ap.add_argument("--levels", type=int, default=10)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--cycles", type=int, default=10)
    ap.add_argument("--tail-sweep", default="", help="comma-separated tail_max_rows values to time the cycle with (e.g. -1,3000,8000,50000)")
    ap.add_argument("--min-rows", type=int, default=0, help="shard_min_rows (0: library default; the round-1 2-GPU profiles used 16384)")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    peak = 6555.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]

    def new_id():
        if world == 1:
            return None
        ids = [G.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def maxr(v):
        if world == 1:
            return v
        vals = [None] * world
        dist.all_gather_object(vals, float(v))
        return max(vals)

    def sumr(v):
        if world == 1:
            return v
        vals = [None] * world
        dist.all_gather_object(vals, float(v))
        return sum(vals)

    t0 = time.time()
    A, rhs = synthetic_system(a.side)
    n, nnz = A.shape[0], A.nnz
    t_asm = time.time() - t0
    out = {"n": n, "nnz": nnz, "levels": a.levels, "n_gpus": world, "peak_gbs_per_gpu": peak, "shard_min_rows": a.min_rows,
           "kernels": {}}

    def make(**kw):
        kw.setdefault("shard_min_rows", a.min_rows)
        return Amg(A.indptr, A.indices, A.data, rhs, levels=a.levels, fast=True, device=local, rank=rank, n_ranks=world,
                   nccl_id=new_id(), **kw)

    t0 = time.time()
    amg = make()
    out["setup_s"] = maxr(time.time() - t0)
    lay = [(amg.info(l)["n"],) + amg.rows(l) for l in range(a.levels)]
    out["layout"] = [{"n": x[0], "row0": x[1], "rows": x[2], "sharded": x[3]} for x in lay]
    if rank == 0:
        print(f"{world} GPU(s); mesh {a.side}^2 -> {n} DoF, {nnz} nnz; assembled in {t_asm:.1f} s, setup {out['setup_s']:.1f} s", flush=True)
        print("  levels:", [(x[0], "sharded" if x[3] else "replicated") for x in lay], flush=True)
    tm = Timer()

    def bench(h, name, fn, reps=a.reps, units=None):
        st = h.stream()
        fn(); h.sync(); h.reset_stats()
        if dist:
            dist.barrier()
        tm.start(st)
        for _ in range(reps):
            fn()
        tm.stop(st)
        ms = maxr(tm.elapsed_ms() / reps)
        s = h.stats()
        gbs = sumr(s["bytes_algorithmic"]) / reps / (ms * 1e-3) / 1e9
        rec = {"ms": ms, "alg_GBs_total": gbs, "frac_of_peak_per_gpu": gbs / world / peak, "launches_per_rank": s["kernel_launches"] / reps}
        if units:
            rec["dof_per_s"] = units / (ms * 1e-3)
        out["kernels"][name] = rec
        if rank == 0:
            extra = f"  {rec['dof_per_s'] / 1e9:8.2f} GDoF/s" if units else ""
            print(f"{name:58s} {ms:9.4f} ms {gbs:9.1f} GB/s alg ({100 * gbs / world / peak:5.1f}% of peak per GPU){extra}", flush=True)

    nrm = C.c_double()
    bench(amg, "L0 multicolour GS sweep (ghosts after every colour)", lambda: amg.smooth(0, M.GS_MULTICOLOUR, 1), units=n)
    bench(amg, "L0 Jacobi sweep", lambda: amg.smooth(0, M.JACOBI, 1), units=n)
    bench(amg, "L0 residual + norm", lambda: amg.lib.mgb_amg_residual(amg.h, 0, C.byref(nrm)), units=n)
    bench(amg, "restrict L0->L1", lambda: amg.restrict(1))
    bench(amg, "prolong-add L1->L0", lambda: amg.prolong(0))
    amg.set_vector(0, 0, np.zeros(n))
    hist = None

    def cyc(h):
        nonlocal hist
        h.set_vector(0, 0, np.zeros(n))
        hist = h.solve(tol=0.0, maxit=a.cycles)

    # K correction-scheme V(2,2) cycles: wall clock around the call (it reads one norm back per cycle), max over ranks
    def timed_cycles(h, name):
        cyc(h); h.sync()
        if dist:
            dist.barrier()
        h.set_vector(0, 0, np.zeros(n)); h.sync(); h.reset_stats()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        hh = h.solve(tol=0.0, maxit=a.cycles)
        dt = maxr(time.perf_counter() - t0)
        s = h.stats()
        rec = {"cycles": a.cycles, "ms_per_cycle": 1e3 * dt / a.cycles, "dof_cycles_per_s": n * a.cycles / dt,
               "alg_GBs_total": sumr(s["bytes_algorithmic"]) / dt / 1e9, "launches_per_rank_per_cycle": s["kernel_launches"] / a.cycles,
               "residual": [float(hh[0]), float(hh[-1])]}
        out["kernels"][name] = rec
        if rank == 0:
            print(f"{name:58s} {rec['ms_per_cycle']:9.3f} ms/cycle {rec['dof_cycles_per_s'] / 1e9:7.3f} GDoF*cycles/s  "
                  f"residual {hh[0]:.3e} -> {hh[-1]:.3e}", flush=True)

    timed_cycles(amg, "correction-scheme V(2,2) cycle, multicolour GS")
    amg.close()
    for cap in [int(v) for v in a.tail_sweep.split(",") if v]:
        for graph in (0, -1):
            t = make(tail_max_rows=cap, cycle_graph=graph)
            timed_cycles(t, f"V(2,2) cycle, tail_max_rows={cap}, cycle_graph={graph}")
            t.close()
    if a.tail_sweep:
        t = make(coop_sweeps=1)
        bench(t, "L0 multicolour GS sweep, one cooperative launch", lambda: t.smooth(0, M.GS_MULTICOLOUR, 1), units=n)
        timed_cycles(t, "V(2,2) cycle, cooperative whole-sweep launches (coop_sweeps=1)")
        t.close()
    if world > 1:
        hy = make(hybrid_gs=1)
        bench(hy, "L0 hybrid multicolour GS sweep (ghosts once per sweep)", lambda: hy.smooth(0, M.GS_MULTICOLOUR, 1), units=n)
        timed_cycles(hy, "correction-scheme V(2,2) cycle, hybrid GS")
        hy.close()
    if rank == 0 and a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
