"""Multi-rank parity check of the slab-decomposed GMG path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/mg_check.py [--size 1025] [--depth 10]

Every rank runs its slab; rank 0 also runs the SAME problem on one rank and the assembled slab
results must equal it bit for bit (the kernels and their arithmetic are identical; only the
order of the norm sums differs).  Rendezvous and result assembly use the gloo backend, so the
only NCCL traffic is the library's own halo exchange / gather / all-reduce.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_prj_b200 import Gmg, GmgConfig                   # noqa: E402
from multigrid_prj_b200 import gmg as G                         # noqa: E402


def assemble(g, n, arr_fn):
    """each rank fills its rows of a zero n x n array; the sum over ranks is the global field"""
    out = np.zeros((n, n))
    arr_fn(out)
    r0, rows = g.rows(0)
    mask = np.zeros((n, n))
    mask[r0:r0 + rows] = 1.0
    t = torch.from_numpy(out * mask)
    dist.all_reduce(t)
    return t.numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1025)
    ap.add_argument("--depth", dest="levels", type=int, default=10)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    n, L = a.n, a.levels
    rng = np.random.default_rng(11)
    u0 = rng.standard_normal((n, n)); b0 = rng.standard_normal((n, n))
    ok = True

    def report(name, cond, extra=""):
        nonlocal ok
        ok = ok and bool(cond)
        if rank == 0:
            print(("PASS " if cond else "FAIL ") + name, extra, flush=True)

    for mode in ("exact", "fast"):
        ids = [G.nccl_unique_id() if rank == 0 else None]      # one id per communicator
        dist.broadcast_object_list(ids, src=0)
        kw = dict(length=10.0, alpha=1.0, rb_fast_arith=int(mode == "fast"))
        cfg = GmgConfig.fast(n, L, device=local, rank=rank, n_ranks=world, nccl_id=ids[0], **kw)
        g = Gmg(cfg)
        ref = Gmg(GmgConfig.fast(n, L, device=local, **kw)) if rank == 0 else None
        # (1) per operator on random data: fused and unfused red-black sweeps, Jacobi, residual
        for name, kind, sweeps in (("rb fused x5", G.GS_RB, 5), ("rb fused x2", G.GS_RB, 2), ("rb fused x1", G.GS_RB, 1),
                                   ("jacobi x3", G.JACOBI, 3)):
            g.set_u(u0); g.set_rhs(b0)
            g.smooth(0, kind, sweeps, sol=G.VEC_U, rhs=G.VEC_F)
            got = assemble(g, n, g.get_u)
            if rank == 0:
                ref.set_u(u0); ref.set_rhs(b0)
                ref.smooth(0, kind, sweeps, sol=G.VEC_U, rhs=G.VEC_F)
                want = ref.get_u()
                report(f"[{mode}] {name}", np.array_equal(got, want), f"maxdiff {np.abs(got - want).max():.3e}")
        g.set_u(u0); g.set_rhs(b0)
        ss = g.residual(0, G.VEC_U, G.VEC_F, store=True)
        got = assemble(g, n, lambda out: out.__setitem__(slice(None), g.get_level(0, G.VEC_R)))
        if rank == 0:
            ref.set_u(u0); ref.set_rhs(b0)
            ss_ref = ref.residual(0, G.VEC_U, G.VEC_F, store=True)
            report(f"[{mode}] residual field", np.array_equal(got, ref.get_level(0, G.VEC_R)))
            report(f"[{mode}] residual norm (all-reduced)", abs(ss - ss_ref) <= 1e-13 * ss_ref, f"{ss} vs {ss_ref}")
        # (2) whole solve of test problem 1
        g.set_rhs_test(1); g.set_u(None)
        hist = g.solve()
        got = assemble(g, n, g.get_u)
        if rank == 0:
            ref.set_rhs_test(1); ref.set_u(None)
            hist_ref = ref.solve()
            want = ref.get_u()
            report(f"[{mode}] solve: cycles", hist.size == hist_ref.size, f"{hist.size} vs {hist_ref.size}")
            report(f"[{mode}] solve: history", hist.size == hist_ref.size and np.allclose(hist, hist_ref, rtol=1e-9))
            report(f"[{mode}] solve: solution bit-identical", np.array_equal(got, want), f"maxdiff {np.abs(got - want).max():.3e}")
            print(f"      exchanges posted by rank 0: {g.lib and g.stats()}", flush=True)
        # (3) the same iterations as CUDA-graph launches (norm all-reduce deferred to the read at the end)
        g.set_rhs_test(1); g.set_u(None)
        rel = g.run_cycles(6)
        got = assemble(g, n, g.get_u)
        if rank == 0:
            ref.set_rhs_test(1); ref.set_u(None)
            rel_ref = ref.run_cycles(6)
            report(f"[{mode}] run_cycles(6): solution bit-identical", np.array_equal(got, ref.get_u()))
            report(f"[{mode}] run_cycles(6): final residual", abs(rel - rel_ref) <= 1e-9 * rel_ref, f"{rel} vs {rel_ref}")
        # (4) cached graphs against state changes between calls (round-1 advisor finding): run_cycles(3) runs uncaptured and
        # leaves the halo of u valid, run_cycles(4) captures a graph without the leading exchange, set_u(random) invalidates
        # the halo, the next run_cycles must not replay that graph on stale halo rows; same for set_cycle(coarse_maxit)
        g.set_rhs_test(1); g.set_u(None)
        g.run_cycles(3); g.run_cycles(4)
        g.set_u(u0)
        g.run_cycles(6)
        got = assemble(g, n, g.get_u)
        cs = g.checksum()
        if rank == 0:
            ref.set_rhs_test(1); ref.set_u(None)
            ref.run_cycles(3); ref.run_cycles(4)
            ref.set_u(u0)
            ref.run_cycles(6)
            report(f"[{mode}] run_cycles(3); run_cycles(4); set_u(random); run_cycles(6): bit-identical", np.array_equal(got, ref.get_u()))
            report(f"[{mode}] checksum over ranks == single-rank checksum", cs == ref.checksum(), f"{cs:#x}")
        g.lib.mgb_gmg_set_cycle(g.h, cfg.smoother, cfg.restriction, cfg.nu, 0.5, 3)
        g.run_cycles(4)
        got = assemble(g, n, g.get_u)
        if rank == 0:
            ref.lib.mgb_gmg_set_cycle(ref.h, cfg.smoother, cfg.restriction, cfg.nu, 0.5, 3)
            ref.run_cycles(4)
            report(f"[{mode}] set_cycle(coarse_tol, coarse_maxit) reaches cached graphs", np.array_equal(got, ref.get_u()))
        g.lib.mgb_gmg_set_cycle(g.h, cfg.smoother, cfg.restriction, cfg.nu, 0.1, 2000)
        if rank == 0:
            ref.lib.mgb_gmg_set_cycle(ref.h, cfg.smoother, cfg.restriction, cfg.nu, 0.1, 2000)
        # (5) the cycle options beyond the sawtooth on slabs: V / W cycles (operator by operator, NCCL halos) and BiCGSTAB
        # preconditioned by the sawtooth cycle
        for cyc, name in ((G.CYCLE_V, "V(2,5)"), (G.CYCLE_W, "W(2,5)")):
            g.set_cycle_type(cyc, 2, 0); g.set_rhs_test(1); g.set_u(None)
            hist = g.solve(tol=1e-10, maxiter=30)
            got = assemble(g, n, g.get_u)
            if rank == 0:
                ref.set_cycle_type(cyc, 2, 0); ref.set_rhs_test(1); ref.set_u(None)
                hist_ref = ref.solve(tol=1e-10, maxiter=30)
                report(f"[{mode}] {name} cycles: solution bit-identical", hist.size == hist_ref.size and np.array_equal(got, ref.get_u()),
                       f"{hist.size - 1} cycles")
        g.set_cycle_type(G.CYCLE_SAWTOOTH, 0, 0); g.set_rhs_test(1); g.set_u(None)
        hist = g.krylov(G.KRYLOV_BICGSTAB, G.PRECOND_MG, tol=1e-10, maxit=20)
        got = assemble(g, n, g.get_u)
        if rank == 0:
            ref.set_cycle_type(G.CYCLE_SAWTOOTH, 0, 0); ref.set_rhs_test(1); ref.set_u(None)
            hist_ref = ref.krylov(G.KRYLOV_BICGSTAB, G.PRECOND_MG, tol=1e-10, maxit=20)
            want = ref.get_u()
            report(f"[{mode}] BiCGSTAB + sawtooth preconditioner: same steps, solution to 1e-9",
                   hist.size == hist_ref.size and hist[-1] <= 1e-10 and np.linalg.norm(got - want) <= 1e-9 * np.linalg.norm(want),
                   f"{hist.size - 1} steps")
        g.close()
        if ref is not None:
            ref.close()
    flag = torch.tensor([0 if ok else 1])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("MG_CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
