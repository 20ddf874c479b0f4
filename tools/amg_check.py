"""Multi-rank parity check of the row-block sharded AMG path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tools/amg_check.py [--side 301] [--levels 4]

Every rank builds the same hierarchy, keeps its row blocks, and ALSO runs the same problem unsharded on its own
GPU; after every operator and after the whole pass the rows a rank owns must equal the unsharded result bit for
bit (same kernels, same colouring, same arithmetic per row; ghost entries refreshed after every colour).  Only the
norms differ in the last bits (order of the partial sums).  Rendezvous uses gloo, so the only NCCL traffic is the
library's own ghost exchange / all-gather / all-reduce.
"""
import argparse
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from multigrid_prj_b200 import Amg                      # noqa: E402
from multigrid_prj_b200 import amg as M                 # noqa: E402
from multigrid_prj_b200 import gmg as G                 # noqa: E402
from amg_bench import synthetic_system                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=301)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--min-rows", type=int, default=2500)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    A, rhs = synthetic_system(a.side)
    n = A.shape[0]
    ok = True

    def report(name, cond, extra=""):
        nonlocal ok
        flags = [None] * world
        dist.all_gather_object(flags, bool(cond))
        ok = ok and all(flags)
        if rank == 0:
            print(("PASS " if all(flags) else "FAIL ") + name, extra, flags if not all(flags) else "", flush=True)

    def new_id():
        ids = [G.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    rng = np.random.default_rng(3)
    x0 = rng.standard_normal(n)
    for exact in (0, 1):
        kw = dict(levels=a.levels, fast=True, device=local, exact_order=exact)
        one = Amg(A.indptr, A.indices, A.data, rhs, **kw)
        sh = Amg(A.indptr, A.indices, A.data, rhs, rank=rank, n_ranks=world, nccl_id=new_id(), shard_min_rows=a.min_rows, **kw)
        tag = "exact-order" if exact else "fast"
        lay = [sh.rows(l) for l in range(a.levels)]
        if rank == 0:
            print(f"[{tag}] levels:", [(one.info(l)["n"], "sharded" if lay[l][2] else "replicated") for l in range(a.levels)], flush=True)
        report(f"[{tag}] level 0 is sharded, the last level is replicated", lay[0][2] and not lay[-1][2])

        def own(v, l):
            r0, rows, _ = lay[l]
            return v[r0:r0 + rows]

        def same(l, which=0):
            return np.array_equal(own(sh.vector(l, which), l), own(one.vector(l, which), l))

        for h in (one, sh):
            h.set_vector(0, 0, x0)
        for h in (one, sh):
            h.smooth(0, M.GS_MULTICOLOUR, 3)
        report(f"[{tag}] 3 multicolour GS sweeps on level 0", same(0))
        for h in (one, sh):
            h.smooth(0, M.JACOBI, 2)
        report(f"[{tag}] 2 Jacobi sweeps on level 0", same(0))
        r1, r2 = one.residual(0), sh.residual(0)
        report(f"[{tag}] residual norm", abs(r1 - r2) <= 1e-12 * r1, f"{r1:.15e} {r2:.15e}")
        report(f"[{tag}] residual vector", same(0, 2))
        for l in range(1, a.levels):
            for h in (one, sh):
                h.restrict(l)
            report(f"[{tag}] restriction to level {l}", same(l))
            for h in (one, sh):
                h.smooth(l, M.GS_MULTICOLOUR, 2)
            report(f"[{tag}] 2 multicolour GS sweeps on level {l}", same(l))
        for l in range(a.levels - 2, -1, -1):
            for h in (one, sh):
                h.prolong(l)
            report(f"[{tag}] prolongation to level {l}", same(l))
            for h in (one, sh):
                h.smooth(l, M.GS_MULTICOLOUR, 1)
            report(f"[{tag}] post-sweep on level {l}", same(l))
        # the reference's whole pass (AMG.cpp:277-308) with the multicolour smoother
        for h in (one, sh):
            for l in range(a.levels):
                h.set_vector(l, 0, np.zeros(h.info(l)["n"]))
        q1, q2 = one.apply(), sh.apply()
        report(f"[{tag}] one pass: solution", same(0))
        report(f"[{tag}] one pass: residual", abs(q1 - q2) <= 1e-12 * q1, f"{q1:.12e} {q2:.12e}")
        # correction-scheme V(2,2) cycles
        for h in (one, sh):
            h.set_vector(0, 0, np.zeros(n))
        h1, h2 = one.solve(tol=1e-8, maxit=60), sh.solve(tol=1e-8, maxit=60)
        report(f"[{tag}] correction-scheme solve: history", h1.size == h2.size and np.allclose(h1, h2, rtol=1e-9), f"{h1.size - 1} cycles to {h1[-1] / h1[0]:.2e}")
        report(f"[{tag}] correction-scheme solve: solution", same(0))
        sh.close()
        if not exact:
            # hybrid Gauss-Seidel: one exchange per sweep, Jacobi-like across the cuts -- a different iterate that still converges
            hy = Amg(A.indptr, A.indices, A.data, rhs, rank=rank, n_ranks=world, nccl_id=new_id(), shard_min_rows=a.min_rows, hybrid_gs=1, **kw)
            h3 = hy.solve(tol=1e-8, maxit=60)
            report("[fast] hybrid GS (1 exchange per sweep) converges", h3[-1] <= max(1e-8 * h3[0], 10 * h1[-1]) and h3.size <= h1.size + 3, f"{h3.size - 1} cycles vs {h1.size - 1}")
            hy.close()
            # weighted Jacobi
            wj1 = Amg(A.indptr, A.indices, A.data, rhs, jacobi_omega=0.8, **kw)
            wj2 = Amg(A.indptr, A.indices, A.data, rhs, rank=rank, n_ranks=world, nccl_id=new_id(), shard_min_rows=a.min_rows, jacobi_omega=0.8, **kw)
            for h in (wj1, wj2):
                h.set_vector(0, 0, x0); h.smooth(0, M.JACOBI, 3)
            r0, rows, _ = wj2.rows(0)
            report("[fast] weighted Jacobi (omega 0.8), 3 sweeps", np.array_equal(wj1.vector(0)[r0:r0 + rows], wj2.vector(0)[r0:r0 + rows]))
            wj1.close(); wj2.close()
        one.close()
    if rank == 0:
        print("AMG_CHECK OK" if ok else "AMG_CHECK FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
