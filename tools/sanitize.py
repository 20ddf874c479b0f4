"""Small workloads that touch every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py

GMG: k_rb_stream in every MODE / PIN / EXACT variant (plain sweeps S = 2, 4, 10; fused residual; fused residual +
restriction; fused prolongation; fused correction + norm), the persistent coarse tail, the marching kernels and the
exact wavefront Gauss-Seidel.  AMG: SELL kernels (all modes), the persistent tail, colouring, exact-order kernels.
Sizes are tiny (the sanitizer slows kernels down 10-100x); odd widths exercise the clamped loads at the tile edges.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multigrid_prj_b200 import Amg, Gmg, GmgConfig      # noqa: E402
from multigrid_prj_b200 import amg as M                 # noqa: E402
from multigrid_prj_b200 import gmg as G                 # noqa: E402


def gmg():
    for n, L in ((513, 9), (257, 5), (321, 7)):
        for fast in (1, 0):
            with Gmg(GmgConfig.fast(n, L, rb_fast_arith=fast)) as g:
                g.set_rhs_test(1); g.set_u(None)
                for sw in (1, 2, 5):
                    g.smooth(0, G.GS_RB, sw, sol=G.VEC_U, rhs=G.VEC_F)
                g.solve(maxiter=3)                      # MODE 3 + PIN legs + MODE 1 + tail, uncaptured
                g.run_cycles(6)                         # the same as graph launches
                g.fine_leg()
                g.sync()
        with Gmg(GmgConfig.fast(n, L, fuse_correction=0, fuse_residual=0, fuse_prolong=0, rb_fused=0)) as g:
            g.set_rhs_test(1); g.set_u(None); g.solve(maxiter=2)
        for sm in (G.GS_LEX, G.JACOBI):
            with Gmg(GmgConfig(n=n, levels=L, smoother=sm)) as g:
                g.set_rhs_test(1); g.set_u(None); g.solve(maxiter=2)
        with Gmg(GmgConfig(n=n, levels=L, smoother=G.JACOBI, jacobi_omega=0.8, restriction=G.FULL_WEIGHTING)) as g:
            g.set_rhs_test(2); g.set_u(None); g.solve(maxiter=2)
    print("gmg ok", flush=True)


def amg():
    from amg_bench import synthetic_system
    A, rhs = synthetic_system(61)
    for fast in (True, False):
        for tail in (0, -1):
            with Amg(A.indptr, A.indices, A.data, rhs, levels=4, fast=fast, tail_max_rows=tail) as a:
                a.apply()
                a.solve(tol=1e-6, maxit=5)
                for kind in (M.GS_MULTICOLOUR, M.JACOBI) + (() if fast else (M.GS_LEX,)):
                    a.smooth(0, kind, 2)
                a.restrict(1); a.prolong(0); a.residual(0)
                a.sync()
    print("amg ok", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "gmg"):
        gmg()
    if what in ("all", "amg"):
        amg()
