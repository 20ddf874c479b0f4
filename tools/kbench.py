"""Per-kernel timing on one B200 (CUDA events on the library's stream).  Development aid:
    python tools/kbench.py [--n 8193] [--levels 13]
Prints one line per operator: ms per launch group, algorithmic GB/s (SURVEY.md 8d bytes), share.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_prj_b200 import Gmg, GmgConfig          # noqa: E402
from multigrid_prj_b200 import gmg as G                # noqa: E402
from multigrid_prj_b200.gmg import Timer               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8193)
    ap.add_argument("--levels", type=int, default=13)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--fast", type=int, default=1)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    cfg = GmgConfig.fast(a.n, a.levels, rb_fast_arith=a.fast)
    g = Gmg(cfg)
    g.set_rhs_test(1); g.set_u(None)
    g.run_cycles(2)
    t = Timer(); st = g.stream()
    out = {}

    def bench(name, fn, reps=a.reps):
        fn(); g.sync(); g.reset_stats()
        t.start(st)
        for _ in range(reps):
            fn()
        t.stop(st)
        ms = t.elapsed_ms() / reps
        s = g.stats()
        gbs = s["bytes_algorithmic"] / reps / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "alg_GBs": gbs, "launches": s["kernel_launches"] / reps}
        print(f"{name:42s} {ms:9.4f} ms  {gbs:9.1f} GB/s alg  {s['kernel_launches'] / reps:6.1f} launches")

    for lvl in (0, 1, 2):
        if lvl >= a.levels:
            break
        bench(f"L{lvl} rb fused 5 sweeps (S=10)", lambda: g.smooth(lvl, G.GS_RB, 5))
        bench(f"L{lvl} rb fused 2 sweeps (S=4)", lambda: g.smooth(lvl, G.GS_RB, 2))
        bench(f"L{lvl} rb fused 1 sweep  (S=2)", lambda: g.smooth(lvl, G.GS_RB, 1))
        bench(f"L{lvl} jacobi 1 sweep", lambda: g.smooth(lvl, G.JACOBI, 1))
        bench(f"L{lvl} residual norm-only", lambda: g.lib.mgb_gmg_residual(g.h, lvl, G.VEC_E, G.VEC_R, 0, None))
    bench("L0 residual + store (U,F -> R)", lambda: g.lib.mgb_gmg_residual(g.h, 0, G.VEC_U, G.VEC_F, 1, None))
    bench("restrict all levels", lambda: g.restrict())
    if a.levels > 1:
        bench("prolong L1 -> L0", lambda: g.prolong(1))
    if a.levels > 2:
        bench("prolong L2 -> L1", lambda: g.prolong(2))
    bench("cycle (no pre-smooth)", lambda: g.cycle())
    bench("driver iteration (run_cycles(1))", lambda: g.run_cycles(1, want_relres=False))
    # coarse part alone: levels >= 3 (everything the fine-level kernels do not cover)
    if a.n > 1025:
        c = Gmg(GmgConfig.fast((a.n - 1) // 8 + 1, a.levels - 3, rb_fast_arith=a.fast))
        c.set_rhs_test(1); c.set_u(None); c.run_cycles(1)
        t2 = Timer(); c.sync()
        t2.start(c.stream())
        c.run_cycles(a.reps * 4, want_relres=False)
        t2.stop(c.stream())
        ms = t2.elapsed_ms() / (a.reps * 4)
        out["coarse cycle from level 3 down"] = {"ms": ms}
        print(f"{'cycle of the (n-1)/8+1 problem (levels 3+)':42s} {ms:9.4f} ms")
    for (cn, cl) in ((129, 7), (257, 8), (513, 9), (1025, 10)):
        c = Gmg(GmgConfig.fast(cn, cl, rb_fast_arith=a.fast))
        c.set_rhs_test(1); c.set_u(None); c.run_cycles(1)
        t2 = Timer(); c.sync()
        t2.start(c.stream())
        c.run_cycles(a.reps * 4, want_relres=False)
        t2.stop(c.stream())
        ms = t2.elapsed_ms() / (a.reps * 4)
        out[f"driver iteration of a {cn}^2 problem"] = {"ms": ms}
        print(f"{'driver iteration of a %d^2 problem' % cn:42s} {ms:9.4f} ms")
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
