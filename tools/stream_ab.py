"""A/B of the two generations of the streaming red-black kernel: bit-identical iterates and time per iteration.

    python tools/stream_ab.py [--n 8193] [--levels 13] [--cycles 12] [--exact]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_prj_b200 import Gmg, GmgConfig          # noqa: E402
from multigrid_prj_b200.gmg import Timer               # noqa: E402


def run(n, levels, cycles, impl, exact, reps=10):
    cfg = GmgConfig.fast(n, levels)
    if exact:
        cfg.rb_fast_arith = 0
    tm = Timer()
    with Gmg(cfg) as g:
        g.set_stream_impl(impl)
        g.set_rhs_test(1); g.set_u(None)
        rel = g.run_cycles(cycles)
        cs = g.checksum()
        g.run_cycles(4)
        g.sync()
        tm.start(g.stream()); g.run_cycles(reps, want_relres=False); tm.stop(g.stream())
        ms = tm.elapsed_ms() / reps
        g.fine_leg(); g.fine_leg(); g.sync()
        tm.start(g.stream())
        for _ in range(reps):
            g.fine_leg()
        tm.stop(g.stream())
        leg = tm.elapsed_ms() / reps
    return {"impl": impl, "relres": rel, "checksum": f"{cs:#018x}", "ms_per_iteration": ms, "fine_leg_ms": leg}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8193)
    ap.add_argument("--levels", type=int, default=0)
    ap.add_argument("--cycles", type=int, default=12)
    ap.add_argument("--exact", action="store_true")
    a = ap.parse_args()
    L = a.levels or (a.n - 1).bit_length() - 1
    L = min(L, 14)
    r1 = run(a.n, L, a.cycles, 1, a.exact)
    r2 = run(a.n, L, a.cycles, 2, a.exact)
    print(json.dumps({"n": a.n, "levels": L, "exact": a.exact, "v1": r1, "v2": r2, "match": r1["checksum"] == r2["checksum"]}))
    sys.exit(0 if r1["checksum"] == r2["checksum"] else 1)
