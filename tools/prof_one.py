"""Runs one operator a few times (for ncu captures):  python tools/prof_one.py rb5|rb2|rb1|jacobi|resid|cycle|fineleg [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_prj_b200 import Gmg, GmgConfig          # noqa: E402
from multigrid_prj_b200 import gmg as G                # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "rb5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8193
L = 13 if what in ("cycle", "fineleg") else 1
g = Gmg(GmgConfig.fast(n, L))
g.set_rhs_test(1); g.set_u(None)
for _ in range(3):
    if what == "rb5": g.smooth(0, G.GS_RB, 5)
    elif what == "rb2": g.smooth(0, G.GS_RB, 2)
    elif what == "rb1": g.smooth(0, G.GS_RB, 1)
    elif what == "jacobi": g.smooth(0, G.JACOBI, 1)
    elif what == "resid": g.residual(0, G.VEC_U, G.VEC_F, store=True)
    elif what == "cycle": g.run_cycles(1)
    elif what == "fineleg": g.fine_leg()
g.sync()
print("done", what, n)
