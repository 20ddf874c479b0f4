import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_prj_b200 import Gmg, GmgConfig
from multigrid_prj_b200.gmg import Timer
for n, L in ((1025, 10), (8193, 13)):
    for tail in (0, 17, 33, 65, 129, 257):
        g = Gmg(GmgConfig.fast(n, L, tail_max_width=tail))
        g.set_rhs_test(1); g.set_u(None); g.run_cycles(6)
        t = Timer(); g.sync(); t.start(g.stream()); g.run_cycles(40, want_relres=False); t.stop(g.stream())
        print(n, "tail", tail, "ms/iter", round(t.elapsed_ms() / 40, 4), flush=True)
        g.close()
