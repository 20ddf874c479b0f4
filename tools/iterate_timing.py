import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from multigrid_prj_b200 import Gmg, GmgConfig
g = Gmg(GmgConfig.fast(8193, 13))
g.set_rhs_test(1); g.set_u(None)
for i in range(4): g.iterate()
g.sync(); t0 = time.perf_counter()
for i in range(100): ss, cr = g.iterate(4e-11)
g.sync(); print("iterate ms", (time.perf_counter() - t0) * 10, g.stats())
g.sync(); t0 = time.perf_counter(); g.run_cycles(100, want_relres=False); g.sync(); print("run_cycles ms", (time.perf_counter() - t0) * 10)
