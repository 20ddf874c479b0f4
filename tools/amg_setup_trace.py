"""Wall-clock of the device-side AMG setup phases:  MGB_TRACE_SETUP=1 python tools/amg_setup_trace.py [side] [levels]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MGB_TRACE_SETUP", "1")
from multigrid_prj_b200 import Amg, System          # noqa: E402
from multigrid_prj_b200 import amg as M             # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 4001
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for rep in range(2):
    t0 = time.perf_counter()
    s = System.synthetic(side)
    t1 = time.perf_counter()
    print(f"assembly {t1 - t0:.3f} s  {s.info()}", flush=True)
    for kw in ({}, {"smoother": M.L1_JACOBI}):
        t0 = time.perf_counter()
        a = Amg.from_system(s, levels=levels, **kw)
        a.sync()
        print(f"setup {kw} {time.perf_counter() - t0:.3f} s, levels {a.levels}", flush=True)
        a.close()
    s.close()
