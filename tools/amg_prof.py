"""One correction-scheme cycle of the AMG fast path for ncu (launch list / full capture), graph off so that every launch
is visible:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/amg_prof.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from multigrid_prj_b200 import Amg                      # noqa: E402
from amg_bench import synthetic_system                  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 10
A, rhs = synthetic_system(side)
amg = Amg(A.indptr, A.indices, A.data, rhs, levels=levels, fast=True, cycle_graph=-1)
amg.set_vector(0, 0, np.zeros(A.shape[0]))
print(amg.solve(tol=0.0, maxit=2))
amg.close()
