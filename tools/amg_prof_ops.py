"""Level-0 operators of the AMG fast path, a few launches each, for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:k_amg -o out python tools/amg_prof_ops.py [side]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from multigrid_prj_b200 import Amg                      # noqa: E402
from multigrid_prj_b200 import amg as M                 # noqa: E402
from amg_bench import synthetic_system                  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
A, rhs = synthetic_system(side)
amg = Amg(A.indptr, A.indices, A.data, rhs, levels=2, fast=True)
amg.set_vector(0, 0, np.random.default_rng(1).standard_normal(A.shape[0]))
for _ in range(2):
    amg.smooth(0, M.GS_MULTICOLOUR, 1)
    amg.smooth(0, M.JACOBI, 1)
    amg.residual(0)
    amg.restrict(1)
    amg.prolong(0)
amg.sync()
amg.close()
print("done")
