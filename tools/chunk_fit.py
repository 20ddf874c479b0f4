"""Fixed cost per row chunk of the streaming kernels: time of the two big fine-level launches against the number of
row chunks per strip (MGB_FORCE_NY), one process per setting.  T = waves * (c + X) * tau  ->  X = guarded-step overhead.

    python tools/chunk_fit.py [--n 8193]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8193)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        from multigrid_prj_b200 import Gmg, GmgConfig
        from multigrid_prj_b200 import gmg as G
        from multigrid_prj_b200.gmg import Timer
        L = (a.n - 1).bit_length() - 1
        tm = Timer()
        with Gmg(GmgConfig.fast(a.n, min(L, 14))) as g:
            g.set_rhs_test(1); g.set_u(None); g.run_cycles(3)
            out = {}
            for name, fn in (("fine_leg", lambda: g.fine_leg()),
                             ("pre2", lambda: g.smooth(0, G.GS_RB, 2, sol=G.VEC_U, rhs=G.VEC_F)),
                             ("sweeps5", lambda: g.smooth(0, G.GS_RB, 5, sol=G.VEC_E, rhs=G.VEC_R))):
                fn(); fn(); g.sync()
                tm.start(g.stream())
                for _ in range(10):
                    fn()
                tm.stop(g.stream())
                out[name] = tm.elapsed_ms() / 10
        print(json.dumps(out))
        sys.exit(0)
    res = {}
    for ny in (0, 6, 12, 24, 48, 96):
        env = dict(os.environ, MGB_FORCE_NY=str(ny))
        o = subprocess.check_output([sys.executable, __file__, "--child", "--n", str(a.n)], env=env, text=True)
        res[ny] = json.loads(o.strip().splitlines()[-1])
        print(ny, res[ny], flush=True)
