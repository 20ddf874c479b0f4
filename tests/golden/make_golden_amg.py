"""Regenerates tests/golden/amg_*.npz (run where /root/reference exists):  python tests/golden/make_golden_amg.py

For each bundled Gmsh mesh: the P1 system (A, rhs) assembled with the reference's own mesh reader and
element, then -- from the REFERENCE'S OWN AMG classes (oracle/_ref/libamgref.so, start index of the C/F
splitting injected = n/2 on every level, 1 OpenMP thread) -- every level's A_l, P_l, rhs_l, the solution
after AMG::apply_AMG() and the residual norm it prints.  The reference ships no AMG golden vectors.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
os.environ.setdefault("OMP_NUM_THREADS", "1")     # AMG.hpp:314-331 races with more than one thread


def pack(prefix, m, out):
    out[prefix + "_ptr"] = m.ptr.astype(np.int32)
    out[prefix + "_col"] = m.col.astype(np.int32)
    out[prefix + "_val"] = m.val
    out[prefix + "_shape"] = np.array([m.n_rows, m.n_cols])


def main():
    import shutil
    os.makedirs(os.path.join(HERE, "mesh"), exist_ok=True)
    shutil.copyfile("/root/reference/AMG/mesh/mesh1.msh", os.path.join(HERE, "mesh", "mesh1.msh"))   # data fixture of config C2
    import oracle
    oracle.build(ref=True)
    r = oracle.ref_amg()
    for mesh, levels in (("mesh2", 5), ("mesh-pipe", 5), ("mesh1", 5)):
        A, b = r.assemble(f"/root/reference/AMG/mesh/{mesh}.msh")
        out = {"levels": np.array([levels]), "rhs0": b}
        t = time.time()
        r.build(A, b, levels, [-1] * (levels - 1))
        for l in range(levels):
            pack(f"A{l}", r.A(l), out)
            out[f"rhs{l}"] = r.rhs(l)
            if l < levels - 1:
                pack(f"P{l}", r.P(l), out)
        x, res = r.apply()
        out["x_after_pass"] = x
        out["residual_after_pass"] = np.array([res])
        out["residual_before"] = np.array([np.linalg.norm(b)])
        name = f"amg_{mesh.replace('-', '_')}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, os.path.getsize(os.path.join(HERE, name)), "bytes", [r.info(l)[0] for l in range(levels)],
              f"residual {np.linalg.norm(b):.4f} -> {res:.6f}  ({time.time() - t:.1f} s)")


if __name__ == "__main__":
    main()
