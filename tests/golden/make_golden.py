"""Regenerates tests/golden/ (run in the build container, where /root/reference exists).

  python tests/golden/make_golden.py

1. The reference's own committed golden runs are re-packed, values untouched:
     GeometricMultigrid/test/{MGGS4.txt,x.mtx}  <= -n 385 -a 1 -w 10 -ml 5 -test 0 -smt 2
     WebInterface/{MGGS4.txt,x.mtx}             <= -n 145 -a 1 -w 10 -ml 5 -test 1 -smt 1
   (text files with 6 significant digits -> float64 arrays in gmg_goldens.npz)
2. Per-operator and whole-solve outputs of the REFERENCE ITSELF (oracle/_ref/libgmgref.so, the
   reference's classes compiled from /root/reference by oracle/Makefile, 1 OpenMP thread) on seeded
   inputs, as raw float64 -> gmg_ref_ops.npz.  The inputs are regenerated in the tests from the
   same seeds, and are also stored so the fixture is self-contained.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = "/root/reference"


def read_vec(path):
    with open(path) as f:
        n = int(f.readline())
        v = np.array([float(x) for x in f.read().split()])
    assert v.size == n, (path, n, v.size)
    return v


def main():
    import oracle
    oracle.build(ref=True)
    g = {
        "n385_hist": read_vec(f"{REF}/GeometricMultigrid/test/MGGS4.txt"),
        "n385_x": read_vec(f"{REF}/GeometricMultigrid/test/x.mtx"),
        "n145_hist": read_vec(f"{REF}/WebInterface/MGGS4.txt"),
        "n145_x": read_vec(f"{REF}/WebInterface/x.mtx"),
    }
    np.savez_compressed(os.path.join(HERE, "gmg_goldens.npz"), **g)

    r = oracle.ref_gmg()
    r.set_threads(1)
    out = {}
    N, W, alpha = 33, 10.0, 1.0
    rng = np.random.default_rng(20261018)
    u0 = rng.standard_normal(N * N)
    b0 = rng.standard_normal(N * N)
    out["ops_N"] = np.array([N]); out["ops_u0"] = u0; out["ops_b0"] = b0
    for level in range(4):
        out[f"gs_l{level}"] = r.sweep(0, N, W, alpha, level, u0.copy(), b0)
        out[f"jacobi_l{level}"] = r.sweep(1, N, W, alpha, level, u0.copy(), b0)
        ss, res, rel = r.residual(N, W, alpha, level, u0, b0)
        out[f"res_l{level}"] = res
        out[f"res_l{level}_sumsq_rel"] = np.array([ss, rel])
    for lc in range(1, 4):
        out[f"prolong_from_l{lc}"] = r.prolong(N, W, alpha, lc, u0.copy())
    # one full cycle each smoother from a random state (test problem 1 rhs)
    b1 = r.rhs(N, W, 1)
    out["rhs_test1_N33"] = b1
    for sm in (0, 1):
        out[f"cycle_smt{sm}"] = r.cycle(N, W, alpha, 4, sm, b1, u0.copy())
    # whole solves: config C1 (257, L=8, GS, test 1) and a Jacobi one at 65
    for name, (n, L, sm, test) in {"c1_257_gs": (257, 8, 0, 1), "n65_jacobi": (65, 5, 1, 1),
                                   "n65_gs_test2": (65, 6, 0, 2)}.items():
        b = r.rhs(n, W, test)
        u, hist, crel = r.solve(n, W, alpha, L, sm, b)
        out[f"{name}_hist"] = hist
        out[f"{name}_coarse_relres"] = crel
        out[f"{name}_u"] = u
    np.savez_compressed(os.path.join(HERE, "gmg_ref_ops.npz"), **out)
    for f in ("gmg_goldens.npz", "gmg_ref_ops.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
