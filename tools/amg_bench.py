"""AMG solve-phase kernels on one B200, on a synthetic unstructured triangulation (BASELINE config 5 shape):

    python tools/amg_bench.py [--side 2001] [--levels 5] [--json out.json]

Mesh: side x side lattice on [0,2]^2, interior nodes jittered by <= 0.2 h (seed 12345), every cell split by a
diagonal chosen by a seeded hash, so connectivity and values are unstructured; P1 stiffness on the interior
nodes with the reference's conventions (vertex quadrature, weights 2*area/3, Dirichlet values lifted into the
right-hand side -- AMG/src/main.cpp:34-117), assembled vectorised on the host (caller side, not the product).
Reports per level-0 kernel: ms, algorithmic GB/s (12*nnz + 28*n per sweep / SpMV) and fraction of
the measured HBM peak; the one-pass cycle of the reference (AMG.cpp:277-308) and correction-scheme cycles.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multigrid_prj_b200 import Amg                      # noqa: E402
from multigrid_prj_b200 import amg as M                 # noqa: E402
from multigrid_prj_b200.gmg import Timer                # noqa: E402


def synthetic_system(side, seed=12345):
    rng = np.random.default_rng(seed)
    h = 2.0 / (side - 1)
    ii, jj = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    x = jj * h
    y = ii * h
    interior = (ii > 0) & (ii < side - 1) & (jj > 0) & (jj < side - 1)
    x = x + np.where(interior, rng.uniform(-0.2, 0.2, x.shape) * h, 0.0)
    y = y + np.where(interior, rng.uniform(-0.2, 0.2, y.shape) * h, 0.0)
    node = (ii * side + jj)
    # two triangles per cell, diagonal by a seeded hash
    c = node[:-1, :-1].ravel()
    a00, a01, a10, a11 = c, c + 1, c + side, c + side + 1
    flip = rng.integers(0, 2, c.size).astype(bool)
    t1 = np.where(flip[:, None], np.stack([a00, a01, a10], 1), np.stack([a00, a01, a11], 1))
    t2 = np.where(flip[:, None], np.stack([a01, a11, a10], 1), np.stack([a00, a11, a10], 1))
    tri = np.concatenate([t1, t2])
    X, Y = x.ravel(), y.ravel()
    px, py = X[tri], Y[tri]                                   # (T, 3)
    area2 = np.abs((px[:, 1] * py[:, 2] - px[:, 2] * py[:, 1]) + (py[:, 0] * px[:, 2] - px[:, 0] * py[:, 2])
                   + (px[:, 0] * py[:, 1] - py[:, 0] * px[:, 1]))
    # gradients of the P1 basis functions
    det = (px[:, 1] - px[:, 0]) * (py[:, 2] - py[:, 0]) - (px[:, 2] - px[:, 0]) * (py[:, 1] - py[:, 0])
    gx = np.stack([py[:, 1] - py[:, 2], py[:, 2] - py[:, 0], py[:, 0] - py[:, 1]], 1) / det[:, None]
    gy = np.stack([px[:, 2] - px[:, 1], px[:, 0] - px[:, 2], px[:, 1] - px[:, 0]], 1) / det[:, None]
    K = (gx[:, :, None] * gx[:, None, :] + gy[:, :, None] * gy[:, None, :]) * area2[:, None, None]   # sum of 3 weights = area2
    bnd = ~interior.ravel()
    dof = np.full(side * side, -1, np.int64)
    dof[~bnd] = np.arange((~bnd).sum())
    rows = np.repeat(tri[:, :, None], 3, 2).ravel()
    cols = np.repeat(tri[:, None, :], 3, 1).ravel()
    vals = K.ravel()
    r = np.sqrt(X * X + Y * Y)
    gval = np.sin(5 * r)                                                              # Utilities.cpp:3-14
    with np.errstate(divide="ignore", invalid="ignore"):
        fval = np.where(r > 0, -5 * (np.cos(5 * r) / r - 5 * np.sin(5 * r)), 0.0)     # Utilities.cpp:16-22
    n = int((~bnd).sum())
    keep = (~bnd[rows]) & (~bnd[cols])
    A = sp.coo_matrix((vals[keep], (dof[rows[keep]], dof[cols[keep]])), shape=(n, n)).tocsr()
    A.sum_duplicates(); A.sort_indices()
    rhs = np.zeros(n)
    lift = (~bnd[rows]) & bnd[cols]
    np.add.at(rhs, dof[rows[lift]], -vals[lift] * gval[cols[lift]])
    w = np.repeat(area2 / 3.0, 3)
    tn = tri.ravel()
    ok = ~bnd[tn]
    np.add.at(rhs, dof[tn[ok]], fval[tn[ok]] * w[ok])
    return A, rhs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=2001)
    ap.add_argument("--levels", type=int, default=5)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    peak = 6555.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    t0 = time.time()
    A, rhs = synthetic_system(a.side)
    n, nnz = A.shape[0], A.nnz
    print(f"mesh {a.side}^2 nodes -> {n} interior DoF, {nnz} nnz ({nnz / n:.2f}/row), assembled in {time.time() - t0:.1f} s", flush=True)
    out = {"n": n, "nnz": nnz, "levels": a.levels, "peak_gbs": peak, "kernels": {}}
    t0 = time.time()
    amg = Amg(A.indptr, A.indices, A.data, rhs, levels=a.levels, fast=True)
    out["setup_s"] = time.time() - t0
    info = [amg.info(l) for l in range(a.levels)]
    out["hierarchy"] = info
    print(f"setup (host O(nnz) + upload + device colouring): {out['setup_s']:.1f} s")
    for l, i in enumerate(info):
        print(f"  level {l}: n={i['n']} nnz={i['nnz_a']} colours={i['colours']} wavefronts={i['wavefronts']}")
    tm, st = Timer(), amg.stream()

    def bench(name, fn, reps=a.reps):
        fn(); amg.sync(); amg.reset_stats()
        tm.start(st)
        for _ in range(reps):
            fn()
        tm.stop(st)
        ms = tm.elapsed_ms() / reps
        s = amg.stats()
        gbs = s["bytes_algorithmic"] / reps / (ms * 1e-3) / 1e9
        out["kernels"][name] = {"ms": ms, "alg_GBs": gbs, "frac_of_peak": gbs / peak, "launches": s["kernel_launches"] / reps}
        print(f"{name:44s} {ms:9.4f} ms {gbs:9.1f} GB/s alg  {100 * gbs / peak:5.1f}% of peak  {s['kernel_launches'] / reps:5.1f} launches", flush=True)

    lib, h = amg.lib, amg.h
    bench("L0 multicolour GS sweep (colour-sorted SELL-32)", lambda: amg.smooth(0, M.GS_MULTICOLOUR, 1))
    bench("L0 Jacobi sweep (natural-order SELL-32)", lambda: amg.smooth(0, M.JACOBI, 1))
    bench("L0 residual r=b-Ax + norm (natural-order SELL-32)", lambda: lib.mgb_amg_residual(h, 0, __import__('ctypes').byref(__import__('ctypes').c_double())))
    bench("restrict L0->L1 (R=P^T, SELL-32 gather)", lambda: amg.restrict(1))
    bench("prolong-add L1->L0", lambda: amg.prolong(0))
    amg.set_vector(0, 0, np.zeros(n))
    bench("one pass (10 pre / 200 coarse / 10 post sweeps), multicolour GS", lambda: amg.apply(False), reps=3)
    amg.set_vector(0, 0, np.zeros(n))
    for l in range(1, a.levels):
        amg.set_vector(l, 0, np.zeros(info[l]["n"]))
    res = amg.apply()
    out["residual"] = {"before": float(np.linalg.norm(rhs)), "after_one_pass_multicolour": res}
    print(f"residual: {np.linalg.norm(rhs):.4e} -> {res:.4e} after one pass (multicolour GS)")
    # beyond the reference: correction-scheme V(2,2) cycles to 1e-8 (its own pass is not an iteration and diverges here)
    amg.set_vector(0, 0, np.zeros(n))
    amg.sync(); amg.reset_stats()
    t0 = time.perf_counter()
    hist = amg.solve(tol=1e-8, maxit=100)
    dt = time.perf_counter() - t0
    cyc = hist.size - 1
    out["correction_scheme"] = {"cycles": cyc, "seconds": dt, "ms_per_cycle": 1e3 * dt / max(cyc, 1),
                                "dof_cycles_per_s": n * cyc / dt, "residual": [float(hist[0]), float(hist[-1])],
                                "mean_reduction_per_cycle": float((hist[-1] / hist[0]) ** (1 / max(cyc, 1)))}
    print(f"correction-scheme V(2,2), multicolour GS: {cyc} cycles to {hist[-1] / hist[0]:.1e} in {dt * 1e3:.1f} ms "
          f"({1e3 * dt / max(cyc, 1):.2f} ms/cycle, {n * cyc / dt / 1e9:.2f} GDoF*cycles/s, reduction {out['correction_scheme']['mean_reduction_per_cycle']:.3f}/cycle)")
    amg.close()
    # (the CPU baseline of these kernels -- the checker's Gauss-Seidel loop on the same matrix -- is timed by bench.py,
    #  the only bench that may execute oracle/: see the `amg.cpu_baseline` object of its JSON line)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
