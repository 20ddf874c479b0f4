"""CPU study (scipy, no GPU): which smoother should the Galerkin levels of the AMG cycle use?

    python tools/amg_smoother_study.py [--side 501] [--levels 6] [--cycles 12]

Background: on B200 the correction-scheme cycle is bound by the NUMBER of dependent steps,
not by bandwidth -- a multicolour Gauss-Seidel sweep is one launch per colour and the Galerkin levels of the
reference's coarsening need 13-16 colours, while a Jacobi / polynomial sweep is ONE launch of the SELL kernel that
runs at ~90 % of the HBM peak.  This script builds the same hierarchy as the library (its host setup stages through
the C ABI: strength, C/F split, direct interpolation, Galerkin product) on the synthetic triangulation of
tools/amg_bench.py, runs V(nu,nu) correction-scheme cycles with different smoothers on the levels >= 1 (level 0
always multicolour Gauss-Seidel, the coarsest level 20 sweeps of the same smoother) and prints, per variant, the
mean residual reduction per cycle and the number of kernel launches a cycle would take -- the inputs for choosing
the smoother by time to solution.  Multicolour Gauss-Seidel is replayed colour by colour with a first-fit colouring
(the colour count matches the device's Jones-Plassmann colouring to within 1-3).
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from amg_bench import synthetic_system                  # noqa: E402
from multigrid_prj_b200 import load                     # noqa: E402
from multigrid_prj_b200._lib import check               # noqa: E402

_p = lambda a: a.ctypes.data_as(C.c_void_p)


def hierarchy(A, levels):
    lib = load()

    def handle(M):
        ptr, col, val = M.indptr.astype(np.int64), M.indices.astype(np.int64), M.data.astype(np.float64)
        h = C.c_void_p()
        check(lib.mgb_csr_create(M.shape[0], M.shape[1], _p(ptr), _p(col), _p(val), C.byref(h)))
        return h

    def fetch(h):
        nr, nc, nz = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(lib.mgb_csr_info(h, C.byref(nr), C.byref(nc), C.byref(nz)))
        pp, cc, vv = np.zeros(nr.value + 1, np.int64), np.zeros(nz.value, np.int64), np.zeros(nz.value)
        check(lib.mgb_csr_get(h, _p(pp), _p(cc), _p(vv)))
        return sp.csr_matrix((vv, cc, pp), shape=(nr.value, nc.value))

    As, Ps = [A.tocsr()], []
    hA = handle(As[0])
    for _ in range(levels - 1):
        n = As[-1].shape[0]
        mask, nc = np.zeros(n, np.uint8), C.c_size_t()
        check(lib.mgb_amg_select_coarse_nodes(hA, 0.2, -1, _p(mask), C.byref(nc)))
        hP = C.c_void_p()
        check(lib.mgb_amg_build_prolongation(hA, 0.2, _p(mask), C.byref(hP)))
        hC = C.c_void_p()
        check(lib.mgb_amg_build_coarse_matrix(hA, hP, C.byref(hC)))
        Ps.append(fetch(hP)); As.append(fetch(hC))
        lib.mgb_csr_destroy(hA); lib.mgb_csr_destroy(hP)
        hA = hC
    lib.mgb_csr_destroy(hA)
    return As, Ps


def first_fit_colours(A):
    n = A.shape[0]
    colour = -np.ones(n, np.int64)
    ip, ix = A.indptr, A.indices
    for i in range(n):
        used = set(colour[ix[ip[i]:ip[i + 1]]].tolist())
        k = 0
        while k in used:
            k += 1
        colour[i] = k
    return colour


class Level:
    def __init__(self, A):
        self.A = A
        self.d = A.diagonal()
        self.colour = first_fit_colours(A)
        self.ncol = int(self.colour.max()) + 1
        self.rows = [np.nonzero(self.colour == c)[0] for c in range(self.ncol)]
        self.Arows = [A[r] for r in self.rows]
        self.l1 = np.asarray(abs(A).sum(axis=1)).ravel()            # l1-Jacobi diagonal
        # largest eigenvalue of D^-1 A (power iteration) for the polynomial smoother
        v = np.random.default_rng(0).standard_normal(A.shape[0])
        for _ in range(30):
            v = (A @ v) / self.d
            lam = np.linalg.norm(v)
            v /= lam
        self.lam = 1.1 * lam

    def gs(self, x, b):
        for r, Ar in zip(self.rows, self.Arows):
            x[r] = (b[r] - (Ar @ x - self.d[r] * x[r])) / self.d[r]

    def jacobi(self, x, b, omega):
        x += omega * (b - self.A @ x) / self.d

    def l1jacobi(self, x, b):
        x += (b - self.A @ x) / self.l1

    def chebyshev(self, x, b, degree):
        # Chebyshev polynomial in D^-1 A on [lam/4, lam] (the usual smoothing interval), `degree` SpMVs
        lmax, lmin = self.lam, self.lam / 4.0
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho = 1.0 / sigma
        r = (b - self.A @ x) / self.d
        dvec = r / theta
        for k in range(degree):
            x += dvec
            if k + 1 < degree:
                r = r - (self.A @ dvec) / self.d
                rho_new = 1.0 / (2.0 * sigma - rho)
                dvec = rho_new * rho * dvec + 2.0 * rho_new / delta * r
                rho = rho_new


def run(levels_, Ps, b, variant, nu, cycles, coarse_sweeps=20):
    L = len(levels_)
    x = [np.zeros(l.A.shape[0]) for l in levels_]
    rhs = [b] + [None] * (L - 1)
    launches = [0]

    def smooth(l, sweeps):
        lv = levels_[l]
        kind = "gs" if l == 0 else variant
        for _ in range(sweeps):
            if kind == "gs":
                lv.gs(x[l], rhs[l]); launches[0] += lv.ncol
            elif kind.startswith("jacobi"):
                lv.jacobi(x[l], rhs[l], float(kind[6:])); launches[0] += 1
            elif kind == "l1jacobi":
                lv.l1jacobi(x[l], rhs[l]); launches[0] += 1
            elif kind.startswith("cheb"):
                deg = int(kind[4:])
                lv.chebyshev(x[l], rhs[l], deg); launches[0] += 2 * deg          # SpMV + vector update per term

    hist = [np.linalg.norm(b)]
    for _ in range(cycles):
        launches[0] = 0
        for l in range(L - 1):
            smooth(l, nu)
            r = rhs[l] - levels_[l].A @ x[l]; launches[0] += 1
            rhs[l + 1] = Ps[l].T @ r; launches[0] += 2
            x[l + 1][:] = 0.0
        smooth(L - 1, coarse_sweeps)
        for l in range(L - 2, -1, -1):
            x[l] += Ps[l] @ x[l + 1]; launches[0] += 1
            smooth(l, nu)
        hist.append(np.linalg.norm(b - levels_[0].A @ x[0])); launches[0] += 2
    hist = np.array(hist)
    return hist, launches[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=501)
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--cycles", type=int, default=12)
    ap.add_argument("--nu", type=int, default=2)
    a = ap.parse_args()
    A, b = synthetic_system(a.side)
    As, Ps = hierarchy(A, a.levels)
    lv = [Level(M) for M in As]
    print(f"{A.shape[0]} DoF; levels: " + ", ".join(f"{l.A.shape[0]} rows / {l.A.nnz / l.A.shape[0]:.1f} nnz per row / {l.ncol} colours" for l in lv))
    print(f"V({a.nu},{a.nu}) correction scheme, level 0 multicolour GS, levels >= 1 as listed; 20 sweeps on the last level")
    print(f"{'smoother on levels >= 1':28s} {'reduction/cycle':>16s} {'launches/cycle':>15s} {'cycles to 1e-8':>15s} {'launches to 1e-8':>17s}")
    for variant in ("gs", "jacobi0.6", "jacobi0.7", "jacobi0.8", "l1jacobi", "cheb2", "cheb3", "cheb4"):
        hist, launches = run(lv, Ps, b, variant, a.nu, a.cycles)
        tail = hist[max(1, len(hist) // 2):]
        rate = (tail[-1] / tail[0]) ** (1.0 / (len(tail) - 1)) if tail[0] > 0 and len(tail) > 1 else float("nan")
        need = np.log(1e-8) / np.log(rate) if 0 < rate < 1 else float("inf")
        print(f"{variant:28s} {rate:16.3f} {launches:15d} {need:15.1f} {need * launches:17.0f}", flush=True)


if __name__ == "__main__":
    main()
