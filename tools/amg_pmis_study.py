"""Host study (scipy, no GPU) of the cycle on a PMIS + direct-interpolation hierarchy -- the algorithm of csrc/amg_setup.cu
restated with numpy -- to choose smoothers / cycle shape for the device path:

    python tools/amg_pmis_study.py [--side 301]

Prints the asymptotic reduction per cycle of V(2,2) with l1-Jacobi / Gauss-Seidel / Chebyshev smoothing below level 0,
of more sweeps on the coarse levels, and of a W-shaped visit of the coarse levels.
"""
import argparse
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from amg_bench import synthetic_system        # noqa: E402

EPS = 0.2


def mix32(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & np.uint64(0xffffffff)
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & np.uint64(0xffffffff)
    x ^= x >> np.uint64(16)
    return x


def strength(A):
    B = abs(A).tocsr().copy(); B.setdiag(0); B.eliminate_zeros()
    big = B.max(axis=1).toarray().ravel()
    C = B.tocoo()
    keep = C.data >= EPS * big[C.row]
    return sp.csr_matrix((np.ones(keep.sum()), (C.row[keep], C.col[keep])), shape=A.shape)


def pmis(S, seed):
    n = S.shape[0]
    St = S.T.tocsr()
    N = ((S + St) > 0).astype(np.float64).tocsr()
    lam = np.asarray(St.sum(axis=1)).ravel() if False else np.asarray(S.sum(axis=0)).ravel()     # points that depend on i
    w = lam.astype(np.float64) * 2.0 ** 33 + mix32(np.arange(n) ^ seed).astype(np.float64) * 2.0 + 0.0
    w = w + np.arange(n) * 1e-7          # index tie-break (weights are distinct integers anyway)
    state = np.where(np.asarray(N.sum(axis=1)).ravel() > 0, -1, 0)
    while (state < 0).any():
        und = state < 0
        wu = np.where(und, w, -1.0)
        # max weight over undecided neighbours
        Nc = N.tocoo()
        m = np.zeros(n); np.maximum.at(m, Nc.row, wu[Nc.col])
        newc = und & (w > m)
        state[newc] = 1
        dep = (S @ (state == 1).astype(np.float64)) > 0
        state[(state < 0) & dep] = 0
    return state == 1


def interpolation(A, S, is_c):
    n = A.shape[0]
    cidx = np.cumsum(is_c) - 1
    Sc = S.multiply(sp.csr_matrix(np.ones((n, 1))) @ sp.csr_matrix(is_c.astype(np.float64)[None, :])).tocsr()
    W = A.multiply(Sc).tocsr()
    denom = np.asarray(W.sum(axis=1)).ravel()
    denom[denom == 0] = 1.0
    W = sp.diags(1.0 / denom) @ W
    W = W.tolil()
    P = W.tocsr()[:, is_c].tolil()
    for i in np.flatnonzero(is_c):
        P.rows[i] = [cidx[i]]; P.data[i] = [1.0]
    return P.tocsr()


def hierarchy(A, levels):
    out = []
    for l in range(levels - 1):
        if A.shape[0] <= 16:
            break
        S = strength(A)
        is_c = pmis(S, 12345 + 7919 * l)
        P = interpolation(A, S, is_c)
        out.append((A, P))
        A = (P.T @ A @ P).tocsr()
    out.append((A, None))
    return out


def smoother(kind, A):
    d = A.diagonal()
    if kind == "l1":
        dl1 = np.asarray(abs(A).sum(axis=1)).ravel()
        return lambda x, b: x + (b - A @ x) / dl1
    if kind == "gs":
        Lm = sp.tril(A).tocsr()
        return lambda x, b: x + spla.spsolve_triangular(Lm, b - A @ x, lower=True)
    if kind == "sgs":
        Lm = sp.tril(A).tocsr(); Um = sp.triu(A).tocsr()
        def f(x, b):
            x = x + spla.spsolve_triangular(Lm, b - A @ x, lower=True)
            return x + spla.spsolve_triangular(Um, b - A @ x, lower=False)
        return f
    if kind.startswith("cheb"):
        deg = int(kind[4:])
        dl1 = np.asarray(abs(A).sum(axis=1)).ravel()
        # Chebyshev on D_l1^-1 A, spectrum in (0, 1]: target interval [1/4 (rough), 1]
        lmax, lmin = 1.0, 0.25
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        def f(x, b):
            r = (b - A @ x) / dl1
            sigma = theta / delta
            rho = 1.0 / sigma
            dvec = r / theta
            x = x + dvec
            for _ in range(deg - 1):
                r = (b - A @ x) / dl1
                rho_new = 1.0 / (2 * sigma - rho)
                dvec = rho_new * rho * dvec + 2 * rho_new / delta * r
                x = x + dvec
                rho = rho_new
            return x
        return f
    raise ValueError(kind)


def cycle(H, sm, l, x, b, nu, gamma, coarse_sweeps):
    A, P = H[l]
    if P is None:
        for _ in range(coarse_sweeps):
            x = sm[l](x, b)
        return x
    n1, n2 = nu(l)
    for _ in range(n1):
        x = sm[l](x, b)
    r = b - A @ x
    xc = np.zeros(P.shape[1])
    for _ in range(gamma(l)):
        xc = cycle(H, sm, l + 1, xc, P.T @ r, nu, gamma, coarse_sweeps)
    x = x + P @ xc
    for _ in range(n2):
        x = sm[l](x, b)
    return x


def rate(H, kinds, nu, gamma, coarse_sweeps=20, K=40):
    sm = [smoother(kinds(l), H[l][0]) for l in range(len(H))]
    A = H[0][0]
    rng = np.random.default_rng(1)
    b = rng.standard_normal(A.shape[0])
    x = np.zeros_like(b)
    hist = [np.linalg.norm(b)]
    for _ in range(K):
        x = cycle(H, sm, 0, x, b, nu, gamma, coarse_sweeps)
        hist.append(np.linalg.norm(b - A @ x))
    hist = np.array(hist)
    return (hist[-1] / hist[-11]) ** 0.1, (hist[8] / hist[0]) ** 0.125


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=301)
    ap.add_argument("--levels", type=int, default=8)
    a = ap.parse_args()
    A, _ = synthetic_system(a.side)
    H = hierarchy(A.tocsr(), a.levels)
    print("levels:", [(h[0].shape[0], h[0].nnz) for h in H])
    V22 = lambda l: (2, 2)
    one = lambda l: 1
    cases = {
        "V(2,2) gs on L0, l1 below": (lambda l: "gs" if l == 0 else "l1", V22, one),
        "V(2,2) l1 everywhere": (lambda l: "l1", V22, one),
        "V(2,2) gs everywhere": (lambda l: "gs", V22, one),
        "V(2,2) gs L0, l1 below with 4+4 sweeps below": (lambda l: "gs" if l == 0 else "l1", lambda l: (2, 2) if l == 0 else (4, 4), one),
        "V(2,2) gs L0, cheb3 below (1+1)": (lambda l: "gs" if l == 0 else "cheb3", lambda l: (2, 2) if l == 0 else (1, 1), one),
        "V(2,2) gs L0, cheb2 below (2+2)": (lambda l: "gs" if l == 0 else "cheb2", V22, one),
        "W below level 1 (gamma=2), gs L0, l1 below": (lambda l: "gs" if l == 0 else "l1", V22, lambda l: 2 if l >= 1 else 1),
        "W everywhere, gs L0, l1 below": (lambda l: "gs" if l == 0 else "l1", V22, lambda l: 2),
    }
    for name, (kinds, nu, gamma) in cases.items():
        asym, first = rate(H, kinds, nu, gamma)
        print(f"{name:52s} asymptotic {asym:.3f}   first 8 cycles {first:.3f}", flush=True)


if __name__ == "__main__":
    main()
