"""ctypes bindings of the CHECKERS (test infrastructure only -- see oracle/README.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing in multigrid_prj_b200/ does.

  oracle.gmg      -> C restatement of the reference algorithm (oracle/gmg_oracle.c)
  oracle.ref_gmg  -> the reference's own classes compiled from /root/reference (oracle/_ref),
                     or None when the prebuilt library is absent.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
_REF_GMG_SO = os.path.join(HERE, "_ref", "libgmgref.so")
_REF_AMG_SO = os.path.join(HERE, "_ref", "libamgref.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the C restatement (always) and, when /root/reference exists, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class _Level(C.Structure):
    _fields_ = [("N", C.c_size_t), ("w", C.c_size_t), ("s", C.c_size_t),
                ("diag", C.c_double), ("off", C.c_double)]


GS, JACOBI, BICGSTAB, RBGS = 0, 1, 2, 3


class GmgOracle:
    """Restatement of GeometricMultigrid/ (fine-sized arrays + stride masks, as the reference)."""

    def __init__(self, path=_ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.gmgo_level_init.argtypes = [C.POINTER(_Level), C.c_size_t, C.c_double, C.c_double, C.c_int]
        L.gmgo_level_init.restype = C.c_int
        L.gmgo_rhs.argtypes = [C.c_size_t, C.c_double, C.c_int, _dp]
        L.gmgo_gs_sweep.argtypes = [C.POINTER(_Level), _dp, _dp]
        L.gmgo_rbgs_sweep.argtypes = [C.POINTER(_Level), _dp, _dp]
        L.gmgo_jacobi_sweep.argtypes = [C.POINTER(_Level), _dp, _dp, _dp]
        L.gmgo_residual.argtypes = [C.POINTER(_Level), _dp, _dp, C.c_void_p]
        L.gmgo_residual.restype = C.c_double
        L.gmgo_sumsq.argtypes = [C.POINTER(_Level), _dp]
        L.gmgo_sumsq.restype = C.c_double
        L.gmgo_prolong.argtypes = [C.POINTER(_Level), C.POINTER(_Level), _dp]
        L.gmgo_cycle_create.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int]
        L.gmgo_cycle_create.restype = C.c_void_p
        L.gmgo_cycle_destroy.argtypes = [C.c_void_p]
        L.gmgo_cycle_set_params.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_double]
        L.gmgo_cycle_apply.argtypes = [C.c_void_p, _dp, _dp]
        L.gmgo_cycle_last_coarse_relres.argtypes = [C.c_void_p]
        L.gmgo_cycle_last_coarse_relres.restype = C.c_double
        L.gmgo_cycle_last_coarse_iters.argtypes = [C.c_void_p]
        L.gmgo_cycle_last_coarse_iters.restype = C.c_long
        L.gmgo_solve.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                 _dp, _dp, C.c_double, C.c_int, _dp, C.c_void_p, C.c_void_p]
        L.gmgo_solve.restype = C.c_int
        L.gmgo_solve_ex.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_long, C.c_double,
                                    _dp, _dp, C.c_double, C.c_int, _dp, C.c_void_p, C.c_void_p]
        L.gmgo_solve_ex.restype = C.c_int
        L.gmgo_cycle_set_restriction.argtypes = [C.c_void_p, C.c_int]

    def level(self, N, length, alpha, level):
        lv = _Level()
        if self.lib.gmgo_level_init(C.byref(lv), N, length, alpha, level) != 0:
            raise ValueError(f"(N-1)={N - 1} is not divisible by 2^{level}")
        return lv

    def rhs(self, N, length, test):
        b = np.empty(N * N)
        self.lib.gmgo_rhs(N, length, test, b)
        return b

    def sweep(self, kind, N, length, alpha, level, sol, b):
        lv = self.level(N, length, alpha, level)
        if kind == JACOBI:
            self.lib.gmgo_jacobi_sweep(C.byref(lv), sol, b, np.zeros_like(sol))
        elif kind == RBGS:
            self.lib.gmgo_rbgs_sweep(C.byref(lv), sol, b)
        else:
            self.lib.gmgo_gs_sweep(C.byref(lv), sol, b)
        return sol

    def residual(self, N, length, alpha, level, sol, b, store=True):
        lv = self.level(N, length, alpha, level)
        res = np.zeros_like(sol) if store else None
        ss = self.lib.gmgo_residual(C.byref(lv), sol, b, res.ctypes.data if store else None)
        return ss, res

    def sumsq(self, N, length, alpha, level, b):
        lv = self.level(N, length, alpha, level)
        return self.lib.gmgo_sumsq(C.byref(lv), b)

    def prolong(self, N, length, alpha, level_coarse, vec):
        c = self.level(N, length, alpha, level_coarse)
        f = self.level(N, length, alpha, level_coarse - 1)
        self.lib.gmgo_prolong(C.byref(c), C.byref(f), vec)
        return vec

    def cycle(self, N, length, alpha, L, kind, b, u, nu=5, coarse_maxit=2000, coarse_tol=0.1,
              restrict_mode=0):
        st = self.lib.gmgo_cycle_create(N, length, alpha, L, kind)
        if not st:
            raise ValueError("bad level count for this N")
        self.lib.gmgo_cycle_set_params(st, nu, coarse_maxit, coarse_tol)
        self.lib.gmgo_cycle_set_restriction(st, restrict_mode)
        self.lib.gmgo_cycle_apply(st, u, b)
        info = (self.lib.gmgo_cycle_last_coarse_relres(st), self.lib.gmgo_cycle_last_coarse_iters(st))
        self.lib.gmgo_cycle_destroy(st)
        return u, info

    def solve(self, N, length, alpha, L, smoother, b, u=None, pre_kind=GS, tol=1e-11, maxiter=1000,
              restrict_mode=0, nu=5, coarse_maxit=2000, coarse_tol=0.1):
        u = np.zeros(N * N) if u is None else u
        hist = np.zeros(maxiter + 1)
        crel = np.zeros(maxiter)
        cits = np.zeros(maxiter, dtype=np.int64)
        n = self.lib.gmgo_solve_ex(N, length, alpha, L, smoother, pre_kind, restrict_mode, nu,
                                   coarse_maxit, coarse_tol, b, u, tol, maxiter, hist,
                                   crel.ctypes.data, cits.ctypes.data)
        if n < 0:
            raise ValueError("bad level count for this N")
        return u, hist[:n].copy(), crel[:n - 1].copy(), cits[:n - 1].copy()


class GmgReference:
    """The reference's own GeometricMultigrid classes (oracle/_ref/libgmgref.so)."""

    def __init__(self, path=_REF_GMG_SO, threads=1):
        L = self.lib = C.CDLL(path)
        L.gmgref_rhs.argtypes = [C.c_size_t, C.c_double, C.c_int, _dp]
        L.gmgref_sweep.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp]
        L.gmgref_residual.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, _dp, _dp,
                                      C.c_void_p, C.POINTER(C.c_double)]
        L.gmgref_residual.restype = C.c_double
        L.gmgref_prolong.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, _dp]
        L.gmgref_solve.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp,
                                   C.c_double, C.c_int, _dp, C.c_void_p]
        L.gmgref_solve.restype = C.c_int
        L.gmgref_cycle.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp]
        L.gmgref_set_threads.argtypes = [C.c_int]
        L.gmgref_openmp_threads.restype = C.c_int
        self.set_threads(threads)

    def set_threads(self, n):
        self.lib.gmgref_set_threads(int(n))

    def threads(self):
        return self.lib.gmgref_openmp_threads()

    def rhs(self, N, length, test):
        b = np.empty(N * N)
        self.lib.gmgref_rhs(N, length, test, b)
        return b

    def sweep(self, kind, N, length, alpha, level, sol, b):
        self.lib.gmgref_sweep(N, length, alpha, level, kind, sol, b)
        return sol

    def residual(self, N, length, alpha, level, sol, b):
        res = np.zeros_like(sol)
        rel = C.c_double()
        ss = self.lib.gmgref_residual(N, length, alpha, level, sol, b, res.ctypes.data, C.byref(rel))
        return ss, res, rel.value

    def prolong(self, N, length, alpha, level_coarse, vec):
        self.lib.gmgref_prolong(N, length, alpha, level_coarse, vec)
        return vec

    def cycle(self, N, length, alpha, L, smoother, b, u):
        self.lib.gmgref_cycle(N, length, alpha, L, smoother, b, u)
        return u

    def solve(self, N, length, alpha, L, smoother, b, u=None, tol=1e-11, maxiter=1000):
        u = np.zeros(N * N) if u is None else u
        hist = np.zeros(maxiter + 1)
        crel = np.zeros(maxiter)
        n = self.lib.gmgref_solve(N, length, alpha, L, smoother, b, u, tol, maxiter, hist,
                                  crel.ctypes.data)
        return u, hist[:n].copy(), crel[:n - 1].copy()


_gmg = None
_ref_gmg = None


def gmg():
    global _gmg
    if _gmg is None:
        _gmg = GmgOracle()
    return _gmg


def ref_gmg():
    """The compiled reference, or None if oracle/_ref was never built (parity then rests on goldens)."""
    global _ref_gmg
    if _ref_gmg is None and os.path.exists(_REF_GMG_SO):
        _ref_gmg = GmgReference()
    return _ref_gmg
