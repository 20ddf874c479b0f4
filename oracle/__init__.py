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
        # textbook cycles / FMG / Krylov (ours)
        L.gmgo_tb_create.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_long, C.c_double]
        L.gmgo_tb_create.restype = C.c_void_p
        L.gmgo_tb_destroy.argtypes = [C.c_void_p]
        L.gmgo_tb_set_omega.argtypes = [C.c_void_p, C.c_double]
        L.gmgo_tb_iteration.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int]
        L.gmgo_tb_fmg.argtypes = [C.c_void_p, _dp, _dp]
        L.gmgo_tb_krylov.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_int, _dp]
        L.gmgo_tb_krylov.restype = C.c_int

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

    def textbook(self, N, length, alpha, L, kind, cycle, bottom, nu_pre=0, nu=5, restrict_mode=2,
                 coarse_maxit=2000, coarse_tol=0.1, omega=1.0):
        """state of the V / W / F cycles, FMG and the Krylov solvers (ours, not in the reference); `bottom` = first level of
        the coarse solver (one sawtooth pass over levels bottom..L-1)"""
        t = _Textbook(self.lib, N, length, alpha, L, kind, cycle, bottom, nu_pre, nu, restrict_mode, coarse_maxit, coarse_tol)
        self.lib.gmgo_tb_set_omega(t.st, omega)
        return t

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


class _Textbook:
    def __init__(self, lib, N, length, alpha, L, kind, cycle, bottom, nu_pre, nu, restrict_mode, coarse_maxit, coarse_tol):
        self.lib, self.N = lib, N
        self.st = lib.gmgo_tb_create(N, length, alpha, L, kind, nu_pre, nu, restrict_mode, cycle, bottom, coarse_maxit, coarse_tol)
        if not self.st:
            raise ValueError("bad level count / bottom for this N")

    def iteration(self, u, b, pre_kind=RBGS, n_pre=2):
        self.lib.gmgo_tb_iteration(self.st, u, b, pre_kind, n_pre)
        return u

    def fmg(self, u, b):
        self.lib.gmgo_tb_fmg(self.st, u, b)
        return u

    def krylov(self, method, precond, b, u, tol=1e-11, maxit=200):
        hist = np.zeros(maxit + 1)
        n = self.lib.gmgo_tb_krylov(self.st, method, precond, b, u, tol, maxit, hist)
        return u, hist[:n].copy()

    def close(self):
        if self.st:
            self.lib.gmgo_tb_destroy(self.st)
            self.st = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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


# ---------------------------------------------------------------------------------------------------
# AMG
# ---------------------------------------------------------------------------------------------------
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


class Csr:
    """plain CSR triple (int64 indices) as the checkers exchange it"""

    def __init__(self, n_rows, n_cols, ptr, col, val):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.val = np.ascontiguousarray(val, dtype=np.float64)

    @property
    def nnz(self):
        return int(self.ptr[-1])

    def same_as(self, o):
        return (self.n_rows == o.n_rows and self.n_cols == o.n_cols and np.array_equal(self.ptr, o.ptr)
                and np.array_equal(self.col, o.col) and np.array_equal(self.val, o.val))

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val, self.col, self.ptr), shape=(self.n_rows, self.n_cols))


class AmgOracle:
    """Restatement of AMG/ (oracle/amg_oracle.c)."""

    def __init__(self, path=_ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.amgo_gs_sweep.argtypes = [C.c_int64, _ip, _ip, _dp, _dp, _dp]
        L.amgo_gs_sweep_masked.argtypes = [C.c_int64, _ip, _ip, _dp, _ip, _dp, _dp]
        L.amgo_residual.argtypes = [C.c_int64, _ip, _ip, _dp, _dp, _dp, C.c_void_p]
        L.amgo_residual.restype = C.c_double
        L.amgo_restrict.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _dp, _dp, _dp]
        L.amgo_prolong_add.argtypes = [C.c_int64, _ip, _ip, _dp, _dp, _dp]
        L.amgo_build.argtypes = [C.c_int64, _ip, _ip, _dp, _dp, C.c_int, C.c_void_p]
        L.amgo_build.restype = C.c_void_p
        L.amgo_free.argtypes = [C.c_void_p]
        L.amgo_levels.argtypes = [C.c_void_p]
        L.amgo_level_info.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int64)] * 4
        L.amgo_get_A.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp]
        L.amgo_get_P.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp]
        L.amgo_get_rhs.argtypes = [C.c_void_p, C.c_int, _dp]
        L.amgo_get_cf.argtypes = [C.c_void_p, C.c_int, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")]
        L.amgo_pass.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int]
        L.amgo_pass.restype = C.c_double

    def gs(self, A, b, x, sweeps=1):
        for _ in range(sweeps):
            self.lib.amgo_gs_sweep(A.n_rows, A.ptr, A.col, A.val, x, b)
        return x

    def residual(self, A, x, b):
        r = np.zeros(A.n_rows)
        nrm = self.lib.amgo_residual(A.n_rows, A.ptr, A.col, A.val, x, b, r.ctypes.data)
        return nrm, r

    def restrict(self, P, xf):
        xc = np.zeros(P.n_cols)
        self.lib.amgo_restrict(P.n_rows, P.n_cols, P.ptr, P.col, P.val, xf, xc)
        return xc

    def prolong_add(self, P, xc, xf):
        self.lib.amgo_prolong_add(P.n_rows, P.ptr, P.col, P.val, xc, xf)
        return xf

    def build(self, A, rhs, levels, starts=None):
        st = None
        if starts is not None:
            st = np.ascontiguousarray(starts, dtype=np.int64)
        h = self.lib.amgo_build(A.n_rows, A.ptr, A.col, A.val, rhs, levels, st.ctypes.data if st is not None else None)
        if not h:
            raise ValueError("bad level count")
        return AmgHierarchy(self, h)


class AmgHierarchy:
    def __init__(self, o, h):
        self.o, self.h = o, h
        self.levels = o.lib.amgo_levels(h)

    def __del__(self):
        try:
            self.o.lib.amgo_free(self.h)
        except Exception:
            pass

    def info(self, l):
        v = [C.c_int64() for _ in range(4)]
        self.o.lib.amgo_level_info(self.h, l, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)          # n, nnzA, nnzP, ncP

    def A(self, l):
        n, nnz, _, _ = self.info(l)
        ptr, col, val = np.zeros(n + 1, np.int64), np.zeros(max(nnz, 1), np.int64), np.zeros(max(nnz, 1))
        self.o.lib.amgo_get_A(self.h, l, ptr, col, val)
        return Csr(n, n, ptr, col[:nnz], val[:nnz])

    def P(self, l):
        n, _, nnz, nc = self.info(l)
        ptr, col, val = np.zeros(n + 1, np.int64), np.zeros(max(nnz, 1), np.int64), np.zeros(max(nnz, 1))
        self.o.lib.amgo_get_P(self.h, l, ptr, col, val)
        return Csr(n, nc, ptr, col[:nnz], val[:nnz])

    def rhs(self, l):
        b = np.zeros(self.info(l)[0])
        self.o.lib.amgo_get_rhs(self.h, l, b)
        return b

    def cf(self, l):
        m = np.zeros(self.info(l)[0], np.uint8)
        self.o.lib.amgo_get_cf(self.h, l, m)
        return m

    def apply(self, x, pre=10, coarse=200, post=10):
        """AMG::apply_AMG after initialization; returns the level-0 residual norm"""
        return self.o.lib.amgo_pass(self.h, x, pre, coarse, post)


class AmgReference:
    """The reference's own AMG classes (oracle/_ref/libamgref.so); one hierarchy at a time."""

    def __init__(self, path=_REF_AMG_SO):
        L = self.lib = C.CDLL(path)
        L.amgref_set_starts.argtypes = [_ip, C.c_int]
        L.amgref_assemble.argtypes = [C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.amgref_get_system.argtypes = [_ip, _ip, _dp, _dp]
        L.amgref_build.argtypes = [C.c_long, _ip, _ip, _dp, _dp, _dp, C.c_int]
        L.amgref_build.restype = C.c_int
        L.amgref_level_info.argtypes = [C.c_int] + [C.POINTER(C.c_long)] * 4
        L.amgref_get_A.argtypes = [C.c_int, _ip, _ip, _dp]
        L.amgref_get_P.argtypes = [C.c_int, _ip, _ip, _dp]
        L.amgref_get_rhs.argtypes = [C.c_int, _dp]
        L.amgref_pass.argtypes = [_dp]
        L.amgref_pass.restype = C.c_double
        L.amgref_gs.argtypes = [C.c_long, _ip, _ip, _dp, _dp, _dp, C.c_int]
        L.amgref_set_threads.argtypes = [C.c_int]
        L.amgref_set_threads(1)       # the reference's OpenMP loops race on std::map (AMG.hpp:314-331)

    def assemble(self, msh_path):
        n, nnz = C.c_long(), C.c_long()
        self.lib.amgref_assemble(msh_path.encode(), C.byref(n), C.byref(nnz))
        ptr, col, val = np.zeros(n.value + 1, np.int64), np.zeros(nnz.value, np.int64), np.zeros(nnz.value)
        rhs = np.zeros(n.value)
        self.lib.amgref_get_system(ptr, col, val, rhs)
        return Csr(n.value, n.value, ptr, col, val), rhs

    def build(self, A, rhs, levels, starts, x0=None):
        st = np.ascontiguousarray(starts, dtype=np.int64)
        self.lib.amgref_set_starts(st, st.size)
        x0 = np.zeros(A.n_rows) if x0 is None else x0
        return self.lib.amgref_build(A.n_rows, A.ptr, A.col, A.val, rhs, x0, levels)

    def info(self, l):
        v = [C.c_long() for _ in range(4)]
        self.lib.amgref_level_info(l, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def A(self, l):
        n, nnz, _, _ = self.info(l)
        ptr, col, val = np.zeros(n + 1, np.int64), np.zeros(max(nnz, 1), np.int64), np.zeros(max(nnz, 1))
        self.lib.amgref_get_A(l, ptr, col, val)
        return Csr(n, n, ptr, col[:nnz], val[:nnz])

    def P(self, l):
        n, _, nnz, nc = self.info(l)
        ptr, col, val = np.zeros(n + 1, np.int64), np.zeros(max(nnz, 1), np.int64), np.zeros(max(nnz, 1))
        self.lib.amgref_get_P(l, ptr, col, val)
        return Csr(n, nc, ptr, col[:nnz], val[:nnz])

    def rhs(self, l):
        b = np.zeros(self.info(l)[0])
        self.lib.amgref_get_rhs(l, b)
        return b

    def apply(self):
        x = np.zeros(self.info(0)[0])
        r = self.lib.amgref_pass(x)
        return x, r

    def gs(self, A, b, x, sweeps=1):
        self.lib.amgref_gs(A.n_rows, A.ptr, A.col, A.val, b, x, sweeps)
        return x


_amg = None
_ref_amg = None


def amg():
    global _amg
    if _amg is None:
        _amg = AmgOracle()
    return _amg


def ref_amg():
    global _ref_amg
    if _ref_amg is None and os.path.exists(_REF_AMG_SO):
        _ref_amg = AmgReference()
    return _ref_amg
