"""Thin Python handle over the GMG part of the C ABI (tests and bench only).

Names follow the reference's driver (GeometricMultigrid/src/main.cpp): `smooth` is
`u * GS`, `residual` is `u * RES; RES.Norm()`, `cycle` is `u * MG`, `solve` is the loop of
main.cpp:73-116.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import GmgConfigStruct, GmgStatsStruct, check, load

GS_LEX, JACOBI, BICGSTAB, GS_RB = 0, 1, 2, 3
INJECTION, HALF_INJECTION, FULL_WEIGHTING = 0, 1, 2
VEC_U, VEC_F, VEC_E, VEC_R = 0, 1, 2, 3
CYCLE_SAWTOOTH, CYCLE_V, CYCLE_W, CYCLE_F = 0, 1, 2, 3
KRYLOV_CG, KRYLOV_BICGSTAB = 0, 1
PRECOND_NONE, PRECOND_MG = 0, 1


@dataclass
class GmgConfig:
    n: int
    levels: int
    length: float = 10.0
    alpha: float = 1.0
    smoother: int = GS_LEX
    pre_smoother: int = GS_LEX
    n_pre: int = 2
    nu: int = 5
    restriction: int = INJECTION
    coarse_tol: float = 0.1
    coarse_maxit: int = 2000
    device: int = 0
    rank: int = 0
    n_ranks: int = 1
    nccl_id: bytes = b""
    tail_max_width: int = 65
    use_graph: int = 1
    rb_fast_arith: int = 0
    rb_fused: int = 1
    fuse_correction: int = 0
    fuse_residual: int = 0
    fuse_prolong: int = 0
    defer_norm: int = 0
    jacobi_omega: float = 1.0
    cycle_type: int = CYCLE_SAWTOOTH
    nu_pre: int = 0
    fmg: int = 0

    @staticmethod
    def fast(n, levels, **kw):
        """the B200 fast path: red-black GS everywhere + full weighting"""
        kw.setdefault("smoother", GS_RB)
        kw.setdefault("pre_smoother", GS_RB)
        kw.setdefault("restriction", FULL_WEIGHTING)
        kw.setdefault("rb_fast_arith", 1)
        kw.setdefault("fuse_correction", 1)
        kw.setdefault("fuse_residual", 1)
        kw.setdefault("fuse_prolong", 1)
        return GmgConfig(n=n, levels=levels, **kw)


def partition(n, levels, n_ranks, rank, level):
    """(sharded, row0, rows) of `rank` on `level` -- host arithmetic only, no GPU needed"""
    lib = load()
    sh, r0, r = C.c_int(), C.c_size_t(), C.c_size_t()
    check(lib.mgb_gmg_partition(n, levels, n_ranks, rank, level, C.byref(sh), C.byref(r0), C.byref(r)))
    return bool(sh.value), r0.value, r.value


def nccl_unique_id():
    lib = load()
    buf = (C.c_ubyte * 128)()
    check(lib.mgb_nccl_unique_id(C.byref(buf)))
    return bytes(buf)


class Gmg:
    def __init__(self, cfg: GmgConfig):
        self.lib = load()
        c = GmgConfigStruct()
        self.lib.mgb_gmg_config_default(C.byref(c))
        for k in ("n", "levels", "length", "alpha", "smoother", "pre_smoother", "n_pre", "nu",
                  "restriction", "coarse_tol", "coarse_maxit", "device", "rank", "n_ranks",
                  "tail_max_width", "use_graph", "rb_fast_arith", "rb_fused", "fuse_correction", "fuse_residual", "fuse_prolong", "defer_norm",
                  "jacobi_omega", "cycle_type", "nu_pre", "fmg"):
            setattr(c, k, getattr(cfg, k))
        if cfg.nccl_id:
            C.memmove(c.nccl_id, cfg.nccl_id, min(128, len(cfg.nccl_id)))
        self.cfg = cfg
        self.h = C.c_void_p()
        check(self.lib.mgb_gmg_create(C.byref(c), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.mgb_gmg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # geometry
    def width(self, level):
        w = C.c_size_t()
        check(self.lib.mgb_gmg_level_width(self.h, level, C.byref(w)))
        return w.value

    def rows(self, level):
        r0, r = C.c_size_t(), C.c_size_t()
        check(self.lib.mgb_gmg_level_rows(self.h, level, C.byref(r0), C.byref(r)))
        return r0.value, r.value

    # data movement (host arrays are the GLOBAL w x w grids; each rank moves its slab)
    @staticmethod
    def _ptr(a):
        assert a.dtype == np.float64 and a.flags.c_contiguous
        return a.ctypes.data_as(C.c_void_p)

    def set_rhs(self, b):
        check(self.lib.mgb_gmg_set_rhs(self.h, self._ptr(b)))

    def set_rhs_test(self, test):
        check(self.lib.mgb_gmg_set_rhs_test(self.h, test))

    def set_u(self, u=None):
        check(self.lib.mgb_gmg_set_u(self.h, self._ptr(u) if u is not None else None))

    def get_u(self, out=None):
        n = self.cfg.n
        out = np.zeros((n, n)) if out is None else out
        check(self.lib.mgb_gmg_get_u(self.h, self._ptr(out)))
        return out

    def set_level(self, level, which, a):
        w = self.width(level)
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(w, w)
        check(self.lib.mgb_gmg_set_level(self.h, level, which, self._ptr(a)))

    def get_level(self, level, which):
        w = self.width(level)
        out = np.zeros((w, w))
        check(self.lib.mgb_gmg_get_level(self.h, level, which, self._ptr(out)))
        return out

    # operators
    def smooth(self, level, kind, sweeps=1, sol=VEC_E, rhs=VEC_R):
        check(self.lib.mgb_gmg_smooth(self.h, level, kind, sweeps, sol, rhs))

    def residual(self, level, sol=VEC_E, rhs=VEC_R, store=False):
        ss = C.c_double()
        check(self.lib.mgb_gmg_residual(self.h, level, sol, rhs, int(store), C.byref(ss)))
        return ss.value

    def sumsq(self, level, which):
        ss = C.c_double()
        check(self.lib.mgb_gmg_sumsq(self.h, level, which, C.byref(ss)))
        return ss.value

    def restrict(self):
        check(self.lib.mgb_gmg_restrict(self.h))

    def prolong(self, level_coarse):
        check(self.lib.mgb_gmg_prolong(self.h, level_coarse))

    def cycle(self):
        rel, its = C.c_double(), C.c_int()
        check(self.lib.mgb_gmg_cycle(self.h, C.byref(rel), C.byref(its)))
        return rel.value, its.value

    def fine_leg(self, want_norm=False):
        ss = C.c_double()
        check(self.lib.mgb_gmg_fine_leg(self.h, C.byref(ss) if want_norm else None))
        return ss.value

    def solve(self, tol=1e-11, maxiter=1000, check_every=1):
        hist = np.zeros(maxiter + 1)
        n = C.c_int()
        check(self.lib.mgb_gmg_solve(self.h, tol, maxiter, check_every, self._ptr(hist), C.byref(n)))
        return hist[:n.value].copy()

    def set_cycle_type(self, cycle_type, nu_pre=0, fmg=0):
        check(self.lib.mgb_gmg_set_cycle_type(self.h, cycle_type, nu_pre, fmg))

    def fmg(self):
        """one full-multigrid pass on the residual equation, applied to u"""
        check(self.lib.mgb_gmg_fmg(self.h))

    def krylov(self, method=KRYLOV_CG, precond=PRECOND_MG, tol=1e-11, maxit=200):
        hist = np.zeros(maxit + 1)
        n = C.c_int()
        check(self.lib.mgb_gmg_krylov(self.h, method, precond, tol, maxit, self._ptr(hist), C.byref(n)))
        return hist[:n.value].copy()

    def iterate(self, confirm_below=0.0):
        """one driver iteration as one library call: (sum of squares of the new residual, coarse relative residual)"""
        ss, cr = C.c_double(), C.c_double()
        check(self.lib.mgb_gmg_iterate(self.h, confirm_below, C.byref(ss), C.byref(cr)))
        return ss.value, cr.value

    def run_cycles(self, cycles, want_relres=True):
        rel = C.c_double()
        check(self.lib.mgb_gmg_run_cycles(self.h, cycles, C.byref(rel) if want_relres else None))
        return rel.value

    def set_stream_impl(self, impl):
        """1 = first-generation streaming kernel everywhere, 2 = bulk-copy fed kernel where instantiated (default)"""
        check(self.lib.mgb_gmg_set_stream_impl(self.h, impl))

    def checksum(self, level=0, which=VEC_U):
        """64-bit checksum over all ranks (collective when n_ranks > 1): equal <=> bit-identical vectors"""
        v = C.c_uint64()
        check(self.lib.mgb_gmg_checksum(self.h, level, which, C.byref(v)))
        return v.value

    def sync(self):
        check(self.lib.mgb_gmg_sync(self.h))

    def stats(self):
        s = GmgStatsStruct()
        check(self.lib.mgb_gmg_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_ if k != "reserved"}

    def reset_stats(self):
        check(self.lib.mgb_gmg_reset_stats(self.h))

    def stream(self):
        return self.lib.mgb_gmg_stream(self.h)


class Timer:
    """CUDA-event pair recorded on the handle's own stream (torch.cuda.Event would only see torch's)."""

    def __init__(self):
        self.lib = load()
        self.t = C.c_void_p()
        check(self.lib.mgb_timer_create(C.byref(self.t)))

    def start(self, stream):
        check(self.lib.mgb_timer_start(self.t, stream))

    def stop(self, stream):
        check(self.lib.mgb_timer_stop(self.t, stream))

    def elapsed_ms(self):
        ms = C.c_double()
        check(self.lib.mgb_timer_elapsed_ms(self.t, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            self.lib.mgb_timer_destroy(self.t)
        except Exception:
            pass
