"""Thin Python handle over the AMG part of the C ABI (tests and bench only).

Names follow the reference (AMG/include/AMG.hpp): `Amg(A, rhs, levels)` is the constructor followed
by `initialization()`, `apply()` is the body of `apply_AMG()`, `smooth/restrict/prolong/residual` are
`apply_smoother_operator / apply_restriction_operator / apply_prolungation_operator / compute_residual`.
"""
import ctypes as C

import numpy as np

from ._lib import AmgConfigStruct, GmgStatsStruct, check, load

GS_LEX, JACOBI, GS_MULTICOLOUR, L1_JACOBI = 0, 1, 3, 4


class System:
    """A linear system assembled ON THE DEVICE (mgb_fem_assemble_p1 / mgb_fem_synthetic): CSR + right-hand side in HBM."""

    def __init__(self, handle):
        self.lib = load()
        self.h = handle

    @staticmethod
    def assemble_p1(x, y, on_boundary, tri, exact_order=True, device=0):
        """P1 Poisson assembly of the reference (AMG/src/main.cpp:34-117) on a triangle mesh; tri: (T, 3) ascending vertex ids"""
        lib = load()
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        b = np.ascontiguousarray(on_boundary, dtype=np.uint8); t = np.ascontiguousarray(tri, dtype=np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        h = C.c_void_p()
        check(lib.mgb_fem_assemble_p1(x.size, p(x), p(y), p(b), t.shape[0], p(t), int(exact_order), device, C.byref(h)))
        return System(h)

    @staticmethod
    def synthetic(side, seed=12345, device=0):
        """BASELINE config 5: jittered side x side lattice with hashed diagonals, generated and assembled on the device"""
        h = C.c_void_p()
        check(load().mgb_fem_synthetic(side, seed, device, C.byref(h)))
        return System(h)

    @staticmethod
    def synthetic_mesh(side, seed=12345):
        """the same mesh on the host: (x, y, on_boundary, tri)"""
        x, y = np.zeros(side * side), np.zeros(side * side)
        b = np.zeros(side * side, np.uint8); t = np.zeros((2 * (side - 1) ** 2, 3), np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(load().mgb_fem_synthetic_mesh(side, seed, p(x), p(y), p(b), p(t)))
        return x, y, b, t

    def info(self):
        n, nnz = C.c_size_t(), C.c_size_t()
        check(self.lib.mgb_system_info(self.h, C.byref(n), C.byref(nnz)))
        return n.value, nnz.value

    def get(self):
        """(ptr, col, val, rhs) on the host"""
        n, nnz = self.info()
        ptr, col, val, rhs = np.zeros(n + 1, np.int64), np.zeros(max(nnz, 1), np.int64), np.zeros(max(nnz, 1)), np.zeros(max(n, 1))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(self.lib.mgb_system_get(self.h, p(ptr), p(col), p(val), p(rhs)))
        return ptr, col[:nnz], val[:nnz], rhs[:n]

    def close(self):
        if self.h:
            self.lib.mgb_system_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Amg:
    def __init__(self, ptr, col, val, rhs, levels=5, fast=False, starts=None, rank=0, n_ranks=1, nccl_id=None, system=None,
                 device_path=False, **kw):
        """rank / n_ranks / nccl_id: row-block sharded over one box (mgb_amg_create_sharded); every rank passes the
        same system.  device_path: mgb_amg_config_device (hierarchy built on the device, l1-Jacobi below level 0).
        system: a device-resident System instead of host CSR arrays (mgb_amg_create_from_system).
        Other keywords set fields of mgb_amg_config (hybrid_gs, shard_min_rows, jacobi_omega, coarse_smoother, ...)."""
        self.lib = load()
        c = AmgConfigStruct()
        (self.lib.mgb_amg_config_device if (device_path or system is not None) else
         (self.lib.mgb_amg_config_fast if fast else self.lib.mgb_amg_config_default))(C.byref(c))
        c.levels = levels
        for k, v in kw.items():
            setattr(c, k, v)
        if starts is not None:
            for i, s in enumerate(starts):
                c.start_index[i] = int(s)
        self.h = C.c_void_p()
        self.rank, self.n_ranks = rank, n_ranks
        if system is not None:
            idb = C.cast((C.c_ubyte * 128)(*bytes(nccl_id)), C.c_void_p) if n_ranks > 1 else None
            check(self.lib.mgb_amg_create_from_system(C.byref(c), system.h, rank, n_ranks, idb, C.byref(self.h)))
            self.n = system.info()[0]
            self.levels = self.lib.mgb_amg_n_levels(self.h)
            return
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int64)
        val = np.ascontiguousarray(val, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        self.n = ptr.size - 1
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if n_ranks > 1:
            idb = (C.c_ubyte * 128)(*bytes(nccl_id))
            check(self.lib.mgb_amg_create_sharded(C.byref(c), self.n, p(ptr), p(col), p(val), p(rhs), rank, n_ranks,
                                                  C.cast(idb, C.c_void_p), C.byref(self.h)))
        else:
            check(self.lib.mgb_amg_create_from_csr(C.byref(c), self.n, p(ptr), p(col), p(val), p(rhs), C.byref(self.h)))
        self.levels = self.lib.mgb_amg_n_levels(self.h)

    @staticmethod
    def from_system(system, levels=10, rank=0, n_ranks=1, nccl_id=None, **kw):
        return Amg(None, None, None, None, levels=levels, rank=rank, n_ranks=n_ranks, nccl_id=nccl_id, system=system, **kw)

    def close(self):
        if self.h:
            self.lib.mgb_amg_destroy(self.h)
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

    def info(self, level):
        v = [C.c_size_t() for _ in range(4)]
        w, k = C.c_int(), C.c_int()
        check(self.lib.mgb_amg_level_info(self.h, level, *[C.byref(x) for x in v], C.byref(w), C.byref(k)))
        return {"n": v[0].value, "nnz_a": v[1].value, "nnz_p": v[2].value, "n_coarse": v[3].value,
                "wavefronts": w.value, "colours": k.value}

    def rows(self, level):
        """(row0, rows, sharded): the rows of `level` this rank works on"""
        r0, rows, sh = C.c_size_t(), C.c_size_t(), C.c_int()
        check(self.lib.mgb_amg_level_rows(self.h, level, C.byref(r0), C.byref(rows), C.byref(sh)))
        return r0.value, rows.value, bool(sh.value)

    def matrix(self, level, which):
        """(ptr, col, val) of A_level (which=0) or P_level (which=1)"""
        i = self.info(level)
        nnz = i["nnz_a"] if which == 0 else i["nnz_p"]
        ptr, col, val = np.zeros(i["n"] + 1, np.int64), np.zeros(max(nnz, 1), np.int64), np.zeros(max(nnz, 1))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(self.lib.mgb_amg_get_matrix(self.h, level, which, p(ptr), p(col), p(val)))
        return ptr, col[:nnz], val[:nnz]

    def schedule(self, level, which):
        """wavefront (which=0) or colour (which=1) of every row"""
        out = np.zeros(max(self.info(level)["n"], 1), np.int32)
        check(self.lib.mgb_amg_get_schedule(self.h, level, which, out.ctypes.data_as(C.c_void_p)))
        return out[:self.info(level)["n"]]

    def vector(self, level, which=0):
        out = np.zeros(max(self.info(level)["n"], 1))
        check(self.lib.mgb_amg_get_vector(self.h, level, which, out.ctypes.data_as(C.c_void_p)))
        return out[:self.info(level)["n"]]

    def set_vector(self, level, which, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        check(self.lib.mgb_amg_set_vector(self.h, level, which, a.ctypes.data_as(C.c_void_p)))

    def smooth(self, level, kind, sweeps):
        check(self.lib.mgb_amg_smooth(self.h, level, kind, sweeps))

    def restrict(self, level):
        check(self.lib.mgb_amg_restrict(self.h, level))

    def prolong(self, level):
        check(self.lib.mgb_amg_prolong(self.h, level))

    def residual(self, level=0):
        r = C.c_double()
        check(self.lib.mgb_amg_residual(self.h, level, C.byref(r)))
        return r.value

    def apply(self, want_residual=True):
        r = C.c_double()
        check(self.lib.mgb_amg_apply(self.h, C.byref(r) if want_residual else None))
        return r.value

    def solve(self, tol=1e-8, maxit=50, nu1=2, nu2=2, coarse=20):
        """correction-scheme V-cycles (not in the reference); returns the residual history"""
        hist = np.zeros(maxit + 1)
        n = C.c_int()
        check(self.lib.mgb_amg_solve(self.h, tol, maxit, nu1, nu2, coarse, hist.ctypes.data_as(C.c_void_p), C.byref(n)))
        return hist[:n.value].copy()

    def checksum(self, level=0, which=0):
        v = C.c_uint64()
        check(self.lib.mgb_amg_checksum(self.h, level, which, C.byref(v)))
        return v.value

    def sync(self):
        check(self.lib.mgb_amg_sync(self.h))

    def stream(self):
        return self.lib.mgb_amg_stream(self.h)

    def stats(self):
        s = GmgStatsStruct()
        check(self.lib.mgb_amg_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_ if k != "reserved"}

    def reset_stats(self):
        check(self.lib.mgb_amg_reset_stats(self.h))


def partition(n, n_ranks, rank):
    """(row0, rows) of the block of an n-long index space that `rank` owns (mgb_amg_partition; host only)"""
    r0, rows = C.c_size_t(), C.c_size_t()
    check(load().mgb_amg_partition(n, n_ranks, rank, C.byref(r0), C.byref(rows)))
    return r0.value, rows.value


def halo_plan(ptr, col, val, shape, n_ranks, rank, group_of_col=None, n_groups=1):
    """ghost-exchange plan of a CSR operator for `rank` (mgb_amg_halo_plan; host only).
    Returns (send_ptr, send_idx, recv_ptr, recv_idx); segment (g, p) = ptr[g * n_ranks + p : g * n_ranks + p + 2]."""
    lib = load()
    ptr = np.ascontiguousarray(ptr, dtype=np.int64); col = np.ascontiguousarray(col, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    m = C.c_void_p()
    check(lib.mgb_csr_create(shape[0], shape[1], p(ptr), p(col), p(val), C.byref(m)))
    try:
        g = None if group_of_col is None else np.ascontiguousarray(group_of_col, dtype=np.int32)
        ng = 1 if g is None else n_groups
        sp, rp = np.zeros(ng * n_ranks + 1, np.int64), np.zeros(ng * n_ranks + 1, np.int64)
        gp = None if g is None else p(g)
        check(lib.mgb_amg_halo_plan(m, n_ranks, rank, gp, ng, p(sp), None, p(rp), None))
        si, ri = np.zeros(max(int(sp[-1]), 1), np.int64), np.zeros(max(int(rp[-1]), 1), np.int64)
        check(lib.mgb_amg_halo_plan(m, n_ranks, rank, gp, ng, p(sp), p(si), p(rp), p(ri)))
        return sp, si[:int(sp[-1])], rp, ri[:int(rp[-1])]
    finally:
        lib.mgb_csr_destroy(m)
