"""Host-side logic of the row-block sharded AMG path, on the CPU.

The library exports the two pure-host pieces of `mgb_amg_create_sharded`: the block partition and the
ghost-exchange plan of an operator.  These tests drive them exactly as the device path does -- every rank
keeps a full-length (globally indexed) vector, updates only its own rows, and refreshes the ghost entries
listed in the plan -- with numpy standing in for the kernels, and require the assembled result to equal the
single-rank computation bit for bit.  The last test runs the same protocol over real messages (gloo,
world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _system(side=31):
    from amg_bench import synthetic_system
    A, rhs = synthetic_system(side)
    return A.tocsr(), rhs


def _greedy_colours(A):
    n = A.shape[0]
    colour = -np.ones(n, np.int32)
    for i in range(n):
        nb = A.indices[A.indptr[i]:A.indptr[i + 1]]
        used = set(colour[nb[nb != i]].tolist())
        c = 0
        while c in used:
            c += 1
        colour[i] = c
    return colour, int(colour.max()) + 1


def _gs_rows(A, diag, x, b, rows):
    """multicolour Gauss-Seidel update of `rows` (mutually independent) reading x"""
    for i in rows:
        s = 0.0
        for k in range(A.indptr[i], A.indptr[i + 1]):
            j = A.indices[k]
            if j != i:
                s += A.data[k] * x[j]
        x[i] = (b[i] - s) / diag[i]


def _plans(M, world, group=None, ng=1):
    from multigrid_prj_b200.amg import halo_plan
    return [halo_plan(M.indptr, M.indices, M.data, M.shape, world, r, group, ng) for r in range(world)]


def _exchange(plans, xs, world, g0, g1):
    """what pack -> ncclSend/ncclRecv -> unpack does, for the groups [g0, g1)"""
    for me in range(world):
        sp_, si, _, _ = plans[me]
        for g in range(g0, g1):
            for p in range(world):
                q = g * world + p
                seg = si[sp_[q]:sp_[q + 1]]
                if seg.size == 0:
                    continue
                _, _, rp, ri = plans[p]
                qq = g * world + me
                dst = ri[rp[qq]:rp[qq + 1]]
                assert np.array_equal(seg, dst), "send list of one rank != receive list of its peer"
                xs[p][dst] = xs[me][seg]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partition_tiles(world):
    from multigrid_prj_b200.amg import partition
    for n in (1, 7, 64, 1000, 6241, 15992001):
        nxt = 0
        for r in range(world):
            r0, rows = partition(n, world, r)
            assert r0 == nxt
            nxt = r0 + rows
        assert nxt == n
        sizes = [partition(n, world, r)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [2, 4])
def test_plan_lists_are_consistent_and_minimal(world):
    from multigrid_prj_b200.amg import partition
    A, _ = _system(25)
    n = A.shape[0]
    plans = _plans(A, world)
    for me in range(world):
        r0, rows = partition(n, world, me)
        sp_, si, rp, ri = plans[me]
        # receive list = exactly the off-block columns my rows reference
        cols = np.unique(A.indices[A.indptr[r0]:A.indptr[r0 + rows]])
        ghosts = cols[(cols < r0) | (cols >= r0 + rows)]
        assert np.array_equal(np.sort(ri), ghosts)
        # what I send lies in my block
        assert np.all((si >= r0) & (si < r0 + rows))
        for p in range(world):
            assert np.array_equal(si[sp_[p]:sp_[p + 1]], plans[p][3][plans[p][2][me]:plans[p][2][me + 1]])
        assert sp_[me + 1] == sp_[me] and rp[me + 1] == rp[me]          # nothing to myself


@pytest.mark.parametrize("world,hybrid", [(2, False), (3, False), (4, False), (2, True)])
def test_sharded_multicolour_gs_equals_single_rank(world, hybrid):
    from multigrid_prj_b200.amg import partition
    A, b = _system(21)
    n = A.shape[0]
    diag = A.diagonal()
    colour, nc = _greedy_colours(A)
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal(n)
    # single rank: colours in order, rows ascending
    ref = x0.copy()
    for _ in range(2):
        for c in range(nc):
            _gs_rows(A, diag, ref, b, np.nonzero(colour == c)[0])
    # sharded: full exchange first (set_vector gives valid ghosts, but start from stale ghosts to exercise it)
    full, coloured = _plans(A, world), _plans(A, world, colour, nc)
    blocks = [partition(n, world, r) for r in range(world)]
    xs = []
    for r in range(world):
        x = np.full(n, np.nan)
        x[blocks[r][0]:blocks[r][0] + blocks[r][1]] = x0[blocks[r][0]:blocks[r][0] + blocks[r][1]]
        xs.append(x)
    _exchange(full, xs, world, 0, 1)
    for _ in range(2):
        for c in range(nc):
            for r in range(world):
                r0, rows = blocks[r]
                own = np.arange(r0, r0 + rows)
                _gs_rows(A, diag, xs[r], b, own[colour[own] == c])
            if not hybrid:
                _exchange(coloured, xs, world, c, c + 1)
        if hybrid:
            _exchange(full, xs, world, 0, 1)
    got = np.concatenate([xs[r][blocks[r][0]:blocks[r][0] + blocks[r][1]] for r in range(world)])
    if hybrid:      # block-Jacobi coupling across the cuts: a different (still convergent) iterate
        assert not np.array_equal(got, ref)
        assert np.linalg.norm(b - A @ got) < np.linalg.norm(b - A @ x0)
    else:
        assert np.array_equal(got, ref)
        # and every ghost a rank holds is the owner's current value
        for r in range(world):
            ri = full[r][3]
            assert np.array_equal(xs[r][ri], ref[ri])


def _transfer_operators():
    """P of the first coarsening of a small system, from the library's host setup stages"""
    import ctypes as C
    from multigrid_prj_b200 import load
    from multigrid_prj_b200._lib import check
    lib = load()
    A, _ = _system(21)
    n = A.shape[0]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ptr, col, val = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)
    hA, hP = C.c_void_p(), C.c_void_p()
    check(lib.mgb_csr_create(n, n, p(ptr), p(col), p(val), C.byref(hA)))
    mask = np.zeros(n, np.uint8)
    ncoarse = C.c_size_t()
    check(lib.mgb_amg_select_coarse_nodes(hA, 0.2, -1, p(mask), C.byref(ncoarse)))
    check(lib.mgb_amg_build_prolongation(hA, 0.2, p(mask), C.byref(hP)))
    nr, ncol, nnz = C.c_size_t(), C.c_size_t(), C.c_size_t()
    check(lib.mgb_csr_info(hP, C.byref(nr), C.byref(ncol), C.byref(nnz)))
    pp, pc, pv = np.zeros(n + 1, np.int64), np.zeros(nnz.value, np.int64), np.zeros(nnz.value)
    check(lib.mgb_csr_get(hP, p(pp), p(pc), p(pv)))
    lib.mgb_csr_destroy(hA); lib.mgb_csr_destroy(hP)
    P = sp.csr_matrix((pv, pc, pp), shape=(n, ncol.value))
    return A, P


@pytest.mark.parametrize("world,coarse_replicated", [(2, False), (4, False), (2, True)])
def test_sharded_transfers_equal_single_rank(world, coarse_replicated):
    from multigrid_prj_b200.amg import partition
    A, P = _transfer_operators()
    n, nc = P.shape
    R = P.T.tocsr(); R.sort_indices()
    rng = np.random.default_rng(9)
    xf, xc = rng.standard_normal(n), rng.standard_normal(nc)
    fb = [partition(n, world, r) for r in range(world)]
    cb = [partition(nc, world, r) for r in range(world)]
    # restriction x_c = R x_f: rows of R in coarse blocks, ghosts of the FINE vector
    planR = _plans(R, world)
    xs = []
    for r in range(world):
        v = np.full(n, np.nan); v[fb[r][0]:fb[r][0] + fb[r][1]] = xf[fb[r][0]:fb[r][0] + fb[r][1]]; xs.append(v)
    _exchange(planR, xs, world, 0, 1)
    parts = [R[cb[r][0]:cb[r][0] + cb[r][1]] @ np.nan_to_num(xs[r], nan=1e300) for r in range(world)]
    assert np.allclose(np.concatenate(parts), R @ xf, rtol=1e-14, atol=1e-14)
    # prolongation x_f += P x_c: rows of P in fine blocks, ghosts of the COARSE vector (none when it is replicated)
    want = xf + P @ xc
    if coarse_replicated:
        got = np.concatenate([xf[fb[r][0]:fb[r][0] + fb[r][1]] + P[fb[r][0]:fb[r][0] + fb[r][1]] @ xc for r in range(world)])
    else:
        planP = _plans(P.tocsr(), world)
        cs = []
        for r in range(world):
            v = np.full(nc, np.nan); v[cb[r][0]:cb[r][0] + cb[r][1]] = xc[cb[r][0]:cb[r][0] + cb[r][1]]; cs.append(v)
        _exchange(planP, cs, world, 0, 1)
        got = np.concatenate([xf[fb[r][0]:fb[r][0] + fb[r][1]] + P[fb[r][0]:fb[r][0] + fb[r][1]] @ np.nan_to_num(cs[r], nan=1e300)
                              for r in range(world)])
    assert np.allclose(got, want, rtol=1e-14, atol=1e-14)


# ---- the same protocol over real messages: gloo, world_size 2 ---------------------------------------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
    from multigrid_prj_b200.amg import halo_plan, partition
    A, b = _system(21)
    n = A.shape[0]
    diag = A.diagonal()
    colour, nc = _greedy_colours(A)
    x0 = np.random.default_rng(5).standard_normal(n)
    ref = x0.copy()
    for c in range(nc):
        _gs_rows(A, diag, ref, b, np.nonzero(colour == c)[0])
    sp_, si, rp, ri = halo_plan(A.indptr, A.indices, A.data, A.shape, world, rank, colour, nc)
    r0, rows = partition(n, world, rank)
    x = np.full(n, np.nan); x[r0:r0 + rows] = x0[r0:r0 + rows]

    def exchange(g0, g1):
        ops, bufs = [], []
        for g in range(g0, g1):
            for p in range(world):
                qi = g * world + p
                if sp_[qi + 1] > sp_[qi]:
                    ops.append(dist.P2POp(dist.isend, torch.from_numpy(x[si[sp_[qi]:sp_[qi + 1]]].copy()), p))
                if rp[qi + 1] > rp[qi]:
                    t = torch.zeros(int(rp[qi + 1] - rp[qi]), dtype=torch.float64)
                    bufs.append((t, ri[rp[qi]:rp[qi + 1]]))
                    ops.append(dist.P2POp(dist.irecv, t, p))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for t, idx in bufs:
            x[idx] = t.numpy()

    exchange(0, nc)
    own = np.arange(r0, r0 + rows)
    for c in range(nc):
        _gs_rows(A, diag, x, b, own[colour[own] == c])
        exchange(c, c + 1)
    ok = np.array_equal(x[r0:r0 + rows], ref[r0:r0 + rows]) and np.array_equal(x[ri], ref[ri])
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(t.item()))
    dist.destroy_process_group()


def test_sharded_sweep_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1
