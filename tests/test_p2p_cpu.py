"""CPU-side checks of the peer-store transport (csrc/p2p.cuh): the pool layout every rank computes for every other rank,
and randomised interleavings of the two exchange protocols against the hazard they are designed to exclude -- a store
into a buffer the receiver has not finished reading.  The models restate the ORDER of operations a
rank enqueues on its stream (push, signal, wait, consume); the device code itself is covered by the multi-GPU tests."""
import ctypes as C
import random

import pytest

from multigrid_prj_b200 import load
from multigrid_prj_b200.gmg import partition


def layout(n, levels, n_ranks, rank, level, which):
    off, tot = C.c_size_t(), C.c_size_t()
    assert load().mgb_gmg_pool_layout(n, levels, n_ranks, rank, level, which, C.byref(off), C.byref(tot)) == 0
    return off.value, tot.value


@pytest.mark.parametrize("n,levels,ranks", [(16385, 14, 8), (16385, 14, 4), (8193, 13, 2), (2049, 11, 8), (257, 8, 1)])
def test_pool_layout_is_disjoint_aligned_and_rank_consistent(n, levels, ranks):
    absent = C.c_size_t(-1).value
    for rank in range(ranks):
        spans, total = [], None
        for level in range(levels):
            w = n
            for _ in range(level):
                w = (w + 1) // 2
            _, _, rows = partition(n, levels, ranks, rank, level)
            pitch = (w + 2 + 15) // 16 * 16
            nbytes = (rows + 2 * 44) * pitch * 8
            for which in range(6):
                off, tot = layout(n, levels, ranks, rank, level, which)
                total = tot if total is None else total
                assert tot == total
                if level > 0 and which in (0, 1, 5):
                    assert off == absent
                    continue
                assert off % 512 == 0 and off >= 16384 and off + nbytes <= tot
                spans.append((off, off + nbytes))
        spans.sort()
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
        assert total % (2 << 20) == 0 and total >= 4 << 20


def test_layout_query_rejects_bad_arguments():
    off, tot = C.c_size_t(), C.c_size_t()
    lib = load()
    assert lib.mgb_gmg_pool_layout(200, 2, 1, 0, 0, 0, C.byref(off), C.byref(tot)) != 0      # (n-1) % 2^(levels-1) != 0
    assert lib.mgb_gmg_pool_layout(257, 8, 2, 2, 0, 0, C.byref(off), C.byref(tot)) != 0      # rank out of range


# ---- protocol models ----------------------------------------------------------------------------------------------------
def run_interleaved(programs, seed):
    """programs[r] = list of steps; a step is a callable returning False while it must wait.  Random fair scheduler."""
    rng = random.Random(seed)
    pc = [0] * len(programs)
    live = [r for r in range(len(programs)) if programs[r]]
    idle = 0
    while live:
        r = rng.choice(live)
        if programs[r][pc[r]]():
            pc[r] += 1
            idle = 0
            if pc[r] == len(programs[r]):
                live.remove(r)
        else:
            idle += 1
            assert idle < 100000, "deadlock"


@pytest.mark.parametrize("seed", range(12))
def test_amg_pair_protocol_never_overwrites_an_unread_half(seed):
    """neighbour-only exchanges (amg_kernels.cuh): per directed pair a message count = flag value = parity of the staging
    half inside the receiver's region for that sender; peers of an exchange = ranks with data in EITHER direction."""
    rng = random.Random(1000 + seed)
    R, K = 5, 40
    # symmetric random peer relation per exchange (a chain plus random extra pairs, some exchanges with isolated ranks)
    masks = []
    for _ in range(K):
        pairs = {(a, a + 1) for a in range(R - 1) if rng.random() < 0.8}
        pairs |= {tuple(sorted(rng.sample(range(R), 2))) for _ in range(rng.randrange(3))}
        masks.append([{b for (x, y) in pairs for b in ((y,) if x == a else (x,) if y == a else ())} for a in range(R)])
    flag = [[0] * R for _ in range(R)]              # flag[dst][src]
    sent = [[0] * R for _ in range(R)]              # pair_push
    got = [[0] * R for _ in range(R)]               # pair_wait
    half = [[[None, None] for _ in range(R)] for _ in range(R)]      # half[dst][src][parity] = message number stored, unread
    programs = []
    for a in range(R):
        prog = []
        for k in range(K):
            peers = sorted(masks[k][a])
            if not peers:
                continue

            def push(a=a, peers=peers):
                for b in peers:
                    m = sent[a][b] + 1
                    assert half[b][a][m & 1] is None, f"rank {a} overwrites message {half[b][a][m & 1]} of its region on rank {b}"
                    half[b][a][m & 1] = m
                for b in peers:                      # the last CTA: counts and flags after all stores
                    sent[a][b] += 1
                    flag[b][a] = sent[a][b]
                return True

            def wait(a=a, peers=peers):
                if any(flag[a][b] < got[a][b] + 1 for b in peers):
                    return False
                for b in peers:
                    got[a][b] += 1
                return True

            def unpack(a=a, peers=peers):
                for b in peers:
                    m = got[a][b]
                    assert half[a][b][m & 1] == m, "unpack reads a half that does not hold the expected message"
                    half[a][b][m & 1] = None
                return True

            prog += [push, wait, unpack]
        programs.append(prog)
    run_interleaved(programs, seed)


@pytest.mark.parametrize("seed", range(12))
def test_gmg_slab_schedule_never_stores_into_a_halo_still_being_read(seed):
    """slab iteration (gmg_solver.cu: one_iteration_ca): two all-to-all exchange points per iteration, U (halo rows of u) and
    Rr (halo rows of the restricted residuals + gather), stored DIRECTLY into the peers' arrays; the pre-sweeps read the u
    halo, the upward leg reads the residual halos.  A store is legal only if the target's previous reader is done."""
    R, K = 4, 30
    flag = {ch: [[0] * R for _ in range(R)] for ch in "UR"}
    pushed = {ch: [0] * R for ch in "UR"}
    waited = {ch: [0] * R for ch in "UR"}
    programs = []
    for a in range(R):
        prog = []

        def push(ch, a=a):
            def f():
                pushed[ch][a] += 1
                for b in range(R):
                    if b != a:
                        # every store of this exchange precedes the flag; the reader of the previous version must be done
                        assert consumed[ch][b] >= pushed[ch][a] - 1, f"rank {a} stores {ch}#{pushed[ch][a]} while rank {b} still reads #{pushed[ch][a] - 1}"
                        flag[ch][b][a] = pushed[ch][a]
                return True
            return f

        def wait(ch, a=a):
            def f():
                if any(flag[ch][a][b] < waited[ch][a] + 1 for b in range(R) if b != a):
                    return False
                waited[ch][a] += 1
                return True
            return f

        def consume(ch, a=a):
            def f():
                consumed[ch][a] += 1             # the launch that reads the halo rows of exchange #consumed has finished
                return True
            return f

        prog += [push("U"), wait("U")]                       # first iteration: the leading exchange of u
        for k in range(K):
            prog += [consume("U"),                           # pre-sweeps (+ residual + restriction) read the u halo
                     push("R"), wait("R"),
                     consume("R"),                           # replicated part + upward leg read the residual halos / gathered rows
                     push("U"), wait("U")]                   # the new u for the next iteration's pre-sweeps (+ norm parts)
        programs.append(prog)
    consumed = {ch: [0] * R for ch in "UR"}
    run_interleaved(programs, seed)
