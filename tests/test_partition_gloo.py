"""Host-side slab logic of the multi-GPU path, exercised with world_size 2 over gloo on the CPU:
every rank asks the library for ITS slab and the ranks agree that the slabs tile every sharded
level exactly once, on aligned boundaries, and that replicated levels are whole on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, cases, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multigrid_prj_b200.gmg import partition
    ok = True
    for n, L in cases:
        for level in range(L):
            sh, r0, rows = partition(n, L, world, rank, level)
            mine = torch.tensor([int(sh), r0, rows], dtype=torch.int64)
            allp = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(allp, mine)
            w = n
            for _ in range(level):
                w = (w + 1) // 2
            if all(int(p[0]) for p in allp):
                # contiguous, disjoint, complete
                nxt = 0
                for p in allp:
                    ok &= int(p[1]) == nxt and int(p[2]) >= 12
                    nxt = int(p[1]) + int(p[2])
                ok &= nxt == w
                # a coarse row lives where its fine row lives
                if level > 0:
                    fsh, fr0, frows = partition(n, L, world, rank, level - 1)
                    ok &= fsh and fr0 == 2 * r0
            else:
                ok &= all(int(p[0]) == 0 and int(p[1]) == 0 and int(p[2]) == w for p in allp)
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slabs_tile_the_levels(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    cases = [(16385, 14), (8193, 13), (1025, 10), (2049, 6), (3073, 11)]
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1



def test_partition_rejects_tiny_grids():
    from multigrid_prj_b200 import MgbError
    from multigrid_prj_b200.gmg import partition
    with pytest.raises(MgbError):
        partition(65, 5, 8, 0, 0)
    assert partition(65, 5, 1, 0, 2) == (False, 0, 17)
