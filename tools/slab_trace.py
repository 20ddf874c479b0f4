"""Phase timeline of the slab iteration (CUDA events between phases, rank 0 prints): run under torchrun.

    MGB_TRACE=1 python -m torch.distributed.run --nproc-per-node N tools/slab_trace.py [--n 16385] [--levels 14]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                            # noqa: E402
import torch.distributed as dist                        # noqa: E402
from multigrid_prj_b200 import Gmg, GmgConfig           # noqa: E402
from multigrid_prj_b200 import gmg as G                 # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16385)
ap.add_argument("--levels", type=int, default=14)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
cfg = GmgConfig.fast(a.n, a.levels, device=local)
cfg.rank, cfg.n_ranks = rank, world
ids = [G.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
cfg.nccl_id = ids[0]
with Gmg(cfg) as g:
    g.set_rhs_test(1); g.set_u(None)
    for _ in range(4):
        g.run_cycles(3)          # fewer than 4 cycles per call: never captured in a graph, so the events are recorded
    if rank == 0:
        print("levels:", [(l, g.rows(l)) for l in range(a.levels)])
dist.destroy_process_group()
