"""Row-block sharded AMG on >= 2 GPUs (skipped on a single-GPU box): tools/amg_check.py under torchrun.
The host-side half (partition, ghost-exchange plans, the exchange protocol over gloo) is covered on the CPU by
tests/test_amg_shard_plan.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from multigrid_prj_b200 import load
    return load().mgb_device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_results_equal_single_rank(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29520 + world), os.path.join(ROOT, "tools", "amg_check.py"),
           "--side", "201", "--levels", "4", "--min-rows", "1000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "AMG_CHECK OK" in p.stdout, p.stdout[-4000:] + p.stderr[-3000:]
