"""Slab-decomposed GMG on >= 2 GPUs (skipped on a single-GPU box): tools/mg_check.py under torchrun."""
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
def test_slab_results_equal_single_rank(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tools", "mg_check.py"),
           "--size", "2049", "--depth", "11"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MG_CHECK OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
