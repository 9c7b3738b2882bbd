"""Bipartite-sharded step vs the single-GPU engine.  World 1 runs in-process on any box; worlds 2 (and 4, 8 when
the box has the GPUs) are launched with torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "tests", "sharded_gpu_check.py")


def test_sharded_engine_world1_matches_single_gpu(built_lib):
    r = subprocess.run([sys.executable, CHECK], capture_output=True, text=True, timeout=600)
    assert "SHARDED_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_step_matches_single_gpu(built_lib, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs at least {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + world), CHECK]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "SHARDED_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
