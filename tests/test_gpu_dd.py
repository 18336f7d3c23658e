"""Multi-GPU parity (needs >= 2 GPUs on the box, skipped otherwise): the slab-decomposed engine against the
single-GPU engine on the same system -- lists, forces, trajectory, USER-LE topology (scripts/dd_check.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_dd_matches_single_gpu(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world), os.path.join(ROOT, "scripts", "dd_check.py"), "120000" if world == 2 else "400000", "200"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert "DD CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_dd_melt_matches_single_gpu():
    """BASELINE configs[2] (bench/in.chain.scaled): the dense FENE melt replicated along x, one replica per GPU"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29711", os.path.join(ROOT, "scripts", "dd_check.py"), "0", "300", "melt"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert "DD CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
