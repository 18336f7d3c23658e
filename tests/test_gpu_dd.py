"""Multi-GPU parity: the slab-decomposed engine against the single-GPU engine on the same system -- lists, forces,
trajectory, USER-LE topology (scripts/dd_check.py; CommBrick::forward_comm/exchange/borders, src/comm_brick.cpp:452-876).
On a box with fewer GPUs than ranks the ranks share device 0 (LE_DD_SHARE_GPU=1): same kernels, same CUDA-IPC peer stores,
same flag rounds -- the slabs time-slice one GPU instead of running side by side, so nothing is skipped on a one-GPU box."""
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


def _run_dd(world, port, args):
    env = dict(os.environ)
    shared = _ngpu() < world
    if shared:
        env["LE_DD_SHARE_GPU"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "dd_check.py")] + [str(a) for a in args]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500, env=env)
    assert "DD CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return shared


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_dd_matches_single_gpu(world):
    """chromatin chain with the three USER-LE fixes: migration, ghosts, flag rounds, replicated fix logic"""
    big = _ngpu() >= world
    n = (120000 if world == 2 else 400000) if big else (40000 if world == 2 else 90000)
    _run_dd(world, 29700 + world, [n, 200])


@pytest.mark.gpu
def test_dd_melt_matches_single_gpu():
    """BASELINE configs[2] (bench/in.chain.scaled): the dense FENE melt replicated along x, one replica per GPU"""
    _run_dd(2, 29711, [0, 300, "melt"])


@pytest.mark.gpu
def test_dd_bond_create_and_break_match_single_gpu():
    """fix bond/create + fix bond/break (src/MC, the ancestors of ex_load / ex_unload) on two slabs: the owner of an atom measures its
    closest eligible neighbor (ghosts included), every GPU decides from the same records -- topology, types and counters equal the
    single-GPU run"""
    _run_dd(2, 29715, [40000, 100, "mc"])


@pytest.mark.gpu
def test_dynamic_rebalance_keeps_the_trajectory():
    """`fix balance` (src/fix_balance.cpp:191-270): equal-width slabs over a chain that crowds one part of the box, re-cut by atom
    count in the middle of the run; trajectory and USER-LE topology equal the single-GPU run (scripts/dd_balance_check.py)"""
    env = dict(os.environ)
    if _ngpu() < 2:
        env["LE_DD_SHARE_GPU"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29721", os.path.join(ROOT, "scripts", "dd_balance_check.py"), "40000", "150"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500, env=env)
    assert "DD BALANCE CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
