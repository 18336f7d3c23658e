#!/usr/bin/env python3
"""BASELINE configs[4]: 10,000,000 beads in 20 chains of 500,000 with 100,000 extruders, 8 GPUs, dynamic load balance.
  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/config5.py [BEADS] [CHAINS] [STEPS]
Equal-width slabs first (the reference's bricks without a balance command), a run, then `fix balance`-style re-cut
(DDEngine.rebalance) and the same run again: imbalance factor and time per MD step before / after."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems
from lammps_le_b200.engine import Engine
from lammps_le_b200.engine_dd import init_process_group

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
nchains = int(sys.argv[2]) if len(sys.argv) > 2 else 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
rank, world, local, group = init_process_group()
t0 = time.time()
s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=777, barriers="random", nchains=nchains, extruder_bond=systems.EXTRUDER_FENE)
v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
e = systems.make_engine(s, device=local, velocities=v, dd=dict(rank=rank, world=world, halo=6.0, group=group, balance=False))
if rank == 0:
    print("# %d beads, %d chains, %d GPUs, set-up %.0f s" % (n, nchains, world, time.time() - t0), flush=True)
systems.relax(e, steps=300)
e.fix_langevin(1.0, 1.0, 1.0, 904297)
e.fix_extrusion(500, 1, 2, 3, 0.5, 2, 4, 12345)
e.fix_ex_load(100, 1, 1, 1.12, 2, 0.01, 684474, (1, 1), (1, 1))
e.fix_ex_unload(100, 2, 0.5, 0.05, 456456)
e.reset_timestep(0)


def timed(label):
    e.run(64)
    b0 = Engine.stats(e)["neigh_builds"]
    e.run(steps)
    st = Engine.stats(e)
    c = e.owned_counts()
    if rank == 0:
        print("%-26s imbalance %.4f  owned atoms min %d max %d  %.4f ms per MD step (%.2f G atom-steps/s), LE %.4f ms, %.2f steps per rebuild" % (
            label, c.max() / c.mean(), c.min(), c.max(), st["last_run_gpu_ms"] / steps, n / (st["last_run_gpu_ms"] / steps) / 1e6,
            st["last_run_le_ms"] / steps, steps / max(1, st["neigh_builds"] - b0)), flush=True)


timed("equal-width slabs")
t1 = time.time()
before, after = e.rebalance(thresh=1.0)
if rank == 0:
    print("rebalance: imbalance %.4f -> %.4f in %.1f s (host-side re-cut, state carried over)" % (before, after, time.time() - t1), flush=True)
timed("after the re-cut")
e.barrier()
e.close()
