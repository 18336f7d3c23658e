#!/usr/bin/env python3
"""MD-only timing of the slab-decomposed engine: torchrun ... scripts/dd_perf.py BEADS_PER_GPU STEPS HALO [le]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems
from lammps_le_b200.engine import Engine
from lammps_le_b200.engine_dd import init_process_group
rank, world, local, group = init_process_group()
nper, steps, halo = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
with_le = len(sys.argv) > 4 and sys.argv[4] == "le"
n = nper * world
s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=12345, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
e = systems.make_engine(s, device=local, velocities=v, dd=dict(rank=rank, world=world, halo=halo, group=group) if world > 1 else None)
systems.relax(e, steps=300)
e.fix_langevin(1.0, 1.0, 1.0, 904297)
if with_le:
    e.fix_extrusion(500, 1, 2, 3, 0.5, 2, 4, 12345)
    e.fix_ex_load(100, 1, 1, 1.12, 2, 0.01, 684474, (1, 1), (1, 1))
    e.fix_ex_unload(100, 2, 0.5, 0.05, 456456)
e.reset_timestep(0)
e.run(steps)
b0 = Engine.stats(e)["neigh_builds"]
e.run(steps)
st = Engine.stats(e)
if rank == 0:
    print("world %d beads/gpu %d halo %.1f le %d: %.4f ms/step, LE %.4f ms/step, %.2f steps/rebuild" % (
        world, nper, halo, with_le, st["last_run_gpu_ms"] / steps, st["last_run_le_ms"] / steps, steps / max(1, st["neigh_builds"] - b0)), flush=True)
if world > 1:
    import torch.distributed as dist
    dist.barrier()
e.close()
