#!/usr/bin/env python3
"""Per-kernel timing of the MD loop on one GPU: python scripts/perf_1m.py [BEADS] [STEPS] [le]
Builds the bench system (chromatin chain, rho 0.2, 1 % extruders), relaxes it, then
  1. runs STEPS timesteps through the captured graphs          -> whole-loop ms per MD step (CUDA events)
  2. runs STEPS timesteps with direct launches + event marks    -> per-kernel table (LE_B200_TIMING=1, printed by the library)
`le` adds the USER-LE fixes at the bench cadence."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LE_B200_TIMING"] = "1"
import numpy as np
from lammps_le_b200 import systems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
with_le = len(sys.argv) > 3 and sys.argv[3] == "le"
t0 = time.time()
s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=12345, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
e = systems.make_engine(s, velocities=v)
systems.relax(e, steps=300)
e.fix_nve(True)
e.fix_langevin(1.0, 1.0, 1.0, 904297)
if with_le:
    e.fix_extrusion(500, 1, 2, 3, 0.5, 2, 4, 12345)
    e.fix_ex_load(100, 1, 1, 1.12, 2, 0.01, 684474, (1, 1), (1, 1))
    e.fix_ex_unload(100, 2, 0.5, 0.05, 456456)
e.reset_timestep(0)
print("# %d beads, set-up %.1f s" % (n, time.time() - t0), flush=True)
e.run(64)
st0 = e.stats()
e.run(steps)
st = e.stats()
print("graphs: %.4f ms/step over %d steps, %d rebuilds (every %.2f steps), nbar_full %.3f, T %.3f" % (
    st["last_run_gpu_ms"] / steps, steps, st["neigh_builds"] - st0["neigh_builds"],
    steps / max(1, st["neigh_builds"] - st0["neigh_builds"]), st["full_entries"] / n, e.thermo(-1)["temp"]), flush=True)
us = e.run_timed(steps)
print("direct: step kernel %.2f us" % us, flush=True)
# in-graph cost of a rebuild: a rebuild on every step (check no) against the normal cadence K: t_all - t = rebuild (1 - 1/K)
t_norm = st["last_run_gpu_ms"] / steps
K = steps / max(1, st["neigh_builds"] - st0["neigh_builds"])
e.set_neighbor(0.4, 1, 0, 0)
e.run(64)
e.run(steps)
t_all = e.stats()["last_run_gpu_ms"] / steps
rb = (t_all - t_norm) / (1.0 - 1.0 / K)
print("in-graph: rebuild every step %.2f us/step, normal %.2f us/step -> rebuild %.2f us, step without rebuild %.2f us" % (1e3 * t_all, 1e3 * t_norm, 1e3 * rb, 1e3 * (t_all - rb)), flush=True)
e.close()
