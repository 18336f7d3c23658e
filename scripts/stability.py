#!/usr/bin/env python3
"""Long-run stability probe of a bench workload: python scripts/stability.py BEADS EXTRUDERS STEPS [harmonic|fene]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems
n, ne, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 else "fene"
eb = systems.EXTRUDER_FENE if kind == "fene" else systems.EXTRUDER_HARMONIC
s = systems.chromatin_chain(n, ne, rho=0.2, seed=12345, barriers="random", extruder_bond=eb)
v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
e = systems.make_engine(s, velocities=v, dt=0.005)
t0 = time.time()
systems.relax(e, steps=1500)
print("relax %.1fs" % (time.time() - t0), e.thermo(-1))
e.fix_langevin(1.0, 1.0, 1.0, 904297)
e.fix_extrusion(500, 1, 2, 3, 0.5, 2, 4, 12345)
e.fix_ex_load(100, 1, 1, 1.12, 2, 0.02, 684474, (1, 1), (1, 1))
e.fix_ex_unload(100, 2, 0.5, 0.05, 456456)
e.reset_timestep(0)
e.thermo_every(500)
done = 0
while done < steps:
    k = min(1000, steps - done)
    e.run(k)
    done += k
    t = e.thermo(-1); st = e.stats()
    print(done, "T=%.3f ep=%.4f em=%.4f bonds=%d fenewarn=%d builds=%d shifts=%d loads=%d unloads=%d ms/step=%.4f" % (
        t["temp"], t["epair"], t["emol"], t["nbonds"], t["fene_warnings"], st["neigh_builds"], st["extrusion_shifts"], st["loads"], st["unloads"], st["last_run_gpu_ms"] / k), flush=True)
