#!/usr/bin/env python3
"""Development aid: duration of k_step<0> with parts switched off (LE_STEP_SKIP bits: 1 gathers, 2 pair fp64, 4 bonds, 8 neighbor rows, 16 Langevin)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems
n = 1000000
s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=12345, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
e = systems.make_engine(s, velocities=v)
if not os.environ.get("LE_STEP_SKIP"):
    systems.relax(e, steps=300)
e.fix_nve(True); e.fix_langevin(1.0, 1.0, 1.0, 904297)
us = e.run_timed(30)
print("skip=%s k_step %.2f us" % (os.environ.get("LE_STEP_SKIP", "0"), us), flush=True)
