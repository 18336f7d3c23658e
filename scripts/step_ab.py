#!/usr/bin/env python3
"""A/B of step / build kernel variants at the bench size: python scripts/step_ab.py BEADS "S:B" "S:B" ...
(S = LE_STEP_VARIANT, B = LE_BUILD_VARIANT).  Every variant starts from the same relaxed state, runs 200 steps through
the graphs (whole-loop time, CUDA events) and 200 with direct launches (per-kernel averages); positions after the runs
are compared bit for bit with the first variant."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
variants = sys.argv[2:] or ["3:3", "4:3"]
s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=12345, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
e = systems.make_engine(s, velocities=v)
systems.relax(e, steps=300)
x, im = e.positions(); v = e.velocities(); e.close()
s = dict(s); s["x"], s["image"] = x, im
first = None
for var in variants:
    sv, bv = var.split(":")
    os.environ["LE_STEP_VARIANT"], os.environ["LE_BUILD_VARIANT"] = sv, bv
    e = systems.make_engine(s, velocities=v)
    e.fix_nve(True); e.fix_langevin(1.0, 1.0, 1.0, 904297)
    e.run(64)
    st0 = e.stats()
    e.run(400)
    st = e.stats()
    ms = st["last_run_gpu_ms"] / 400
    us = e.run_timed(200)
    xx, _ = e.positions()
    h = hashlib.sha1(np.ascontiguousarray(xx).tobytes()).hexdigest()[:12]
    if first is None: first = h
    print("variant step=%s build=%s kernel %s: %.2f us/step in graphs (%d rebuilds in 400 steps), step kernel %.2f us direct, positions %s %s"
          % (sv, bv, e.step_kernel_name() if hasattr(e, "step_kernel_name") else "?", 1e3 * ms, st["neigh_builds"] - st0["neigh_builds"], us, h,
             "== first" if h == first else "DIFFERENT from first"), flush=True)
    e.close()
