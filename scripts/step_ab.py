#!/usr/bin/env python3
"""A/B of the plain step kernels on one GPU: python scripts/step_ab.py [BEADS] [VARIANTS...]

For every LE_STEP_VARIANT (see plain_step_kernel in csrc/le_engine.cu; 0 = k_step, the reference arm of the comparison)
the same relaxed 1M-bead chromatin state is uploaded, run for 64 + 40 timesteps (captured graphs, then direct launches)
and downloaded again: positions, image flags and velocities must be IDENTICAL, bit for bit, to variant 0 -- k_step2
does the same arithmetic in the same order.  Printed per variant: the live duration of the step kernel (CUDA events,
le_run_timed) and the whole-loop time per MD step of a 400-step graph run (rebuilds included).
Results are appended to gpurun_out/step_ab.txt as they come."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
variants = [int(v) for v in sys.argv[2:]] or [0, 1, 3, 33, 35, 545, 1057]
if variants[0] != 0:
    variants = [0] + variants
os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/step_ab.txt", "a")


def say(msg):
    print(msg, flush=True)
    log.write(msg + "\n")
    log.flush()


t0 = time.time()
os.environ["LE_STEP_VARIANT"] = "0"
s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=12345, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
e = systems.make_engine(s, velocities=v)
systems.relax(e, steps=300)
e.fix_nve(True)
e.fix_langevin(1.0, 1.0, 1.0, 904297)
start = e.owned_buffers(pinned=False)
n0 = e.download_owned(start)
say("# %d beads, set-up %.1f s" % (n, time.time() - t0))

ref = None
best = (None, 1e9)
for var in variants:
    os.environ["LE_STEP_VARIANT"] = str(var)
    try:
        e.upload_owned(n0, start)
        e.reset_timestep(0)
        e.run(64)
        us = e.run_timed(40)
        got = e.owned_buffers(pinned=False)
        ng = e.download_owned(got)
        order = np.argsort(got[0][:ng], kind="stable")
        state = (got[0][:ng][order], got[1][:ng][order], got[2][:ng][order], got[3][:ng][order])
        if ref is None:
            ref, same = state, True
        else:
            same = ng == n0 and all(np.array_equal(a, b) for a, b in zip(state, ref))
        e.run(400)
        st = e.stats()
        ms = st["last_run_gpu_ms"] / 400
        say("variant %2d  identical=%s  step kernel %.2f us  whole loop %.4f ms/step" % (var, same, us, ms))
        if same and var and ms < best[1]:
            best = (var, ms)
    except Exception as ex:  # keep going: the next variant may be fine (a sticky CUDA error ends them all)
        say("variant %2d  FAILED: %s" % (var, str(ex)[:300]))
say("best identical variant by whole-loop time: %s (%.4f ms/step)" % best)
with open("gpurun_out/step_ab_best.txt", "w") as f:
    f.write(str(best[0] if best[0] is not None else 0))
