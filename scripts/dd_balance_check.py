#!/usr/bin/env python3
"""Dynamic load balance check, run under torchrun: a chain that fills only part of the box in x is distributed over equal-width
slabs (imbalanced), run, re-cut with DDEngine.rebalance() (`fix balance`, src/fix_balance.cpp:191-270), run again -- and compared
with the single-GPU engine making the same two runs: the trajectory and the USER-LE topology must not notice the re-cut.
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dd_balance_check.py [BEADS] [STEPS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems
from lammps_le_b200.engine import unpack_image, pack_image
from lammps_le_b200.engine_dd import init_process_group
from tests import lehelpers as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
rank, world, local, group = init_process_group()
s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=4321, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
# the same chain in a box twice as long in x: unwrap, wrap into the long box -- the beads now crowd one part of it
lo, hi = (np.asarray(a, dtype=np.float64) for a in s["box"])
L = hi - lo
xu = s["x"] + unpack_image(s["image"]) * L
big = L * np.array([2.0, 1.0, 1.0])
w = np.floor((xu - lo) / big)
s = dict(s)
s["x"], s["image"], s["box"] = xu - w * big, pack_image(w.astype(np.int64)), (lo, lo + big)
v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
dd = systems.make_engine(s, device=local, velocities=v, dd=dict(rank=rank, world=world, halo=6.2, group=group, balance=False))
ref = systems.make_engine(s, device=local, velocities=v) if rank == 0 else None
ok = True

def report(name, good, extra=""):
    global ok
    ok = ok and good
    if rank == 0:
        print("%-40s %s %s" % (name, "ok" if good else "FAIL", extra), flush=True)

for e in (dd, ref):
    if e is None: continue
    e.fix_nve_limit(0.05); e.fix_langevin(1.0, 1.0, 1.0, 4711)
    e.fix_extrusion(100, 1, 2, 3, 0.5, 2, 4, 12345)
    e.fix_ex_load(50, 1, 1, 1.12, 2, 0.05, 684474, (1, 1), (1, 1))
    e.fix_ex_unload(50, 2, 0.5, 0.05, 456456)
dd.run(steps)
counts0 = dd.owned_counts()
before, after = dd.rebalance(thresh=1.05)
counts1 = dd.owned_counts()
report("imbalance factor goes down", before > 1.05 and after < before and after < 1.05 + 0.1,
       "%.3f -> %.3f, owned atoms %s -> %s" % (before, after, counts0.tolist(), counts1.tolist()))
dd.run(steps)
x, im = dd.positions(); topo = dd.topology(); ty = dd.types(); st = dd.stats()
if rank == 0:
    ref.run(steps); ref.run(steps)
    x2, im2 = ref.positions(); topo2 = ref.topology(); ty2 = ref.types(); st2 = ref.stats()
    res = H.compare_topology(topo, topo2)
    report("USER-LE topology after the re-cut", not any(res.values()) and (ty == ty2).all(), "loads %d/%d unloads %d/%d" % (st["loads"], st2["loads"], st["unloads"], st2["unloads"]))
    report("positions equal the single-GPU run", np.array_equal(x, x2) and np.array_equal(im, im2), "max |dx| %.2e" % np.abs(x - x2).max())
dd.barrier()
if rank == 0:
    print("DD BALANCE CHECK", "PASSED" if ok else "FAILED", flush=True)
dd.close()
sys.exit(0 if ok else 1)
