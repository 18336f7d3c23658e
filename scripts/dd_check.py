#!/usr/bin/env python3
"""Multi-GPU check, run under torchrun: the slab-decomposed engine against the single-GPU engine on the same system.
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dd_check.py [BEADS] [STEPS]
rank 0 also runs the whole system on its own GPU; compared: half neighbor lists, bond list, step-0 forces/energies,
positions after a short Langevin run (same counter-based noise), thermo after a longer one, atom conservation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems
from lammps_le_b200.engine_dd import init_process_group
from tests import lehelpers as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
rank, world, local, group = init_process_group()
melt = len(sys.argv) > 3 and sys.argv[3] == "melt"
mc = len(sys.argv) > 3 and sys.argv[3] == "mc"      # fix bond/create + fix bond/break (src/MC) instead of the three USER-LE fixes
if melt:
    # BASELINE configs[2], bench/in.chain.scaled: the 32,000-bead FENE melt of bench/data.chain replicated `world` times along x
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bench_chain.npz"))
    z = {k: z[k] for k in z.files}
    from lammps_le_b200.engine import pack_image
    base = {"name": "melt", "box": (z["boxlo"], z["boxhi"]), "types": z["type"], "x": z["x"], "image": pack_image(z["image"]),
            "bonds": (z["bonds"][:, 0].copy(), z["bonds"][:, 1].copy(), z["bonds"][:, 2].copy()), "masses": np.ones(1), "nbondtypes": 1,
            "ntypes": 1, "bond_coeffs": {1: ("fene", (30.0, 1.5, 1.0, 1.0))}, "bond_per_atom": 2, "maxspecial": 8, "v": z["v"]}
    s = systems.replicate(base, world, 1, 1)
    n = len(s["types"])
    v = s["v"]
    halo = 0.0
else:
    s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=12345, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    v = systems.maxwell_velocities(n, 1.0, np.ones(n), 1)
    halo = 6.2
    if mc:
        s["bond_per_atom"] = 6      # two backbone bonds + up to two extruder bonds + two created ones
dd = systems.make_engine(s, device=local, velocities=v, dd=dict(rank=rank, world=world, halo=halo, group=group))
ref = systems.make_engine(s, device=local, velocities=v) if rank == 0 else None
ok = True

def report(name, good, extra=""):
    global ok
    ok = ok and good
    if rank == 0:
        print("%-34s %s %s" % (name, "ok" if good else "FAIL", extra), flush=True)

dd.force_rebuild()
off, ent = dd.neighlist(half=True)
bl = dd.bondlist()
if rank == 0:
    ref.force_rebuild()
    o2, e2 = ref.neighlist(half=True)
    bad = H.compare_neighlists(off, ent, o2, e2)
    report("half neighbor list", not bad and len(ent) == len(e2), "%d entries, %d atoms differ" % (len(ent), len(bad)))
    b2 = ref.bondlist()
    report("bond list", bl.shape == b2.shape and (bl == b2).all(), "%d rows" % len(bl))
f, th = dd.compute_forces()
if rank == 0:
    f2, th2 = ref.compute_forces()
    mag = np.sqrt((f2 ** 2).sum(1)); rms = np.sqrt((mag ** 2).mean())
    err = (np.sqrt(((f - f2) ** 2).sum(1)) / np.maximum(mag, rms)).max()
    report("step-0 forces", err < 1e-10, "max rel err %.2e" % err)
    report("step-0 energies", abs(th["epair"] - th2["epair"]) < 1e-10 and abs(th["emol"] - th2["emol"]) < 1e-9 * abs(th2["emol"]),
           "epair %.12g emol %.12g" % (th["epair"], th["emol"]))
for e in (dd, ref):
    if e is None: continue
    e.fix_nve_limit(0.05); e.fix_langevin(1.0, 1.0, 1.0, 4711)
short = 40
dd.run(short)
x, im = dd.positions()
if rank == 0:
    ref.run(short)
    x2, im2 = ref.positions()
    L = s["box"][1][0]
    dx = np.abs(((x - x2) + L / 2) % L - L / 2).max()
    report("positions after %d steps" % short, dx < 1e-4, "max |dx| %.2e" % dx)
for e in (dd, ref):
    if e is None: continue
    e.fix_nve(True); e.fix_langevin(1.0, 1.0, 1.0, 904297); e.thermo_every(100)
t0 = time.time(); dd.run(steps); t1 = time.time()
st = dd.stats(); th = dd.thermo(-1)
x, im = dd.positions()
cnt = np.isfinite(x).all() and (np.abs(x).sum(1) > 0).sum()
if rank == 0:
    ref.run(steps)
    th2 = ref.thermo(-1); st2 = ref.stats()
    report("atoms conserved", cnt >= n - 1, "%d of %d have an owner" % (cnt, n))
    report("thermo after %d steps" % steps, abs(th["temp"] - th2["temp"]) < 0.02 and abs(th["emol"] - th2["emol"]) < 0.05 and abs(th["epair"] - th2["epair"]) < 0.01,
           "T %.4f/%.4f epair %.4f/%.4f emol %.4f/%.4f builds %d/%d" % (th["temp"], th2["temp"], th["epair"], th2["epair"], th["emol"], th2["emol"], st["neigh_builds"], st2["neigh_builds"]))
    print("dd: %.4f ms/step (GPU events, rank 0) over %d ranks; single GPU %.4f ms/step" % (st["last_run_gpu_ms"] / steps, world, st2["last_run_gpu_ms"] / steps), flush=True)
if melt:
    dd.barrier()
    if rank == 0:
        print("DD CHECK", "PASSED" if ok else "FAILED", flush=True)
    dd.close()
    sys.exit(0 if ok else 1)
# USER-LE events on the decomposed system: the decision logic is replicated, the geometry comes from the owners
for e in (dd, ref):
    if e is None: continue
    if mc:
        e.fix_bond_create(10, 1, 1, 1.05, 2, 0.5, 684474, (2, 4), (2, 4))
        e.fix_bond_break(10, 2, 1.25, 0.5, 456456)
    else:
        e.fix_extrusion(200, 1, 2, 3, 0.5, 2, 4, 12345)
        e.fix_ex_load(100, 1, 1, 1.12, 2, 0.02, 684474, (1, 1), (1, 1))
        e.fix_ex_unload(100, 2, 0.5, 0.05, 456456)
    e.reset_timestep(0)
le_steps = 120 if mc else 650
dd.run(le_steps)
st = dd.stats(); topo = dd.topology(); ty = dd.types(); x, im = dd.positions()
if rank == 0:
    ref.run(le_steps)
    st2 = ref.stats(); topo2 = ref.topology(); ty2 = ref.types(); x2, im2 = ref.positions()
    res = H.compare_topology(topo, topo2)
    if mc:
        report("bonds were created and broken", st2["loads"] > 50 and st2["unloads"] > 5, "%d / %d" % (st2["loads"], st2["unloads"]))
    report("%s topology after %d steps" % ("bond/create + bond/break" if mc else "USER-LE", le_steps), not any(res.values()) and (ty == ty2).all(),
           "shifts %d/%d loads %d/%d unloads %d/%d %s" % (st["extrusion_shifts"], st2["extrusion_shifts"], st["loads"], st2["loads"],
                                                        st["unloads"], st2["unloads"], {k: v for k, v in res.items() if v}))
    L = s["box"][1][0]
    dx = np.abs(((x - x2) + L / 2) % L - L / 2).max()
    report("positions after the LE run", dx < 1e-3, "max |dx| %.2e" % dx)
    print("dd with LE: %.4f ms/step; single GPU %.4f ms/step" % (st["last_run_gpu_ms"] / le_steps, st2["last_run_gpu_ms"] / le_steps), flush=True)
dd.barrier()
if rank == 0:
    print("DD CHECK", "PASSED" if ok else "FAILED", flush=True)
dd.close()
sys.exit(0 if ok else 1)
