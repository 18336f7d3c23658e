#!/usr/bin/env python3
"""Generate golden vectors from the compiled reference (oracle/_ref) -- run in the build container.

TEST INFRASTRUCTURE ONLY.  `python oracle/make_golden.py` rewrites tests/golden/*.npz:
  le_trace_small.npz   USER-LE event trace: state before/after every fix extrusion / ex_load / ex_unload
                       call of a 1-rank reference run (1200-bead chain), incl. the pair list, the bond list,
                       Neighbor::xhold and the Marsaglia draw counters
  forces_chain.npz     step-0 conservative forces, energies, virials, half neighbor list of a 3000-bead chain
The functions are also imported by the tests to produce larger cases on the fly when oracle/_ref is present.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refio  # noqa: E402


def le_trace(system, steps, le_lines, workdir=None, minimize=True, dt=0.005, langevin_seed=904297, damp=1.0, min_args="1e-6 1e-8 2000 20000"):
    """Run the reference with the three USER-LE fixes and le/snap fixes around them; returns (pre, post)."""
    wd = workdir or tempfile.mkdtemp(prefix="le_trace_")
    refio.write_data_file(os.path.join(wd, "data.le"), system)
    deck = refio.deck_header(system, "data.le")
    if minimize:
        deck += ["minimize " + min_args, "reset_timestep 0"]
    deck += ["fix 1 all nve", "fix 2 all langevin 1.0 1.0 %g %d" % (damp, langevin_seed)]
    deck += le_lines
    deck += ["thermo_style custom step temp epair emol bonds f_loop[1] f_loading[1] f_unloading[1]", "thermo 500",
             "timestep %g" % dt, "run %d" % steps]
    out, _ = refio.run_reference(deck, workdir=wd)
    pre = refio.read_records(os.path.join(wd, "pre.bin"))
    post = refio.read_records(os.path.join(wd, "post.bin"))
    assert len(pre) == len(post)
    return pre, post, out


def force_case(system, workdir=None, minimize=True, velocities=False, min_args="1e-4 1e-6 200 2000"):
    """run 0 of the reference without thermostat: forces, energies, virial, lists on grid-snapped positions."""
    wd = workdir or tempfile.mkdtemp(prefix="le_force_")
    refio.write_data_file(os.path.join(wd, "data.le"), system)
    deck = refio.deck_header(system, "data.le")
    if minimize:
        deck += ["minimize " + min_args, "reset_timestep 0"]
    if velocities:
        deck += ["velocity all create 1.0 4928459 dist gaussian"]
    deck += ["fix s0 all le/snap dummy.bin pre grid", "fix 1 all nve",
             "thermo_style custom step temp epair emol etotal press", "thermo_modify norm yes", "run 0"]
    final = os.path.join(wd, "final.bin")
    out, _ = refio.run_reference(deck, workdir=wd, final=final)
    rec = refio.read_records(final)[0]
    rec["thermo"] = refio.parse_thermo(out)[-1]
    return rec


_PRE_KEYS = ["x", "xhold", "image", "type", "num_bond", "bond_type", "bond_atom", "nspecial", "special", "bondlist",
             "neigh_offsets", "neigh_entries"]
_POST_KEYS = ["type", "num_bond", "bond_type", "bond_atom", "nspecial", "special"]


def pack_trace(pre, post):
    d = {"n_events": np.array(len(pre)), "boxlo": pre[0]["boxlo"], "boxhi": pre[0]["boxhi"],
         "bpa": np.array(pre[0]["bpa"]), "maxspecial": np.array(pre[0]["maxspecial"])}
    for k, (a, b) in enumerate(zip(pre, post)):
        d["ev%d_meta" % k] = np.array([a["step"], a["which"]] + a["rngc"] + b["rngc"] + b["counters"], dtype=np.int64)
        for key in _PRE_KEYS:
            v = a[key]
            if key in ("x", "xhold"):
                v = v  # float64, exact
            d["ev%d_pre_%s" % (k, key)] = v
        for key in _POST_KEYS:
            d["ev%d_post_%s" % (k, key)] = b[key]
    return d


def unpack_trace(npz):
    n = int(npz["n_events"])
    pre, post = [], []
    for k in range(n):
        meta = npz["ev%d_meta" % k]
        a = {"step": int(meta[0]), "which": int(meta[1]), "rngc": [int(v) for v in meta[2:5]], "boxlo": npz["boxlo"],
             "boxhi": npz["boxhi"], "bpa": int(npz["bpa"]), "maxspecial": int(npz["maxspecial"])}
        b = {"step": int(meta[0]), "which": int(meta[1]), "rngc": [int(v) for v in meta[5:8]], "counters": [int(v) for v in meta[8:12]]}
        for key in _PRE_KEYS:
            a[key] = npz["ev%d_pre_%s" % (k, key)]
        for key in _POST_KEYS:
            b[key] = npz["ev%d_post_%s" % (k, key)]
        a["n"] = len(a["type"])
        pre.append(a)
        post.append(b)
    return pre, post


BOND_CREATE_CFG = dict(nevery=20, itype=1, jtype=1, rc=1.05, btype=2, prob=0.5, seed=456456, iparam=(2, 4), jparam=(2, 4))


def bond_create_trace(nbeads=1200, steps=200, cfg=BOND_CREATE_CFG):
    """a run of the compiled reference with fix bond/create (src/MC, the ancestor of fix ex_load) between two le/snap fixes:
    (pre, post) records of every event"""
    from lammps_le_b200 import systems
    s = systems.chromatin_chain(nbeads, nbeads * 3 // 100, rho=0.2, seed=11, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    wd = tempfile.mkdtemp(prefix="le_bc_trace_")
    refio.write_data_file(os.path.join(wd, "data.le"), s)
    deck = refio.deck_header(s, "data.le") + [
        "minimize 1e-6 1e-8 2000 20000", "reset_timestep 0", "fix 1 all nve", "fix 2 all langevin 1.0 1.0 1.0 904297",
        "fix s0 all le/snap pre.bin pre grid",
        "fix cr all bond/create %d %d %d %g %d prob %g %d iparam %d %d jparam %d %d" % (
            cfg["nevery"], cfg["itype"], cfg["jtype"], cfg["rc"], cfg["btype"], cfg["prob"], cfg["seed"], *cfg["iparam"], *cfg["jparam"]),
        "fix s1 all le/snap post.bin post",
        "thermo_style custom step temp bonds f_cr[1] f_cr[2]", "thermo 100", "timestep 0.005", "run %d" % steps]
    refio.run_reference(deck, workdir=wd)
    pre = refio.read_records(os.path.join(wd, "pre.bin"))
    post = refio.read_records(os.path.join(wd, "post.bin"))
    assert len(pre) == len(post) == steps // cfg["nevery"]
    return pre, post


def main():
    from lammps_le_b200 import systems
    from tests.lehelpers import le_deck_lines
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    s = systems.chromatin_chain(1200, 24, rho=0.2, seed=11)
    pre, post, _ = le_trace(s, 1510, le_deck_lines())
    # keep: the 4 extrusion events and a subset of load/unload events (all those that changed something + a few idle)
    ext, other = [], []
    for k, (a, b) in enumerate(zip(pre, post)):
        changed = (a["num_bond"] != b["num_bond"]).any() or (a["bond_atom"] != b["bond_atom"]).any()
        if a["which"] == 1:
            ext.append(k)
        elif changed:
            other.append(k)
    keep = sorted(ext + other[:10])
    np.savez_compressed(os.path.join(gold, "le_trace_small.npz"), **pack_trace([pre[k] for k in keep], [post[k] for k in keep]))
    print("le_trace_small: %d of %d events kept" % (len(keep), len(pre)))
    s2 = systems.chromatin_chain(3000, 40, rho=0.2, seed=5)
    rec = force_case(s2)
    np.savez_compressed(os.path.join(gold, "forces_chain.npz"),
                        **{k: rec[k] for k in ["x", "image", "type", "num_bond", "bond_type", "bond_atom", "nspecial", "special", "f",
                                               "bondlist", "neigh_offsets", "neigh_entries", "boxlo", "boxhi", "virial_pair", "virial_bond"]},
                        energy=np.array(rec["energy"]), bpa=np.array(rec["bpa"]), maxspecial=np.array(rec["maxspecial"]),
                        thermo=np.array([rec["thermo"][k] for k in ("Temp", "E_pair", "E_mol", "TotEng", "Press")]))
    print("forces_chain written")
    # fix bond/create: four events of a 600-bead run (the later ones see beads that already carry created bonds and beads that have
    # changed type)
    pre, post = bond_create_trace(600, 200)
    keep = [1, 2, 5, 9]
    np.savez_compressed(os.path.join(gold, "bond_create_trace_small.npz"), **pack_trace([pre[k] for k in keep], [post[k] for k in keep]))
    print("bond_create_trace_small: %d of %d events kept, %d bonds created in them" % (len(keep), len(pre), sum(post[k]["counters"][2] for k in keep)))


if __name__ == "__main__":
    main()
