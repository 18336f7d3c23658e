#!/usr/bin/env python3
"""Long-run observables of the compiled reference (oracle/_ref) for the statistical parity test.

TEST INFRASTRUCTURE ONLY.  `python oracle/make_stats_golden.py` (build container, ~2 min of CPU) writes
tests/golden/stats_chain.npz:
  * the relaxed, thermalised start state the reference produced (positions, velocities, images) -- the GPU test
    starts from exactly this state,
  * per-frame observables of TWO reference runs with different Langevin seeds (the seed-to-seed scatter is the
    empirical statistical error the test's tolerances are stated in): temperature, E_pair, E_mol, radius of
    gyration (compute gyration, unwrapped), number of extruder bonds, mean loop size, contact probability P(s).
north_star: "Thermodynamic and polymer observables (Rg, contact-probability P(s), loop-size distribution) over
long runs must agree within stated statistical error."
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refio  # noqa: E402

N, NEXT, STEPS, EVERY = 4000, 60, 60000, 1000
S_LIST = np.array([2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64])
CONTACT = 1.5
LE_LINES = ["fix loop all extrusion 200 1 2 3 0.5 2 4",
            "fix loading all ex_load 100 1 1 1.12 2 prob 0.05 684474 iparam 1 1 jparam 1 1",
            "fix unloading all ex_unload 100 2 0.5 prob 0.02 456456"]


def observables(xu, bonds2):
    """xu: unwrapped positions in tag order; bonds2: (a, b) tag pairs of the extruder bonds"""
    com = xu.mean(0)
    rg = np.sqrt(((xu - com) ** 2).sum(1).mean())
    ps = np.array([(np.sqrt(((xu[s:] - xu[:-s]) ** 2).sum(1)) < CONTACT).mean() for s in S_LIST])
    sizes = np.abs(bonds2[:, 1] - bonds2[:, 0]) if len(bonds2) else np.zeros(0)
    return rg, ps, sizes


def read_dump_atoms(path):
    frames = []
    with open(path) as f:
        lines = f.read().splitlines()
    k = 0
    while k < len(lines):
        assert lines[k].startswith("ITEM: TIMESTEP")
        n = int(lines[k + 3])
        rows = np.array([l.split() for l in lines[k + 9:k + 9 + n]], dtype=np.float64)
        frames.append((int(lines[k + 1]), rows[np.argsort(rows[:, 0])][:, 1:4]))
        k += 9 + n
    return frames


def read_dump_local(path):
    frames = []
    with open(path) as f:
        lines = f.read().splitlines()
    k = 0
    while k < len(lines):
        n = int(lines[k + 3])
        rows = np.array([l.split() for l in lines[k + 9:k + 9 + n]], dtype=np.float64).reshape(n, 3) if n else np.zeros((0, 3))
        frames.append((int(lines[k + 1]), rows[rows[:, 2] == 2][:, :2].astype(np.int64)))
        k += 9 + n
    return frames


def relaxed_start(system, wd):
    """minimize + 4000 thermostatted steps in the reference; returns x (wrapped), image, v from its write_data"""
    refio.write_data_file(os.path.join(wd, "data.le"), system)
    deck = refio.deck_header(system, "data.le")
    deck += ["minimize 1e-6 1e-8 2000 20000", "reset_timestep 0", "velocity all create 1.0 4928459 dist gaussian",
             "fix 1 all nve", "fix 2 all langevin 1.0 1.0 1.0 12345", "timestep 0.005", "run 4000", "write_data relaxed.data nocoeff"]
    refio.run_reference(deck, workdir=wd, harness=False)
    lines = open(os.path.join(wd, "relaxed.data")).read().splitlines()
    n = len(system["types"])
    ia = [k for k, l in enumerate(lines) if l.startswith("Atoms")][0]
    iv = [k for k, l in enumerate(lines) if l.startswith("Velocities")][0]
    a = np.array([l.split() for l in lines[ia + 2:ia + 2 + n]], dtype=np.float64)
    v = np.array([l.split() for l in lines[iv + 2:iv + 2 + n]], dtype=np.float64)
    a = a[np.argsort(a[:, 0])]; v = v[np.argsort(v[:, 0])]
    img = a[:, 6:9].astype(np.int64)
    image = (((img[:, 0] + 512) & 1023) | (((img[:, 1] + 512) & 1023) << 10) | (((img[:, 2] + 512) & 1023) << 20)).astype(np.int32)
    return a[:, 3:6].copy(), image, v[:, 1:4].copy()


def production(system, x, image, v, seed, wd):
    s2 = dict(system)
    s2["x"], s2["image"], s2["v"] = x, image, v
    refio.write_data_file(os.path.join(wd, "start.data"), s2)
    deck = refio.deck_header(s2, "start.data")
    deck += ["fix 1 all nve", "fix 2 all langevin 1.0 1.0 1.0 %d" % seed] + LE_LINES
    deck += ["compute bl all property/local batom1 batom2 btype", "dump 1 all custom %d atoms.dump id xu yu zu" % EVERY,
             "dump_modify 1 format float %.10g", "dump 2 all local %d bonds.dump c_bl[1] c_bl[2] c_bl[3]" % EVERY,
             "thermo_style custom step temp epair emol bonds", "thermo %d" % EVERY, "timestep 0.005", "run %d" % STEPS]
    out, _ = refio.run_reference(deck, workdir=wd, harness=False, timeout=3600)
    th = refio.parse_thermo(out)
    fa, fb = read_dump_atoms(os.path.join(wd, "atoms.dump")), read_dump_local(os.path.join(wd, "bonds.dump"))
    rows = []
    sizes_all = []
    for (st, xu), (st2, b2), t in zip(fa, fb, th):
        assert st == st2 == int(t["Step"])
        rg, ps, sizes = observables(xu, b2)
        rows.append([st, t["Temp"], t["E_pair"], t["E_mol"], rg, len(sizes), sizes.mean() if len(sizes) else 0.0] + ps.tolist())
        sizes_all.append(sizes)
    return np.array(rows), np.concatenate(sizes_all[len(sizes_all) // 2:])


def main():
    from lammps_le_b200 import systems
    system = systems.chromatin_chain(N, NEXT, rho=0.2, seed=21, barriers="periodic", extruder_bond=systems.EXTRUDER_FENE)
    wd = tempfile.mkdtemp(prefix="le_stats_")
    x, image, v = relaxed_start(system, wd)
    runs, sizes = [], []
    for seed in (904297, 31337):
        r, sz = production(system, x, image, v, seed, tempfile.mkdtemp(prefix="le_stats_run_"))
        runs.append(r); sizes.append(sz)
        print("seed %d: T %.4f Rg %.3f loops %.1f mean size %.2f P(2) %.4f" % (seed, r[10:, 1].mean(), r[10:, 4].mean(), r[10:, 5].mean(), r[10:, 6].mean(), r[10:, 7].mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stats_chain.npz"), x=x, image=image, v=v, run_a=runs[0], run_b=runs[1],
                        sizes_a=sizes[0], sizes_b=sizes[1], s_list=S_LIST, columns=np.array(["step", "temp", "epair", "emol", "rg", "nloops", "mean_loop"] + ["P(%d)" % s for s in S_LIST]),
                        params=np.array([N, NEXT, STEPS, EVERY]))
    print("wrote tests/golden/stats_chain.npz")


if __name__ == "__main__":
    main()
