#!/usr/bin/env python3
"""tests/golden/ref_restart_small.{bin,npz}: a restart file the compiled reference (oracle/_ref) wrote for a 300-bead chromatin chain
with extruder bonds, plus the arrays it must read back as (from the reference's own write_data of the same state).
TEST INFRASTRUCTURE ONLY.  python oracle/make_restart_golden.py"""
import os, re, shutil, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lammps_le_b200 import systems
from oracle import refio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wd = tempfile.mkdtemp(prefix="le_restart_")
s = systems.chromatin_chain(300, 12, rho=0.2, seed=9, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
refio.write_data_file(os.path.join(wd, "data.le"), s)
deck = refio.deck_header(s, "data.le", sort=True) + ["velocity all create 1.0 4711", "timestep 0.005", "reset_timestep 4321", "run 0",
                                                     "write_restart ref.restart", "write_data ref.data"]
refio.run_reference(deck, workdir=wd, harness=False)
txt = open(os.path.join(wd, "ref.data")).read()
sec = {m.group(1): m.end() for m in re.finditer(r"^(Atoms|Velocities|Bonds)[^\n]*\n\n", txt, re.M)}
rows = lambda name: np.array([[float(v) for v in line.split()] for line in txt[sec[name]:].split("\n\n")[0].strip().splitlines()])
at = rows("Atoms"); at = at[np.argsort(at[:, 0])]
ve = rows("Velocities"); ve = ve[np.argsort(ve[:, 0])]
bo = rows("Bonds").astype(int)
n, bpa = len(at), 4
nb = np.zeros(n, np.int32); bt = np.zeros((n, bpa), np.int32); ba = np.zeros((n, bpa), np.int32)
for _, t, a, b in bo:                                   # newton_bond off: every bond on both atoms
    for i, j in ((a, b), (b, a)):
        bt[i - 1, nb[i - 1]] = t; ba[i - 1, nb[i - 1]] = j; nb[i - 1] += 1
img = ((at[:, 6].astype(int) + 512) & 1023) | (((at[:, 7].astype(int) + 512) & 1023) << 10) | (((at[:, 8].astype(int) + 512) & 1023) << 20)
lo, hi = s["box"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_restart_small.npz"), x=at[:, 3:6], v=ve[:, 1:4], type=at[:, 2].astype(np.int32),
                    image=img.astype(np.int32), num_bond=nb, bond_type=bt, bond_atom=ba, boxlo=np.asarray(lo, float), boxhi=np.asarray(hi, float),
                    ntimestep=4321, nbonds=len(bo), dt=0.005, bond_k=np.array([30.0, 10.0]), bond_r0=np.array([1.5, 4.0]))
shutil.copy(os.path.join(wd, "ref.restart"), os.path.join(ROOT, "tests", "golden", "ref_restart_small.bin"))
print("wrote tests/golden/ref_restart_small.bin (%d bytes) and .npz" % os.path.getsize(os.path.join(wd, "ref.restart")))
