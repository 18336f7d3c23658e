"""Restart files (SURVEY.md 8f "wire formats"): the reference's binary format for this path, read and written by the host-only
entry points le_host_restart_* (lammps_le_b200/csrc/le_restart.cpp; Python: lammps_le_b200/restart.py).  CPU tests.
  * a restart file the compiled reference wrote (tests/golden/ref_restart_small.bin, made by oracle/make_restart_golden.py from a
    chromatin chain with extruder bonds) is read back: header, coefficients, atoms, bond tables equal the recorded arrays;
  * with oracle/_ref present: the reference writes -> we read; we write -> the reference's read_restart accepts the file, `run 0`
    prints the same thermo line as from its own restart, and write_data gives back the same atoms, velocities and bonds."""
import os
import re

import numpy as np
import pytest

from lammps_le_b200 import restart as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def bond_sets(num_bond, bond_type, bond_atom):
    return [frozenset((int(bond_type[t, m]), int(bond_atom[t, m])) for m in range(num_bond[t])) for t in range(len(num_bond))]


def test_reads_a_restart_file_the_reference_wrote():
    z = np.load(os.path.join(GOLD, "ref_restart_small.npz"))
    info, a = RS.read_restart(os.path.join(GOLD, "ref_restart_small.bin"))
    n = len(z["x"])
    assert info["natoms"] == n and info["units"] == "lj" and info["atom_style"] == "bond" and info["pair_style"] == "lj/cut"
    assert info["ntimestep"] == int(z["ntimestep"]) and info["nbonds"] == int(z["nbonds"]) and info["bond_style"] == "fene"
    assert info["newton_pair"] == 1 and info["newton_bond"] == 0 and list(info["special_lj"]) == [0.0, 1.0, 1.0]
    assert np.array_equal(info["boxlo"], z["boxlo"]) and np.array_equal(info["boxhi"], z["boxhi"]) and info["dt"] == float(z["dt"])
    assert info["offset_flag"] == 1 and info["pair_setflag"][0][0] == 1 and info["pair_cut"][0][0] == 1.12246
    assert np.array_equal(info["bond_k"], z["bond_k"]) and np.array_equal(info["bond_r0"], z["bond_r0"])
    assert np.array_equal(a["x"], z["x"]) and np.array_equal(a["v"], z["v"]) and np.array_equal(a["type"], z["type"])
    assert np.array_equal(a["image"], z["image"]) and np.array_equal(a["num_bond"], z["num_bond"])
    assert bond_sets(a["num_bond"], a["bond_type"], a["bond_atom"]) == bond_sets(z["num_bond"], z["bond_type"], z["bond_atom"])
    assert (a["num_bond"] == 3).sum() > 0, "the fixture must carry extruder bonds"


def test_write_then_read_round_trip(tmp_path):
    info, a = RS.read_restart(os.path.join(GOLD, "ref_restart_small.bin"))
    p = str(tmp_path / "again.restart")
    RS.write_restart_arrays(p, info, a)
    info2, b = RS.read_restart(p)
    for k in ("natoms", "nbonds", "ntimestep", "dt", "bond_style", "pair_style", "maxspecial", "bond_per_atom", "extra_bond_per_atom"):
        assert info[k] == info2[k], k
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    # what the reference wrote and what we write are the same bytes up to the version string and the atom order
    assert abs(os.path.getsize(p) - os.path.getsize(os.path.join(GOLD, "ref_restart_small.bin"))) < 64


def test_rejects_what_is_not_a_restart_file(tmp_path):
    p = tmp_path / "junk"
    p.write_bytes(b"not a restart file at all, sorry" * 4)
    with pytest.raises(Exception, match="Invalid LAMMPS restart file"):
        RS.read_restart(str(p))
    with pytest.raises(Exception, match="Cannot open restart file"):
        RS.read_restart(str(tmp_path / "missing"))


def parse_data(path):
    txt = open(path).read()
    sec = {m.group(1): m.end() for m in re.finditer(r"^(Atoms|Velocities|Bonds)[^\n]*\n\n", txt, re.M)}

    def rows(name):
        body = txt[sec[name]:].split("\n\n")[0]
        return np.array([[float(v) for v in line.split()] for line in body.strip().splitlines()])
    at = rows("Atoms"); at = at[np.argsort(at[:, 0])]
    ve = rows("Velocities"); ve = ve[np.argsort(ve[:, 0])]
    bo = rows("Bonds").astype(int)
    return at, ve, {(int(r[1]), min(int(r[2]), int(r[3])), max(int(r[2]), int(r[3]))) for r in bo}


def test_both_directions_against_the_compiled_reference(tmp_path):
    from oracle import refio
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    from lammps_le_b200 import systems
    s = systems.chromatin_chain(800, 16, rho=0.2, seed=5, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    refio.write_data_file(str(tmp_path / "data.le"), s)
    head = refio.deck_header(s, "data.le", sort=True)
    out, _ = refio.run_reference(head + ["velocity all create 1.0 4711", "timestep 0.005", "reset_timestep 1234", "thermo_style custom step temp epair emol bonds",
                                         "run 0", "write_restart ref.restart", "write_data ref.data"], workdir=str(tmp_path), harness=False)
    # the reference wrote -> we read
    info, a = RS.read_restart(str(tmp_path / "ref.restart"))
    at, ve, bonds = parse_data(str(tmp_path / "ref.data"))
    assert info["ntimestep"] == 1234 and info["natoms"] == 800 and info["nbonds"] == len(bonds)
    assert np.array_equal(a["x"], at[:, 3:6]) and np.array_equal(a["v"], ve[:, 1:4]) and np.array_equal(a["type"], at[:, 2].astype(int))
    mine = {(int(a["bond_type"][t, m]), min(t + 1, int(a["bond_atom"][t, m])), max(t + 1, int(a["bond_atom"][t, m])))
            for t in range(800) for m in range(a["num_bond"][t])}
    assert mine == bonds
    # we write -> the reference reads
    RS.write_restart_arrays(str(tmp_path / "ours.restart"), info, a)
    tail = ["thermo_style custom step temp epair emol bonds", "run 0", "write_data back.data"]
    o1, _ = refio.run_reference(["read_restart ref.restart"] + tail, workdir=str(tmp_path), harness=False)
    o2, _ = refio.run_reference(["read_restart ours.restart"] + tail, workdir=str(tmp_path), harness=False)
    row = lambda o: re.search(r"Step Temp E_pair E_mol Bonds \n\s*(.*?)\n", o).group(1).split()
    assert row(o1) == row(o2) and row(o2)[0] == "1234"
    at2, ve2, bonds2 = parse_data(str(tmp_path / "back.data"))
    assert np.array_equal(at2, at) and np.array_equal(ve2, ve) and bonds2 == bonds
