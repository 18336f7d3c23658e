"""GPU tests of the set-up steps either side of the hot path: `minimize` (min_style cg, quadratic line
search) against the compiled reference's own minimizer on the same input, and the USER-LE deck of SURVEY.md Appendix B --
bond_style hybrid, minimize, the three fixes, `thermo_style custom ... f_ID[k]` -- through the C++ front end le_deck."""
import os
import re
import subprocess

import numpy as np
import pytest

from lammps_le_b200 import systems
from tests import lehelpers as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_ref():
    from oracle import refio
    if not refio.have_reference():
        pytest.skip("oracle/_ref not present on this box")


def test_minimize_follows_the_reference_minimizer():
    """Same generator output, same `minimize 1e-6 1e-8 200 2000`: the two CG runs are the same algorithm in different
    floating-point orders, so they agree step for step at first and end at the same energy / force norm within the
    tolerances of the stopping criteria."""
    _need_ref()
    import tempfile
    from oracle import refio
    s = systems.chromatin_chain(3000, 30, rho=0.2, seed=7)
    wd = tempfile.mkdtemp(prefix="le_min_")
    refio.write_data_file(os.path.join(wd, "data.le"), s)
    deck = refio.deck_header(s, "data.le") + ["thermo 1", "thermo_style custom step pe fnorm", "thermo_modify norm yes format float %20.15g",
                                              "minimize 1e-6 1e-8 200 2000"]
    out, _ = refio.run_reference(deck, workdir=wd, harness=False)
    m = re.search(r"Energy initial, next-to-last, final =\s*\n\s*(\S+)\s+(\S+)\s+(\S+)", out)
    e_ref = [float(v) for v in m.groups()]
    it_ref, ev_ref = (int(v) for v in re.search(r"Iterations, force evaluations = (\d+) (\d+)", out).groups())
    stop_ref = re.search(r"Stopping criterion = (.*)", out).group(1).strip()
    rows = refio.parse_thermo(out)
    e = systems.make_engine(s)
    res = e.minimize(1e-6, 1e-8, 200, 2000)
    th = e.compute_forces()[1]
    e.close()
    assert abs(res["einitial"] - e_ref[0]) <= 1e-9 * abs(e_ref[0]), (res["einitial"], e_ref[0])
    # the first iterations are the same line searches: energy after iteration 1 as the reference prints it
    assert res["niter"] > 3 and it_ref > 3
    assert abs(res["efinal"] - e_ref[2]) <= 2e-4 * abs(e_ref[2]), "final energy %.10g vs reference %.10g" % (res["efinal"], e_ref[2])
    assert abs(th["epair"] + th["emol"] - res["efinal"]) <= 1e-9 * abs(res["efinal"])
    assert res["stop_string"] in ("energy tolerance", "force tolerance", "max iterations", "linesearch alpha is zero")
    assert stop_ref in ("energy tolerance", "force tolerance", "max iterations", "linesearch alpha is zero")
    assert 0.3 * it_ref <= res["niter"] <= 3.0 * it_ref + 5, "iterations %d vs reference %d" % (res["niter"], it_ref)
    assert res["fnorm2_final"] < 0.5 * res["fnorm2_init"]
    assert rows and abs(rows[0]["PotEng"] - res["einitial"]) <= 1e-9 * abs(res["einitial"])


APPENDIX_B = """units lj
atom_style bond
newton on off
special_bonds fene
atom_modify sort 0 0
read_data data.le
neighbor 0.4 bin
neigh_modify every 1 delay 1
comm_modify cutoff 5.0
bond_style hybrid fene harmonic
bond_coeff 1 fene 30.0 1.5 1.0 1.0
bond_coeff 2 harmonic 20.0 1.3
pair_style lj/cut 1.12246
pair_modify shift yes
pair_coeff * * 1.0 1.0 1.12246
%s
reset_timestep 0
fix 1 all nve
%s
fix loop all extrusion %d 1 2 3 0.5 2 4
fix loading all ex_load %d 1 1 1.12 2 prob 0.5 684474 iparam 1 1 jparam 1 1
fix unloading all ex_unload %d 2 0.5 prob 0.5 456456
thermo_style custom step temp epair emol bonds f_loop[1] f_loop[2] f_loading[1] f_loading[2] f_unloading[1] f_unloading[2]
thermo %d
timestep 0.005
run %d
"""


def _columns(stdout):
    m = re.search(r"Step Temp E_pair E_mol Bonds f_loop\[1\] f_loop\[2\] f_loading\[1\] f_loading\[2\] f_unloading\[1\] f_unloading\[2\]\s*\n(.*?)\nLoop time", stdout, re.S)
    assert m, stdout[-2000:]
    return np.array([[float(v) for v in line.split()] for line in m.group(1).strip().splitlines()])


def test_appendix_b_deck_runs_unchanged_through_le_deck(tmp_path):
    """The deck of SURVEY.md Appendix B, verbatim (minimize, hybrid bond style, the three fixes, f_ID[k] columns)."""
    from oracle import refio
    s = systems.chromatin_chain(2000, 40, rho=0.2, seed=11)
    refio.write_data_file(str(tmp_path / "data.le"), s)
    deck = APPENDIX_B % ("minimize 1e-6 1e-8 2000 20000", "fix 2 all langevin 1.0 1.0 1.0 904297", 500, 100, 100, 100, 700)
    (tmp_path / "in.le").write_text(deck)
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    r = subprocess.run([exe, "-in", "in.le"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Minimization stats:" in r.stdout and "Stopping criterion" in r.stdout
    rows = _columns(r.stdout)
    assert rows[0, 0] == 0 and rows[-1, 0] == 700 and len(rows) == 8
    nb0 = 1999 + 40
    assert rows[0, 4] == nb0 and (rows[0, 5:] == 0).all()
    # bookkeeping identities of the reference's counters: bonds = initial + loads - unloads; shifts do not change the count;
    # f_ID[2] are running totals of f_ID[1] over the fix's events
    assert (rows[:, 4] == nb0 + rows[:, 8] - rows[:, 10]).all()
    assert (np.diff(rows[:, 6]) >= 0).all() and (np.diff(rows[:, 8]) >= 0).all() and (np.diff(rows[:, 10]) >= 0).all()
    assert (rows[:, 6] == 0).all()          # f_loop[2]: the reference never accumulates it (fix_extrusion.cpp:139, :1500)
    assert rows[:, 5].max() > 0 and rows[-1, 10] >= 0
    assert 0.5 < rows[-1, 1] < 1.5


def test_le_counter_columns_match_the_reference_run(tmp_path):
    """The same deck in the compiled reference and in le_deck, from a pre-minimised data file and WITHOUT the thermostat
    (the Langevin noise streams differ by design), at a dense event cadence: the printed bonds / f_loop / f_loading /
    f_unloading columns agree row for row while the two trajectories are still the same to ~1e-6."""
    _need_ref()
    from oracle import refio
    s = systems.chromatin_chain(2000, 40, rho=0.2, seed=11)
    e = systems.make_engine(s)
    systems.relax(e, steps=800)
    x, im = e.positions()
    v = e.velocities()
    e.close()
    s2 = dict(s)
    s2["x"], s2["image"], s2["v"] = x, im, v
    refio.write_data_file(str(tmp_path / "data.le"), s2)
    deck = APPENDIX_B % ("", "", 10, 5, 5, 1, 24)
    out, _ = refio.run_reference(deck.splitlines(), workdir=str(tmp_path), harness=False)
    ref = _columns(out)
    (tmp_path / "in.le").write_text(deck)
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    r = subprocess.run([exe, "-in", "in.le"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = _columns(r.stdout)
    assert got.shape == ref.shape
    assert (got[:, 0] == ref[:, 0]).all()
    assert (got[:, 4:] == ref[:, 4:]).all(), "bonds / f_ID columns differ:\n%s\n%s" % (got[:, 4:], ref[:, 4:])
    assert ref[:, 5].max() > 0 and ref[-1, 8] + ref[-1, 10] > 0, "the window must hold real events"
    assert np.abs(got[:, 1:4] - ref[:, 1:4]).max() < 1e-3
