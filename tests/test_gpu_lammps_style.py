"""The LAMMPS-side binding, compiled: `run_style le/b200` (lammps_le_b200/lammps_style/verlet_le_b200.{h,cpp}) built against the
reference's own sources into oracle/_ref/b200/lmp_b200 (oracle/build_ref.py, test-only).  The reference's Input::file parses the
deck, its ReadData / Special / Force / Modify objects hold the state, and the timestep loop runs in libleb200.so through the C ABI.
  * bench/in.chain + one added line reproduces the step-0 thermo line of the reference's published log;
  * a short NVE run of the same deck gives the stock `run_style verlet` thermo columns (same binary, same deck) to 2e-5;
  * a chromatin deck with the three USER-LE fixes runs, and Thermo's own `bonds` / f_ID[k] columns follow the engine."""
import os
import re
import subprocess

import numpy as np
import pytest

from tests.test_gpu_deck import GOLD, IN_CHAIN, write_data_chain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "b200", "lmp_b200")
pytestmark = pytest.mark.gpu


def run_lmp(deck, cwd):
    (cwd / "in.deck").write_text(deck)
    r = subprocess.run([EXE, "-in", "in.deck", "-echo", "none"], cwd=cwd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return r.stdout


def thermo_rows(out):
    blocks = re.findall(r"\n(Step [^\n]*)\n(.*?)\nLoop time", out, re.S)
    return [(h.split(), np.array([[float(v) for v in row.split()] for row in body.splitlines()])) for h, body in blocks]


def need_exe():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/b200/lmp_b200 not present on this box")


def test_in_chain_through_the_reference_input_parser(tmp_path):
    need_exe()
    z = np.load(os.path.join(GOLD, "bench_chain.npz"))
    write_data_chain(tmp_path / "data.chain", z)
    out = run_lmp(IN_CHAIN.replace("run\t\t100", "run_style le/b200\nrun\t\t100"), tmp_path)
    assert "Setting up le/b200 run" in out
    (hdr, th), = thermo_rows(out)
    ref = z["ref_thermo"]
    assert hdr == ["Step", "Temp", "E_pair", "E_mol", "TotEng", "Press"] and th.shape == ref.shape
    rel = np.abs(th[0, 1:] - ref[0, 1:]) / np.abs(ref[0, 1:])
    assert rel.max() < 2e-6, (th[0], ref[0])                       # the published log's step-0 line
    assert abs(th[1, 1] - ref[1, 1]) < 0.05 and abs(th[1, 3] - ref[1, 3]) < 0.5   # step 100: another noise stream, same physics


def test_nve_run_matches_run_style_verlet_of_the_same_binary(tmp_path):
    need_exe()
    z = np.load(os.path.join(GOLD, "bench_chain.npz"))
    write_data_chain(tmp_path / "data.chain", z)
    base = IN_CHAIN.replace("fix\t\t2 all langevin 1.0 1.0 10.0 904297\n", "").replace("thermo          100", "thermo 10").replace("timestep\t0.012", "timestep 0.005")
    a = thermo_rows(run_lmp(base.replace("run\t\t100", "run 30"), tmp_path))[0][1]
    b = thermo_rows(run_lmp(base.replace("run\t\t100", "run_style le/b200\nrun 30"), tmp_path))[0][1]
    assert a.shape == b.shape == (4, 6) and (a[:, 0] == b[:, 0]).all()
    rel = np.abs(a[:, 1:] - b[:, 1:]) / np.maximum(np.abs(a[:, 1:]), 1e-3)
    assert rel.max() < 2e-5, (a, b)


def test_chromatin_deck_with_user_le_fixes(tmp_path):
    need_exe()
    from lammps_le_b200 import systems
    from oracle import refio
    s = systems.chromatin_chain(4000, 40, rho=0.2, seed=7, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    refio.write_data_file(str(tmp_path / "data.le"), s)
    deck = "\n".join(refio.deck_header(s, "data.le", sort=True) + [
        "velocity all create 1.0 4711", "fix 1 all nve/limit 0.05", "fix 2 all langevin 1.0 1.0 1.0 904297", "timestep 0.005",
        "run_style le/b200", "thermo 100", "run 300", "unfix 1", "fix 1 all nve",
        "fix loop all extrusion 200 1 2 3 0.5 2 4", "fix loading all ex_load 50 1 1 1.12 2 prob 0.02 684474 iparam 1 1 jparam 1 1",
        "fix unloading all ex_unload 50 2 0.5 prob 0.1 456456",
        "thermo_style custom step temp epair emol bonds f_loop[1] f_loading[1] f_loading[2] f_unloading[2]", "thermo 50", "run 400"]) + "\n"
    out = run_lmp(deck, tmp_path)
    hdr, th = thermo_rows(out)[-1]
    assert hdr[:5] == ["Step", "Temp", "E_pair", "E_mol", "Bonds"]
    bonds = th[:, 4]
    nb0 = len(s["bonds"][0])
    # loaded minus unloaded extruders = change of the bond count, row by row (Thermo's own `bonds` keyword and the fixes' vectors)
    assert (bonds - nb0 == th[:, 7] - th[:, 8] + (bonds[0] - nb0 - th[0, 7] + th[0, 8])).all()
    assert th[-1, 7] > 0 and np.isfinite(th).all() and abs(th[-1, 1] - 1.0) < 0.2


def test_angle_cosine_deck_through_the_binding(tmp_path):
    """atom_style angle + angle_style cosine through run_style le/b200: E_angle of the reference's own Thermo follows the engine and
    equals stock run_style verlet of the same binary at step 0 (2e-6) and over a short NVE run (2e-5)"""
    need_exe()
    from lammps_le_b200 import systems
    n = 1500
    s = systems.chromatin_chain(n, 0, rho=0.2, seed=23)
    x, im = s["x"], s["image"]
    img = np.stack([(im & 1023) - 512, ((im >> 10) & 1023) - 512, ((im >> 20) & 1023) - 512], axis=1)
    bt, b1, b2 = s["bonds"]
    lo, hi = s["box"]
    with open(tmp_path / "data.angle", "w") as f:
        f.write("chain with stiffness\n\n%d atoms\n%d bonds\n%d angles\n\n4 atom types\n2 bond types\n1 angle types\n\n" % (n, len(bt), n - 2))
        for k, ax in enumerate("xyz"):
            f.write("%.17g %.17g %slo %shi\n" % (lo[k], hi[k], ax, ax))
        f.write("\nMasses\n\n1 1\n2 1\n3 1\n4 1\n\nAtoms # angle\n\n")
        f.write("\n".join("%d 1 %d %.17g %.17g %.17g %d %d %d" % (t + 1, s["types"][t], *x[t], *img[t]) for t in range(n)))
        f.write("\n\nBonds\n\n" + "\n".join("%d %d %d %d" % (k + 1, bt[k], b1[k], b2[k]) for k in range(len(bt))))
        f.write("\n\nAngles\n\n" + "\n".join("%d 1 %d %d %d" % (c - 1, c - 1, c, c + 1) for c in range(2, n)) + "\n")
    base = """units lj
atom_style angle
newton on off
special_bonds fene
atom_modify sort 0 0
read_data data.angle
neighbor 0.4 bin
neigh_modify every 1 delay 0 check yes
comm_modify cutoff 5.0
bond_style fene
bond_coeff * 30.0 1.5 1.0 1.0
angle_style cosine
angle_coeff 1 1.5
pair_style lj/cut 1.12246
pair_modify shift yes
pair_coeff * * 1.0 1.0 1.12246
velocity all create 1.0 4711
fix 1 all nve/limit 0.02
thermo_style custom step temp epair emol eangle etotal
thermo 10
timestep 0.002
%s
run 30
"""
    a = thermo_rows(run_lmp(base % "", tmp_path))[0][1]
    b = thermo_rows(run_lmp(base % "run_style le/b200", tmp_path))[0][1]
    assert a.shape == b.shape == (4, 6) and a[0, 4] > 0.05
    rel = np.abs(a[:, 1:] - b[:, 1:]) / np.maximum(np.abs(a[:, 1:]), 1e-3)
    assert rel[0].max() < 2e-6 and rel.max() < 5e-5, (a, b)


def test_mc_fixes_through_the_binding_match_run_style_verlet(tmp_path):
    """fix bond/create + fix bond/break (src/MC, the ancestors of ex_load / ex_unload) in a deck the reference's Input::file parses:
    `run_style le/b200` against stock `run_style verlet` of the SAME binary -- Thermo's `bonds` and the four f_ID[k] columns row for row
    (NVE, grid-snapped start: both integrate the same doubles)"""
    need_exe()
    from lammps_le_b200 import systems
    from oracle import refio
    n = 3000
    s = systems.chromatin_chain(n, 90, rho=0.2, seed=21, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    v = systems.maxwell_velocities(n, 1.0, np.ones(n), 5)
    e = systems.make_engine(s, velocities=v, dt=0.005)
    x, im = e.positions()
    e.close()
    s2 = dict(s); s2["x"], s2["image"], s2["v"] = x, im, v
    refio.write_data_file(str(tmp_path / "data.le"), s2)
    base = "\n".join(refio.deck_header(s2, "data.le", sort=False) + [
        "fix 1 all nve", "fix cr all bond/create 10 1 1 1.05 2 prob 0.5 684474 iparam 2 4 jparam 2 4", "fix br all bond/break 10 2 1.2 prob 0.5 456456",
        "timestep 0.005", "%s", "thermo_style custom step bonds f_cr[1] f_cr[2] f_br[1] f_br[2]", "thermo 10", "run 40"]) + "\n"
    a = thermo_rows(run_lmp(base % "", tmp_path))[-1][1]
    b = thermo_rows(run_lmp(base % "run_style le/b200", tmp_path))[-1][1]
    assert a.shape == b.shape == (5, 6), (a, b)
    assert a[-1, 3] > 100 and a[-1, 5] > 5, "the run must create and break bonds"
    assert np.array_equal(a, b), (a, b)
