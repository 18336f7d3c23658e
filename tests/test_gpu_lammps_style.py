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
