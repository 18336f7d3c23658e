"""BASELINE.json configs[0] through the C++ host front end: the reference's bench/in.chain input script runs UNCHANGED
on le_deck (lammps_le_b200/csrc/le_deck.cpp -> C ABI -> CUDA engine) and reproduces the thermo output of the
reference's own published log (bench/log.6Oct16.chain.fixed.icc.1; fixture tests/golden/bench_chain.npz)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# bench/in.chain of the reference, verbatim (20 input lines)
IN_CHAIN = """# FENE beadspring benchmark

units		lj
atom_style	bond
special_bonds   fene

read_data	data.chain

neighbor	0.4 bin
neigh_modify	every 1 delay 1

bond_style      fene
bond_coeff	1 30.0 1.5 1.0 1.0

pair_style	lj/cut 1.12
pair_modify	shift yes
pair_coeff	1 1 1.0 1.0 1.12

fix		1 all nve
fix		2 all langevin 1.0 1.0 10.0 904297

thermo          100
timestep	0.012

run		100
"""


def write_data_chain(path, z):
    z = {k: z[k] for k in z.files} if hasattr(z, "files") else z      # an NpzFile decompresses on every access
    n = len(z["x"])
    with open(path, "w") as f:
        f.write("LAMMPS data file (tests/golden/bench_chain.npz)\n\n%d atoms\n%d bonds\n\n1 atom types\n1 bond types\n\n" % (n, len(z["bonds"])))
        for k, ax in enumerate("xyz"):
            f.write("%.17g %.17g %slo %shi\n" % (z["boxlo"][k], z["boxhi"][k], ax, ax))
        f.write("\nMasses\n\n1 1\n\nAtoms\n\n")
        for t in range(n):
            f.write("%d %d %d %.17g %.17g %.17g %d %d %d\n" % (t + 1, z["mol"][t], z["type"][t], *z["x"][t], *z["image"][t]))
        f.write("\nVelocities\n\n")
        for t in range(n):
            f.write("%d %.17g %.17g %.17g\n" % (t + 1, *z["v"][t]))
        f.write("\nBonds\n\n")
        for k, (bt, a, b) in enumerate(z["bonds"]):
            f.write("%d %d %d %d\n" % (k + 1, bt, a, b))


def parse_thermo(out):
    m = re.search(r"Step Temp E_pair E_mol TotEng Press \n(.*?)\nLoop time", out, re.S)
    return np.array([[float(v) for v in row.split()] for row in m.group(1).splitlines()])


@pytest.mark.gpu
def test_bench_in_chain_runs_unchanged_and_matches_the_reference_log(tmp_path):
    z = np.load(os.path.join(GOLD, "bench_chain.npz"))
    write_data_chain(tmp_path / "data.chain", z)
    (tmp_path / "in.chain").write_text(IN_CHAIN)
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    assert os.path.exists(exe), "le_deck is not built (make -C lammps_le_b200/csrc)"
    r = subprocess.run([exe, "-in", "in.chain"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    th = parse_thermo(r.stdout)
    ref = z["ref_thermo"]
    assert th.shape == ref.shape and (th[:, 0] == ref[:, 0]).all()
    # step 0 is deterministic: Temp, E_pair, E_mol, TotEng, Press as the reference prints them (8 significant digits);
    # positions are snapped to a 2^-32 box-fraction grid (8e-9 sigma here), which moves the energies by < 1e-6 relative
    rel = np.abs(th[0, 1:] - ref[0, 1:]) / np.abs(ref[0, 1:])
    assert rel.max() < 2e-6, (th[0], ref[0])
    # step 100: a different (counter-based) noise stream -> agreement within thermal fluctuations of a 32,000-bead melt
    assert abs(th[1, 1] - ref[1, 1]) < 0.02           # Temp
    assert abs(th[1, 2] - ref[1, 2]) < 0.02           # E_pair
    assert abs(th[1, 3] - ref[1, 3]) < 0.08           # E_mol
    assert abs(th[1, 5] - ref[1, 5]) < 0.15           # Press
    builds = int(re.search(r"Neighbor list builds = (\d+)", r.stdout).group(1))
    assert abs(builds - int(z["ref_builds"])) <= 6


@pytest.mark.gpu
def test_deck_write_data_roundtrip(tmp_path):
    """write_data -> read_data through the front end: same step-0 thermo from the written file"""
    z = np.load(os.path.join(GOLD, "bench_chain.npz"))
    write_data_chain(tmp_path / "data.chain", z)
    deck = IN_CHAIN.replace("run\t\t100", "run 0\nwrite_data out.data")
    (tmp_path / "in.a").write_text(deck)
    (tmp_path / "in.b").write_text(deck.replace("data.chain", "out.data").replace("write_data out.data", ""))
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    a = subprocess.run([exe, "-in", "in.a"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    b = subprocess.run([exe, "-in", "in.b"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert a.returncode == 0 and b.returncode == 0, a.stderr[-1500:] + b.stderr[-1500:]
    ta, tb = parse_thermo(a.stdout), parse_thermo(b.stdout)
    assert np.abs(ta[0] - tb[0]).max() < 1e-9


@pytest.mark.gpu
def test_deck_dump_custom(tmp_path):
    """dump ID all custom N file id type xu yu zu: frames in the reference's text format on the dump steps"""
    z = np.load(os.path.join(GOLD, "bench_chain.npz"))
    write_data_chain(tmp_path / "data.chain", z)
    deck = IN_CHAIN.replace("run\t\t100", "dump 1 all custom 20 traj.lammpstrj id type xu yu zu\nrun 50")
    (tmp_path / "in.d").write_text(deck)
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    r = subprocess.run([exe, "-in", "in.d"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = (tmp_path / "traj.lammpstrj").read_text().splitlines()
    n = 32000
    per = 9 + n
    assert len(lines) == 3 * per                         # steps 0, 20, 40
    assert [lines[k * per + 1] for k in range(3)] == ["0", "20", "40"]
    assert lines[8] == "ITEM: ATOMS id type xu yu zu"
    first = np.array(lines[9].split(), dtype=float)
    assert first[0] == 1 and np.allclose(first[2:], z["x"][0], atol=1e-5)
    th = parse_thermo(r.stdout)
    assert th[0, 0] == 0 and th[-1, 0] == 50


@pytest.mark.gpu
def test_deck_velocity_create(tmp_path):
    """velocity all create 1.0 seed before `run 0`: the temperature of step 0 is the requested one, as in the reference"""
    z = np.load(os.path.join(GOLD, "bench_chain.npz"))
    write_data_chain(tmp_path / "data.chain", z)
    deck = IN_CHAIN.replace("run\t\t100", "velocity all create 1.25 4928459 dist gaussian\nrun 0")
    (tmp_path / "in.v").write_text(deck)
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    r = subprocess.run([exe, "-in", "in.v"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    th = parse_thermo(r.stdout)
    assert abs(th[0, 1] - 1.25) < 1e-6        # velocities are stored in fp32


@pytest.mark.gpu
def test_deck_thermo_style_custom(tmp_path):
    """thermo_style custom step temp pe ke etotal bonds atoms vol: the reference's column titles and consistent values"""
    z = np.load(os.path.join(GOLD, "bench_chain.npz"))
    write_data_chain(tmp_path / "data.chain", z)
    deck = IN_CHAIN.replace("run\t\t100", "thermo_style custom step temp pe ke etotal bonds atoms vol\nrun 0")
    (tmp_path / "in.t").write_text(deck)
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    r = subprocess.run([exe, "-in", "in.t"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    m = re.search(r"Step Temp PotEng KinEng TotEng Bonds Atoms Volume \n(.*?)\nLoop time", r.stdout, re.S)
    assert m, r.stdout[-1500:]
    row = [float(v) for v in m.group(1).split()]
    ref = z["ref_thermo"][0]                 # Step Temp E_pair E_mol TotEng Press
    assert row[0] == 0 and abs(row[1] - ref[1]) < 1e-5
    assert abs(row[2] - (ref[2] + ref[3])) < 1e-4 and abs(row[2] + row[3] - row[4]) < 1e-6
    assert row[5] == 31680 and row[6] == 32000
    assert abs(row[7] - 33.5919 ** 3) < 1.0


@pytest.mark.gpu
def test_deck_dump_local_bonds(tmp_path):
    """compute b all property/local batom1 batom2 btype + dump ... local: one entry per bond in the reference's format"""
    z = np.load(os.path.join(GOLD, "bench_chain.npz"))
    write_data_chain(tmp_path / "data.chain", z)
    deck = IN_CHAIN.replace("run\t\t100", "compute b all property/local batom1 batom2 btype\n"
                            "dump d all local 10 bonds.dump index c_b[1] c_b[2] c_b[3]\nrun 10")
    (tmp_path / "in.l").write_text(deck)
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    r = subprocess.run([exe, "-in", "in.l"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = (tmp_path / "bonds.dump").read_text().splitlines()
    per = 9 + 31680
    assert len(lines) == 2 * per and lines[2] == "ITEM: NUMBER OF ENTRIES" and lines[3] == "31680"
    assert lines[8] == "ITEM: ENTRIES index c_b[1] c_b[2] c_b[3] "
    first = [float(v) for v in lines[9].split()]
    assert first == [1.0, 1.0, 2.0, 1.0]


@pytest.mark.gpu
def test_restart_files_through_le_deck(tmp_path):
    """read_restart of a file the compiled reference wrote (tests/golden/ref_restart_small.bin: chain with extruder bonds): the step-0
    thermo line of `run 0` equals the energies of the same state uploaded through the Python binding; after a run, write_restart ->
    read_restart in a second le_deck process continues with identical thermo (the bond tables incl. extruder bonds survive)."""
    import shutil
    from lammps_le_b200 import restart as RS
    shutil.copy(os.path.join(GOLD, "ref_restart_small.bin"), tmp_path / "ref.restart")
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    common = "neighbor 0.4 bin\nneigh_modify every 1 delay 0 check yes\nthermo_style custom step temp epair emol bonds\nthermo 20\n"
    (tmp_path / "in.a").write_text("read_restart ref.restart\n" + common + "fix 1 all nve\nrun 40\nwrite_restart mid.restart\nrun 20\n")
    (tmp_path / "in.b").write_text("read_restart mid.restart\n" + common + "fix 1 all nve\nrun 20\n")
    ra = subprocess.run([exe, "-in", "in.a"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert ra.returncode == 0, ra.stdout[-2000:] + ra.stderr[-2000:]
    rb = subprocess.run([exe, "-in", "in.b"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert rb.returncode == 0, rb.stdout[-2000:] + rb.stderr[-2000:]
    rows = lambda out: [np.array([[float(v) for v in r.split()] for r in body.splitlines()])
                        for body in re.findall(r"Step Temp E_pair E_mol Bonds \n(.*?)\nLoop time", out, re.S)]
    a, b = rows(ra.stdout), rows(rb.stdout)
    info, atoms = RS.read_restart(os.path.join(GOLD, "ref_restart_small.bin"))
    assert a[0][0, 0] == info["ntimestep"] and a[0][0, 4] == info["nbonds"]
    e = RS.engine_from_restart(info, atoms)
    _, th = e.compute_forces()
    e.close()
    assert abs(a[0][0, 2] - th["epair"]) < 1e-6 * max(1.0, abs(th["epair"])) and abs(a[0][0, 3] - th["emol"]) < 1e-6 * abs(th["emol"])
    # the second process starts where the first one wrote the file and prints what the first one printed from there on
    assert np.array_equal(a[1][0], b[0][0]) and np.allclose(a[1], b[0], rtol=2e-6, atol=1e-9), (a[1], b[0])


@pytest.mark.gpu
def test_angle_cosine_deck_matches_the_reference(tmp_path):
    """atom_style angle + angle_style cosine (chain stiffness): the same input script and data file through
    le_deck and through the compiled reference -- thermo columns incl. E_angle at step 0 to print precision, over a short NVE run to 2e-5.
    (comm_modify cutoff: the reference needs the end atoms of every angle among its ghosts, 2 bonds = up to 3 sigma away; with the
    default 1.52 it silently tallies a few boundary angles against the wrong image.)"""
    from oracle import refio
    if not refio.have_reference():
        pytest.skip("oracle/_ref not present on this box")
    from lammps_le_b200 import systems
    n = 2000
    s = systems.chromatin_chain(n, 20, rho=0.2, seed=17, extruder_bond=systems.EXTRUDER_FENE)
    e = systems.make_engine(s, velocities=systems.maxwell_velocities(n, 1.0, np.ones(n), 6), dt=0.005)
    systems.relax(e, steps=500)
    x, im = e.positions(); v = e.velocities(); e.close()
    img = np.stack([(im & 1023) - 512, ((im >> 10) & 1023) - 512, ((im >> 20) & 1023) - 512], axis=1)
    bt, b1, b2 = s["bonds"]
    lo, hi = s["box"]
    with open(tmp_path / "data.angle", "w") as f:
        f.write("chain with stiffness\n\n%d atoms\n%d bonds\n%d angles\n\n4 atom types\n2 bond types\n2 angle types\n\n" % (n, len(bt), n - 2))
        for k, ax in enumerate("xyz"):
            f.write("%.17g %.17g %slo %shi\n" % (lo[k], hi[k], ax, ax))
        f.write("\nMasses\n\n1 1\n2 1\n3 1\n4 1\n\nAtoms # angle\n\n")
        f.write("\n".join("%d 1 %d %.17g %.17g %.17g %d %d %d" % (t + 1, s["types"][t], *x[t], *img[t]) for t in range(n)))
        f.write("\n\nVelocities\n\n" + "\n".join("%d %.17g %.17g %.17g" % (t + 1, *v[t]) for t in range(n)))
        f.write("\n\nBonds\n\n" + "\n".join("%d %d %d %d" % (k + 1, bt[k], b1[k], b2[k]) for k in range(len(bt))))
        f.write("\n\nAngles\n\n" + "\n".join("%d %d %d %d %d" % (c - 1, 1 + c % 2, c - 1, c, c + 1) for c in range(2, n)) + "\n")
    deck = """units lj
atom_style angle
newton on off
special_bonds fene
atom_modify sort 0 0
read_data data.angle
neighbor 0.4 bin
neigh_modify every 1 delay 0 check yes
comm_modify cutoff 5.0
bond_style fene
bond_coeff 1 30.0 1.5 1.0 1.0
bond_coeff 2 10.0 4.0 1.0 1.0
angle_style cosine
angle_coeff 1 2.0
angle_coeff 2 0.75
pair_style lj/cut 1.12246
pair_modify shift yes
pair_coeff * * 1.0 1.0 1.12246
fix 1 all nve
thermo_style custom step temp epair emol eangle etotal press
thermo 10
timestep 0.005
run 30
"""
    (tmp_path / "in.angle").write_text(deck)
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    r = subprocess.run([exe, "-in", "in.angle"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out, _ = refio.run_reference(deck.splitlines(), workdir=str(tmp_path), harness=False)
    rows = lambda o: np.array([[float(q) for q in ln.split()] for ln in re.search(r"Step Temp E_pair E_mol E_angle TotEng Press \n(.*?)\nLoop time", o, re.S).group(1).splitlines()])
    a, b = rows(r.stdout), rows(out)
    assert a.shape == b.shape == (4, 7) and (a[:, 0] == b[:, 0]).all()
    assert b[0, 4] > 0.05, "the deck must have angle energy"
    rel = np.abs(a[:, 1:] - b[:, 1:]) / np.maximum(np.abs(b[:, 1:]), 1e-3)
    assert rel[0].max() < 2e-6 and rel.max() < 2e-5, (a, b)
