"""GPU tests of the timestep itself: integration, thermostat, reneighboring, determinism, error paths."""
import numpy as np
import pytest

from lammps_le_b200 import systems
from lammps_le_b200.engine import LeError

pytestmark = pytest.mark.gpu


def relaxed_chain(n=5000, next_=50, seed=9, steps=1500):
    s = systems.chromatin_chain(n, next_, rho=0.2, seed=seed)
    e = systems.make_engine(s, velocities=systems.maxwell_velocities(n, 1.0, np.ones(n), seed))
    systems.relax(e, steps=steps)
    return s, e


def test_nve_conserves_energy():
    s, e = relaxed_chain()
    e.fix_langevin(1.0, 1.0, 1.0, 1)          # define then switch the thermostat off by unfixing: plain NVE
    x, im = e.positions()
    v = e.velocities()
    topo = e.topology()
    e.close()
    e = systems.make_engine(s, velocities=v, dt=0.005)
    e.set_positions(x, im)
    e.upload_topology(topo["num_bond"], topo["bond_type"], topo["bond_atom"], topo["nspecial"], topo["special"])
    e.fix_nve(True)
    e.thermo_every(100)
    e.run(2000)
    th = e.thermo()
    et = np.array([t["etotal"] for t in th])
    ke = np.array([t["ke"] for t in th]) / e.natoms
    assert len(th) == 21
    drift = np.abs(et - et[0]).max()
    assert drift < 2e-3 * ke.mean(), "NVE energy drift %.3g per atom (KE/atom %.3g)" % (drift, ke.mean())
    st = e.stats()
    assert st["neigh_builds"] > 5 and st["dangerous_builds"] == 0
    e.close()


def test_langevin_holds_temperature():
    s, e = relaxed_chain(n=20000, next_=200, steps=1000)
    e.fix_langevin(1.0, 1.0, 1.0, 904297)
    e.thermo_every(50)
    e.run(3000)
    t = np.array([r["temp"] for r in e.thermo()][-40:])
    assert abs(t.mean() - 1.0) < 0.03, "mean T %.4f" % t.mean()
    e.close()


def test_run_is_deterministic():
    out = []
    for _ in range(2):
        s, e = relaxed_chain(n=3000, next_=30, steps=300)
        e.fix_langevin(1.0, 1.0, 1.0, 77)
        e.run(400)
        out.append((e.positions()[0], e.velocities()))
        e.close()
    assert (out[0][0] == out[1][0]).all() and (out[0][1] == out[1][1]).all()


def test_bad_fene_bond_aborts_the_run():
    s = systems.chromatin_chain(500, 0, rho=0.2, seed=1)
    s["x"] = s["x"].copy()
    s["x"][250] += 2.9          # bond 250-251 and 251-252 now > 2 R0
    e = systems.make_engine(s)
    e.fix_nve(True)
    with pytest.raises(LeError) as ei:
        e.run(1)
    assert "Bad FENE bond" in str(ei.value)
    e.close()


def test_short_nve_trajectory_matches_reference():
    from oracle import refio
    if not refio.have_reference():
        pytest.skip("oracle/_ref not present on this box")
    import os
    import tempfile
    from oracle.make_golden import force_case
    s = systems.chromatin_chain(3000, 30, rho=0.2, seed=17)
    rec0 = force_case(s, velocities=True)
    # reference: 20 NVE steps from the snapped state
    wd = tempfile.mkdtemp(prefix="le_traj_")
    s2 = dict(s)
    s2["x"], s2["image"], s2["v"] = rec0["x"], rec0["image"], rec0["v"]
    refio.write_data_file(os.path.join(wd, "data.le"), s2)
    deck = refio.deck_header(s2, "data.le") + ["fix 1 all nve", "thermo_style custom step temp epair emol etotal press",
                                                 "thermo 20", "timestep 0.005", "run 20"]
    final = os.path.join(wd, "final.bin")
    out, _ = refio.run_reference(deck, workdir=wd, final=final)
    ref = refio.read_records(final)[0]
    rows = refio.parse_thermo(out)
    from tests import lehelpers as H
    e = H.engine_from_record(rec0, s["bond_coeffs"], positions="x")
    e.set_timestep(0.005)
    e.fix_nve(True)
    e.run(20)
    x, im = e.positions()
    L = rec0["boxhi"] - rec0["boxlo"]
    d = x - ref["x"]
    d -= L * np.rint(d / L)
    assert np.abs(d).max() < 2e-5, "max position deviation after 20 steps %.3g" % np.abs(d).max()
    assert np.abs(e.velocities() - ref["v"]).max() < 2e-4
    th = e.thermo(-1)
    assert abs(th["etotal"] - rows[-1]["TotEng"]) < 2e-5 * abs(rows[-1]["TotEng"])
    assert abs(th["temp"] - rows[-1]["Temp"]) < 1e-4
    e.close()


@pytest.mark.gpu
def test_owned_bulk_exchange_roundtrip():
    """le_download_owned / le_upload_owned (device-side double <-> fixed point) against the tag-order calls"""
    from lammps_le_b200 import systems
    s = systems.chromatin_chain(5000, 50, rho=0.2, seed=3)
    v = systems.maxwell_velocities(5000, 1.0, np.ones(5000), 2)
    e = systems.make_engine(s, velocities=v)
    e.fix_nve(True); e.fix_langevin(1.0, 1.0, 1.0, 77)
    e.run(50)
    x, im = e.positions(); vv = e.velocities()
    bufs = e.owned_buffers(pinned=False)
    n = e.download_owned(bufs)
    tag, xb, imb, vb = bufs
    assert n == 5000 and sorted(tag[:n].tolist()) == list(range(1, 5001))
    assert (xb[:n] == x[tag[:n] - 1]).all() and (imb[:n] == im[tag[:n] - 1]).all() and (vb[:n] == vv[tag[:n] - 1]).all()
    # upload shifted coordinates (some leave the box) and compare with le_set_positions on a twin engine
    L = s["box"][1][0]
    xb[:n] += 0.37
    e2 = systems.make_engine(s, velocities=v)
    e2.fix_nve(True); e2.fix_langevin(1.0, 1.0, 1.0, 77)
    e2.run(50)
    xs = x + 0.37
    e2.set_positions(xs, im)
    e.upload_owned(n, bufs)
    xa, ia = e.positions(); xc, ic = e2.positions()
    assert (xa == xc).all() and (ia == ic).all()
    assert (xa >= 0).all() and (xa < L).all()
    e.close(); e2.close()


@pytest.mark.gpu
def test_two_live_engines_with_different_boxes():
    """the constant parameter block is one symbol per process and device: every entry point must refresh it, so a second
    context with another box cannot leak its mapping into the first one's downloads"""
    from lammps_le_b200 import systems
    sa = systems.chromatin_chain(3000, 30, rho=0.2, seed=3)
    sb = systems.chromatin_chain(24000, 240, rho=0.05, seed=4)
    assert abs(sa["box"][1][0] - sb["box"][1][0]) > 1.0
    a = systems.make_engine(sa, velocities=np.zeros((3000, 3)))
    xa0, ia0 = a.positions()
    b = systems.make_engine(sb, velocities=np.zeros((24000, 3)))
    b.fix_nve(True); b.run(5)                         # b's parameters are now the ones in constant memory
    bufs = a.owned_buffers(pinned=False)
    n = a.download_owned(bufs)
    tag, xb, imb, vb = bufs
    assert n == 3000 and (xb[:n] == xa0[tag[:n] - 1]).all()
    xa1, _ = a.positions()
    assert (xa1 == xa0).all()
    f, _ = a.compute_forces()
    b.run(3)
    f2, _ = a.compute_forces()
    assert np.array_equal(f, f2)
    a.close(); b.close()


@pytest.mark.gpu
def test_device_observables_match_host():
    """le_observables (Rg, contact counts, loop-size histogram on the GPU) against numpy on downloaded state"""
    s = systems.chromatin_chain(6000, 80, rho=0.2, seed=9, extruder_bond=systems.EXTRUDER_FENE)
    v = systems.maxwell_velocities(6000, 1.0, np.ones(6000), 5)
    e = systems.make_engine(s, velocities=v)
    systems.relax(e, steps=1500)
    e.fix_langevin(1.0, 1.0, 1.0, 99)
    e.fix_extrusion(200, 1, 2, 3, 0.5, 2, 4, 12345)
    e.run(1300)
    s_list = [2, 3, 5, 8, 16, 50]
    o = e.observables(s_list, rc=1.5, btype=2, nbins=16, bin_width=4)
    xu, _ = e.positions(unwrap=True)
    rg = np.sqrt(((xu - xu.mean(0)) ** 2).sum(1).mean())
    assert abs(o["rg"] - rg) < 1e-9 * rg and abs(o["rg"] - e.rg()) < 1e-9 * rg
    x, _ = e.positions()
    L = s["box"][1][0]
    for k, sp in enumerate(s_list):
        dd = x[sp:] - x[:-sp]
        dd -= L * np.rint(dd / L)
        r2 = (dd ** 2).sum(1)
        cnt = int((r2 < 1.5 ** 2).sum())
        near_cut = int((np.abs(np.sqrt(r2) - 1.5) < 1e-5).sum())        # fp32 distance on the device
        assert abs(int(round(o["ps"][k] * (6000 - sp))) - cnt) <= near_cut
    topo = e.topology()
    nb, bt, ba = topo["num_bond"], topo["bond_type"], topo["bond_atom"]
    mask = (np.arange(bt.shape[1])[None, :] < nb[:, None]) & (bt == 2) & (ba > (np.arange(6000) + 1)[:, None])
    ii, mm = np.nonzero(mask)
    sizes = ba[ii, mm] - (ii + 1)
    hist = np.bincount(np.minimum(sizes // 4, 15), minlength=16)
    assert (o["loop_hist"] == hist).all() and o["nloops"] == len(sizes)
    e.close()


@pytest.mark.gpu
def test_fix_bond_break_matches_the_reference():
    """fix bond/break (src/MC/fix_bond_break.cpp, the ancestor of fix ex_unload: same body, events on multiples of N): bond counts and
    the fix's counters after every event equal a run of the compiled reference from the same state"""
    import os
    import re
    import tempfile
    from oracle import refio
    from lammps_le_b200 import systems
    if not refio.have_reference():
        pytest.skip("oracle/_ref not present on this box")
    n = 3000
    s = systems.chromatin_chain(n, 90, rho=0.2, seed=21, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    v = systems.maxwell_velocities(n, 1.0, np.ones(n), 5)
    e = systems.make_engine(s, velocities=v, dt=0.005)
    x, im = e.positions()                                   # the engine's grid-snapped start, handed to the reference
    s2 = dict(s); s2["x"], s2["image"], s2["v"] = x, im, v
    wd = tempfile.mkdtemp(prefix="le_bb_")
    refio.write_data_file(os.path.join(wd, "data.le"), s2)
    deck = refio.deck_header(s2, "data.le", sort=False) + ["fix 1 all nve", "fix br all bond/break 10 2 1.2 prob 0.5 456456", "timestep 0.005",
                                                          "thermo_style custom step bonds f_br[1] f_br[2]", "thermo 10", "run 40"]
    out, _ = refio.run_reference(deck, workdir=wd, harness=False)
    ref = np.array([[float(q) for q in r.split()] for r in re.search(r"Step Bonds f_br\[1\] f_br\[2\] \n(.*?)\nLoop time", out, re.S).group(1).splitlines()])
    e.fix_nve(True)
    e.fix_bond_break(10, 2, 1.2, 0.5, 456456)
    e.thermo_every(10)
    e.run(40)
    th = e.thermo()
    got = np.array([[t["step"], t["nbonds"], t["le_f1"][1], t["le_f2"][1]] for t in th], dtype=float)
    e.close()
    assert ref.shape == got.shape == (5, 4), (ref, got)
    assert ref[-1, 3] > 5, "the run must break bonds"
    assert np.array_equal(ref, got), (ref, got)


@pytest.mark.gpu
@pytest.mark.parametrize("through", ["abi", "le_deck"])
def test_fix_bond_create_matches_the_reference(through, tmp_path):
    """fix bond/create (src/MC/fix_bond_create.cpp, the ancestor of fix ex_load: the closest eligible listed neighbor within Rmin, no
    loop-extrusion rules, events on multiples of N): bond counts and the fix's counters after every event equal a run of the compiled
    reference from the same state.  `abi`: prob 0.5 (the Marsaglia stream), two bonds per bead, then the bead changes type;
    `le_deck`: the reference's own input lines through the C++ front end, every pair within reach bonds (fraction 1)."""
    import os
    import re
    import subprocess
    from oracle import refio
    from lammps_le_b200 import systems
    if not refio.have_reference():
        pytest.skip("oracle/_ref not present on this box")
    n = 3000
    s = systems.chromatin_chain(n, 90, rho=0.2, seed=21, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    v = systems.maxwell_velocities(n, 1.0, np.ones(n), 5)
    e = systems.make_engine(s, velocities=v, dt=0.005)
    x, im = e.positions()                                   # the engine's grid-snapped start, handed to the reference
    s2 = dict(s); s2["x"], s2["image"], s2["v"] = x, im, v
    wd = str(tmp_path)
    refio.write_data_file(os.path.join(wd, "data.le"), s2)
    fixline = ("fix cr all bond/create 10 1 1 1.05 2 prob 0.5 456456 iparam 2 4 jparam 2 4" if through == "abi"
               else "fix cr all bond/create 5 1 1 1.0 2")
    deck = refio.deck_header(s2, "data.le", sort=False) + ["fix 1 all nve", fixline, "timestep 0.005",
                                                          "thermo_style custom step bonds f_cr[1] f_cr[2]", "thermo 10", "run 40"]
    out, _ = refio.run_reference(deck, workdir=wd, harness=False)
    table = lambda o: np.array([[float(q) for q in r.split()] for r in re.search(r"Step Bonds f_cr\[1\] f_cr\[2\] \n(.*?)\nLoop time", o, re.S).group(1).splitlines()])
    ref = table(out)
    if through == "abi":
        e.fix_nve(True)
        e.fix_bond_create(10, 1, 1, 1.05, 2, 0.5, 456456, (2, 4), (2, 4))
        e.thermo_every(10)
        e.run(40)
        got = np.array([[t["step"], t["nbonds"], t["le_f1"][2], t["le_f2"][2]] for t in e.thermo()], dtype=float)
        types = e.types() if hasattr(e, "types") else None
        e.close()
        if types is not None:
            assert (types == 4).sum() > (s["types"] == 4).sum(), "beads with two created bonds must have changed type"
    else:
        e.close()
        with open(os.path.join(wd, "in.create"), "w") as f:
            f.write("\n".join(deck) + "\n")
        r = subprocess.run([os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lammps_le_b200", "le_deck"), "-in", "in.create"],
                           cwd=wd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        got = table(r.stdout)
    assert ref.shape == got.shape == (5, 4), (ref, got)
    assert ref[-1, 3] > 100, "the run must create bonds"
    assert np.array_equal(ref, got), (ref, got)


@pytest.mark.gpu
def test_langevin_ramp_spans_the_segments_of_one_run():
    """a run cut into segments (le_deck does that at dump steps) must not restart fix langevin's Tstart -> Tstop
    ramp in every segment: le_set_run_span (`run N start S stop E`, src/run.cpp:90-120).  With the span the segmented run stays
    within 1e-3 of the uncut one (what remains is Verlet::setup re-evaluating the drag with the updated velocity at every
    segment start, as the reference does); without it the temperature is a sawtooth and the trajectories part."""
    from lammps_le_b200 import systems
    n = 4000
    s = systems.chromatin_chain(n, 40, rho=0.2, seed=8)
    e = systems.make_engine(s, velocities=systems.maxwell_velocities(n, 1.0, np.ones(n), 2), dt=0.005)
    systems.relax(e, steps=400)
    x, im = e.positions(); v = e.velocities(); e.close()
    s = dict(s); s["x"], s["image"] = x, im

    def traj(segments, span):
        e = systems.make_engine(s, velocities=v, dt=0.005)
        e.fix_nve(True)
        e.fix_langevin(0.5, 1.5, 1.0, 4242)
        if span:
            e.set_run_span(0, sum(segments))
        for k in segments:
            e.run(k)
        out = e.positions()[0]
        e.close()
        return out

    L = s["box"][1][0]
    dist = lambda a, b: np.abs(((a - b) + L / 2) % L - L / 2).max()
    whole, cut, saw = traj([120], False), traj([40, 40, 40], True), traj([40, 40, 40], False)
    assert dist(whole, cut) < 2e-3, dist(whole, cut)
    assert dist(whole, saw) > 20 * dist(whole, cut), (dist(whole, saw), dist(whole, cut))


@pytest.mark.gpu
def test_angle_cosine_matches_the_oracle_and_conserves_energy():
    """angle_style cosine (src/MOLECULE/angle_cosine.cpp:47-140): forces / energies of a relaxed chain with a
    stiffness term against the oracle restatement (itself pinned on the reference's angle-cosine.yaml), then NVE energy conservation
    with the angle energy in E_mol"""
    from oracle import restate as R
    from lammps_le_b200 import systems
    n = 3000
    s = systems.chromatin_chain(n, 30, rho=0.2, seed=13)
    e = systems.make_engine(s, velocities=systems.maxwell_velocities(n, 1.0, np.ones(n), 4), dt=0.005)
    systems.relax(e, steps=500)
    a2 = np.arange(2, n, dtype=np.int32)                      # every interior bead is the centre of one angle
    ty = np.where(a2 % 2 == 0, 1, 2).astype(np.int32)
    e.set_angle_types(2)
    e.set_angle(1, "cosine", (3.0,)); e.set_angle(2, "cosine", (1.5,))
    e.upload_angles(ty, a2 - 1, a2, a2 + 1)
    f, th = e.compute_forces()
    fp = e.compute_forces_plain()
    assert np.array_equal(f, fp)
    x, _ = e.positions()
    L = np.asarray(s["box"][1]) - np.asarray(s["box"][0])
    topo = e.topology()
    rows = R.half_neighbor_list(x, np.asarray(s["box"][0]), np.asarray(s["box"][1]), 1.12246 + 0.4, topo["nspecial"], topo["special"])
    pi = np.array([t for t in range(n) for _ in rows[t]], dtype=int)
    pj = np.array([(w & R.NEIGHMASK) - 1 for t in range(n) for w in rows[t]], dtype=int)
    fo, evdwl, _ = R.pair_lj_cut(x, L, pi, pj, np.zeros(len(pi), int), R.lj_coeffs(1.0, 1.0, 1.12246, True))
    b1, b2, bt = R.unique_bonds(topo["num_bond"], topo["bond_type"], topo["bond_atom"])
    fb, ebond, _, _ = R.bond_forces(x, L, b1, b2, bt, s["bond_coeffs"])
    fa, eang, _ = R.angle_cosine(x, L, a2 - 2, a2 - 1, a2, ty, {1: 3.0, 2: 1.5})
    ft = fo + fb + fa
    mag = np.sqrt((ft ** 2).sum(1))
    rel = (np.sqrt(((f - ft) ** 2).sum(1)) / np.maximum(mag, np.sqrt((mag ** 2).mean()))).max()
    assert rel <= 1e-5, "max per-atom relative force error %.3g" % rel
    assert abs(th["eangle"] * n - eang) <= 1e-9 * abs(eang) and abs(th["emol"] * n - (ebond + eang)) <= 1e-9 * abs(ebond + eang)
    assert abs(th["epair"] * n - evdwl) <= 1e-6 * max(abs(evdwl), 1.0)
    # let the chain settle with its new stiffness under the thermostat, then check NVE conservation with the angle energy in E_mol
    e.fix_nve_limit(0.05); e.fix_langevin(1.0, 1.0, 1.0, 99)
    e.run(600)
    e2 = systems.make_engine(dict(s, x=e.positions()[0], image=e.positions()[1]), velocities=e.velocities(), dt=0.005)
    e2.set_angle_types(2)
    e2.set_angle(1, "cosine", (3.0,)); e2.set_angle(2, "cosine", (1.5,))
    e2.upload_angles(ty, a2 - 1, a2, a2 + 1)
    e2.fix_nve(True)
    e2.thermo_every(100)
    e2.run(1000)
    th = e2.thermo()
    et = np.array([t["etotal"] for t in th])
    assert sum(t["fene_warnings"] for t in th) == 0
    assert np.abs(et - et[0]).max() < 2e-3 * abs(et[0]), et
    assert th[-1]["eangle"] > 0.1
    e.close(); e2.close()
