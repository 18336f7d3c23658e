"""CPU tests: the oracle restatement (oracle/restate.py) against the golden vectors.

Pins the oracle on (1) the reference's own unit-test vectors (tests/golden/ref_yaml_*.npz, extracted from
unittest/force-styles/tests/*.yaml by oracle/extract_ref_yaml.py) and (2) outputs of the compiled reference
(tests/golden/forces_chain.npz, le_trace_small.npz; oracle/make_golden.py).
"""
import os

import numpy as np
import pytest

from oracle import restate as R
from oracle.make_golden import unpack_trace
from oracle import refio
from tests import lehelpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHROMATIN_BONDS = {1: ("fene", (30.0, 1.5, 1.0, 1.0)), 2: ("harmonic", (20.0, 1.3))}


def test_ranmars_counter_and_range():
    r = R.RanMars(12345)
    vals = [r.uniform() for _ in range(1000)]
    assert all(0.0 <= v < 1.0 for v in vals)
    assert all(v * 16777216.0 == int(v * 16777216.0) for v in vals)      # 24-bit values: the GPU integer restatement is exact
    assert refio.draws_consumed(r.c24()) == 1000 == r.ncalls
    # known answers: first draws of the seeds the decks use, as the compiled reference produced them
    z = np.load(os.path.join(GOLD, "ranmars_ref.npz"))
    for seed, ref in zip(z["seeds"], z["draws"]):
        rr = R.RanMars(int(seed))
        got = np.array([rr.uniform() for _ in range(ref.shape[0])])
        assert (got == ref).all(), "RanMars(%d) differs from the reference" % seed


@pytest.fixture(scope="module")
def force_rec():
    z = np.load(os.path.join(GOLD, "forces_chain.npz"))
    rec = {k: z[k] for k in z.files}
    return rec


def test_special_build_matches_reference(force_rec):
    got = R.special_build(force_rec["num_bond"], force_rec["bond_atom"], special_lj=(0.0, 1.0, 1.0))   # special_bonds fene
    ref = R.special_tiers(force_rec["nspecial"], force_rec["special"])
    assert got == ref


def test_bond_list_matches_reference(force_rec):
    L = force_rec["boxhi"] - force_rec["boxlo"]
    bl = R.bond_list(force_rec["x"], L, force_rec["num_bond"], force_rec["bond_type"], force_rec["bond_atom"])
    assert bl.shape == force_rec["bondlist"].shape and (bl == force_rec["bondlist"]).all()


def test_half_neighbor_list_matches_reference(force_rec):
    rows = R.half_neighbor_list(force_rec["x"], force_rec["boxlo"], force_rec["boxhi"], 1.12246 + 0.4,
                                force_rec["nspecial"], force_rec["special"])
    ref = H.neigh_sets(force_rec["neigh_offsets"], force_rec["neigh_entries"])
    bad = [t + 1 for t in range(len(rows)) if frozenset(rows[t]) != ref[t]]
    assert not bad, "half-list sets differ for tags %s" % bad[:10]


def test_forces_energy_virial_match_reference(force_rec):
    x = force_rec["x"]
    n = len(x)
    L = force_rec["boxhi"] - force_rec["boxlo"]
    off, ent = force_rec["neigh_offsets"], force_rec["neigh_entries"]
    pi = np.repeat(np.arange(n), np.diff(off))
    pj = (ent & R.NEIGHMASK) - 1
    which = (ent >> R.SBBITS) & 3
    co = R.lj_coeffs(1.0, 1.0, 1.12246, True)
    fp, evdwl, vp = R.pair_lj_cut(x, L, pi, pj, which, co)
    b1, b2, bt = R.unique_bonds(force_rec["num_bond"], force_rec["bond_type"], force_rec["bond_atom"])
    fb, eb, vb, _ = R.bond_forces(x, L, b1, b2, bt, CHROMATIN_BONDS)
    f = fp + fb
    fr = force_rec["f"]
    mag = np.sqrt((fr ** 2).sum(1))
    err = np.sqrt(((f - fr) ** 2).sum(1)) / np.maximum(mag, np.sqrt((mag ** 2).mean()))
    assert err.max() < 1e-11
    assert abs(evdwl - force_rec["energy"][0]) <= 1e-10 * abs(force_rec["energy"][0])
    assert abs(eb - force_rec["energy"][1]) <= 1e-10 * abs(force_rec["energy"][1])
    assert np.abs(vp - force_rec["virial_pair"]).max() <= 1e-9 * np.abs(force_rec["virial_pair"]).max()
    assert np.abs(vb - force_rec["virial_bond"]).max() <= 1e-9 * np.abs(force_rec["virial_bond"]).max()


def _yaml_case(name):
    p = os.path.join(GOLD, "ref_yaml_%s.npz" % name)
    z = np.load(p, allow_pickle=False)
    return {k: z[k] for k in z.files}


def test_reference_unit_vectors_lj_cut():
    """unittest/force-styles/tests/mol-pair-lj_cut.yaml: init_forces / init_vdwl / init_stress of the 29-atom molecule"""
    c = _yaml_case("mol-pair-lj_cut")
    x, typ = c["x"], c["type"]
    n = len(x)
    L = c["boxhi"] - c["boxlo"]
    # all pairs within the pair cutoff; special weights from special_tiers (special_bonds lj 0 0 0 of the yaml's prerequisites)
    tiers = R.special_build(c["num_bond"], c["bond_atom"])
    # the 8.0 cutoff exceeds half the 15.0 box: a pair can see more than one periodic image, each is a ghost of its own.
    # Special weights apply to every image of a special partner (the list stores them per local/ghost index by tag).
    import itertools
    pi, pj, which, sh = [], [], [], []
    for i in range(n - 1):
        for j in range(i + 1, n):
            w = 0
            for k in range(3):
                if (j + 1) in tiers[i][k]:
                    w = k + 1
            for s in itertools.product((-1, 0, 1), repeat=3):
                pi.append(i); pj.append(j); which.append(w); sh.append(s)
    pi, pj, which, sh = np.array(pi), np.array(pj), np.array(which), np.array(sh, float)
    eps, sig, cut = c["epsilon"], c["sigma"], c["cut"]     # [ntypes, ntypes] after mixing, as the yaml's pair_coeff lines give them
    ti, tj = typ[pi] - 1, typ[pj] - 1
    co = {k: np.empty(len(pi)) for k in ("lj1", "lj2", "lj3", "lj4", "offset", "cutsq")}
    for a in range(eps.shape[0]):
        for b in range(eps.shape[0]):
            m = (ti == a) & (tj == b)
            cc = R.lj_coeffs(eps[a, b], sig[a, b], cut[a, b], False)
            for k in co:
                co[k][m] = cc[k]
    f, evdwl, vir = R.pair_lj_cut(x, L, pi, pj, which, co, special_lj=tuple(c["special_lj"]), shift=sh)
    assert np.abs(f - c["init_forces"]).max() <= 5e-13 * max(1.0, np.abs(c["init_forces"]).max())
    assert abs(evdwl - c["init_vdwl"]) <= 5e-13 * abs(c["init_vdwl"])
    assert np.abs(vir - c["init_stress"]).max() <= 5e-13 * np.abs(c["init_stress"]).max()


def test_reference_unit_vectors_angle_cosine():
    """unittest/force-styles/tests/angle-cosine.yaml: init_forces / init_energy / init_stress of the 29-atom molecule"""
    c = _yaml_case("angle-cosine")
    L = c["boxhi"] - c["boxlo"]
    ang = c["angles"]
    f, e, vir = R.angle_cosine(c["x"], L, ang[:, 1] - 1, ang[:, 2] - 1, ang[:, 3] - 1, ang[:, 0], {t + 1: k[0] for t, k in enumerate(c["angle_coeff"])})
    assert np.abs(f - c["init_forces"]).max() <= 5e-13 * max(1.0, np.abs(c["init_forces"]).max())
    assert abs(e - c["init_energy"]) <= 5e-13 * abs(c["init_energy"])
    assert np.abs(vir - c["init_stress"]).max() <= 5e-13 * np.abs(c["init_stress"]).max()


@pytest.mark.parametrize("name", ["bond-fene", "bond-harmonic"])
def test_reference_unit_vectors_bonds(name):
    """unittest/force-styles/tests/bond-fene.yaml / bond-harmonic.yaml: init_forces / init_energy / init_stress"""
    c = _yaml_case(name)
    x = c["x"]
    L = c["boxhi"] - c["boxlo"]
    style = "fene" if name == "bond-fene" else "harmonic"
    coeffs = {k + 1: (style, tuple(c["bond_coeff"][k])) for k in range(len(c["bond_coeff"]))}
    b1, b2, bt = R.unique_bonds(c["num_bond"], c["bond_type"], c["bond_atom"])
    f, e, vir, _ = R.bond_forces(x, L, b1, b2, bt, coeffs)
    assert np.abs(f - c["init_forces"]).max() <= 5e-13 * np.abs(c["init_forces"]).max()
    assert abs(e - c["init_energy"]) <= 5e-13 * abs(c["init_energy"])
    assert np.abs(vir - c["init_stress"]).max() <= 5e-13 * np.abs(c["init_stress"]).max()


def replay_with_oracle(pre, post, cfg):
    """every recorded USER-LE event: oracle post-state == reference post-state, bit-exact bond rows, special lists in
    exact order, types, counters and number of Marsaglia draws; returns the number of events per fix"""
    seen = {1: 0, 2: 0, 3: 0}
    changed = {1: 0, 2: 0, 3: 0}
    for a, b in zip(pre, post):
        w = a["which"]
        S = R.copy_state(a)
        slot = H.RNG_SLOT[w]
        rng = R.RanMars(cfg[H.SEED_KEY[w]]["seed"]).skip(refio.draws_consumed(a["rngc"][slot]))
        L = S["L"]
        # the lists of the last rebuild are part of the recorded pre-state; the restated bond list must equal it
        bl = R.bond_list(a["xhold"], L, a["num_bond"], a["bond_type"], a["bond_atom"])
        assert bl.shape == a["bondlist"].shape and (bl == a["bondlist"]).all()
        if w == 1:
            c = cfg["extrusion"]
            cnt, _ = R.fix_extrusion(S, rng, c["neutral"], c["left"], c["right"], c["p_through"], c["btype"], c["roadblock"])
        elif w == 2:
            c = cfg["ex_unload"]
            cnt = R.fix_ex_unload(S, rng, c["btype"], c["rc"], c["prob"])
        else:
            c = cfg["ex_load"]
            cnt = R.fix_ex_load(S, rng, c["itype"], c["jtype"], c["rc"], c["btype"], c["prob"], c["iparam"][0], c["iparam"][1],
                                c["jparam"][0], c["jparam"][1], a["neigh_offsets"], a["neigh_entries"])
        res = H.compare_topology(S, b)
        assert not any(res.values()), "event step %d fix %d: %s" % (a["step"], w, res)
        assert (S["type"] == b["type"]).all()
        assert cnt == b["counters"][w - 1]
        assert rng.c24() == b["rngc"][slot], "draw count differs at step %d fix %d" % (a["step"], w)
        seen[w] += 1
        changed[w] += int((a["bond_atom"] != b["bond_atom"]).any() or (a["num_bond"] != b["num_bond"]).any())
    return seen, changed


def test_le_events_replay_golden_trace():
    pre, post = unpack_trace(np.load(os.path.join(GOLD, "le_trace_small.npz")))
    seen, _ = replay_with_oracle(pre, post, H.LE_DECK)
    assert seen[1] >= 3 and seen[2] >= 1 and seen[3] >= 1


LE_VARIANTS = {
    # transparent barriers, dense cadence: many slides, loads on a chain that already carries extruders
    "open": dict(barriers="random", nbeads=900, next_=30, steps=620, seed=3, deck={
        "extrusion": dict(nevery=100, neutral=1, left=2, right=3, p_through=1.0, btype=2, roadblock=4, seed=12345),
        "ex_load": dict(nevery=50, itype=1, jtype=1, rc=1.12, btype=2, prob=0.2, seed=99, iparam=(1, 1), jparam=(1, 1)),
        "ex_unload": dict(nevery=50, btype=2, rc=0.5, prob=0.2, seed=7)}),
    # closed barriers every 100 beads (configs[1] layout): extruders stall at CTCF sites and behind each other
    "closed": dict(barriers="periodic", nbeads=1000, next_=60, steps=620, seed=4, deck={
        "extrusion": dict(nevery=100, neutral=1, left=2, right=3, p_through=0.0, btype=2, roadblock=4, seed=777),
        "ex_load": dict(nevery=100, itype=1, jtype=1, rc=1.12, btype=2, prob=0.05, seed=684474, iparam=(1, 1), jparam=(1, 1)),
        "ex_unload": dict(nevery=100, btype=2, rc=0.5, prob=0.05, seed=456456)}),
}


@pytest.mark.parametrize("name", sorted(LE_VARIANTS))
def test_le_events_replay_live_reference_variants(name):
    """the restatement against the compiled reference (oracle/_ref) on USER-LE settings the golden trace does not hold:
    fully transparent and fully closed barriers, dense event cadence, more extruders per bead"""
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    from lammps_le_b200 import systems
    from oracle.make_golden import le_trace
    v = LE_VARIANTS[name]
    s = systems.chromatin_chain(v["nbeads"], v["next_"], rho=0.2, seed=v["seed"], barriers=v["barriers"],
                                p_left=0.03, p_right=0.03, p_block=0.01)
    pre, post, _ = le_trace(s, v["steps"], H.le_deck_lines(v["deck"]))
    seen, changed = replay_with_oracle(pre, post, v["deck"])
    assert seen[1] >= 5 and seen[2] >= 5 and seen[3] >= 5
    assert changed[1] >= 3, "the extrusion fix must have moved bonds in this trace"


def test_fp32_pair_math_error_budget(force_rec):
    """Error budget of the optional fp32 pair path (LE_PAIR_FP32=1, pair_term32 in csrc/le_md.cuh), emulated operation for
    operation in numpy float32 on the golden chain: WCA pair terms in fp32 on the fixed-point differences, bonds in
    fp64.  The per-atom force error must stay below the 1e-5 bar (it sits near 6e-6: in a minimised state pair and
    bond terms of order 10 cancel to a net force of order 1); the default fp64 pair path is at 1e-12."""
    c = force_rec
    x, lo, hi = c["x"], c["boxlo"], c["boxhi"]
    L = hi - lo
    n = len(x)
    rows = R.half_neighbor_list(x, lo, hi, 1.12246 + 0.4, c["nspecial"], c["special"])
    pi = np.array([t for t in range(n) for _ in rows[t]], dtype=int)
    pj = np.array([(v & R.NEIGHMASK) - 1 for t in range(n) for v in rows[t]], dtype=int)
    b1, b2, bt = R.unique_bonds(c["num_bond"], c["bond_type"], c["bond_atom"])
    fb, _, _, _ = R.bond_forces(x, L, b1, b2, bt, {1: ("fene", (30.0, 1.5, 1.0, 1.0)), 2: ("harmonic", (20.0, 1.3))})
    u = np.rint((x - lo) / L * 4294967296.0).astype(np.int64)
    d = ((u[pi] - u[pj]) + 2 ** 31) % 2 ** 32 - 2 ** 31                  # signed 32-bit difference = minimum image
    df = d.astype(np.float32) * (L / 4294967296.0).astype(np.float32)[None, :]

    def fma(a, b, cc):
        return (a.astype(np.float64) * b.astype(np.float64) + cc.astype(np.float64)).astype(np.float32)

    rsq = fma(df[:, 2], df[:, 2], fma(df[:, 0], df[:, 0], df[:, 1] * df[:, 1]))
    inn = rsq < np.float32(1.12246 ** 2)
    r2 = np.float32(1.0) / np.where(inn, rsq, np.float32(1.0))
    r6 = (r2 * r2) * r2
    fp = np.where(inn, (r6 * fma(np.full_like(r6, 48.0), r6, np.full_like(r6, -24.0))) * r2, np.float32(0.0))
    contrib = (df * fp[:, None]).astype(np.float64)
    f = fb.copy()
    np.add.at(f, pi, contrib)
    np.add.at(f, pj, -contrib)
    mag = np.sqrt((c["f"] ** 2).sum(1))
    err = np.sqrt(((f - c["f"]) ** 2).sum(1)) / np.maximum(mag, np.sqrt((mag ** 2).mean()))
    assert inn.sum() > 100
    assert 1e-7 < err.max() <= 1e-5, "fp32 pair path: max per-atom relative force error %.3g" % err.max()


def test_lists_and_forces_match_live_reference_on_a_dense_melt():
    """the restatement against the compiled reference on a dense FENE melt (rho* = 0.8442, ten half neighbors per atom,
    many pairs stored through periodic ghosts): special lists, bond list, half list sets, forces, energies, virials"""
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    from lammps_le_b200 import systems
    from oracle.make_golden import force_case
    m = systems.fene_melt(8, 60, rho=0.8442)
    rec = force_case(m, velocities=True)
    n = len(rec["x"])
    L = rec["boxhi"] - rec["boxlo"]
    assert R.special_build(rec["num_bond"], rec["bond_atom"], special_lj=(0.0, 1.0, 1.0)) == R.special_tiers(rec["nspecial"], rec["special"])
    bl = R.bond_list(rec["x"], L, rec["num_bond"], rec["bond_type"], rec["bond_atom"])
    assert bl.shape == rec["bondlist"].shape and (bl == rec["bondlist"]).all()
    rows = R.half_neighbor_list(rec["x"], rec["boxlo"], rec["boxhi"], 1.12246 + 0.4, rec["nspecial"], rec["special"])
    ref = H.neigh_sets(rec["neigh_offsets"], rec["neigh_entries"])
    assert sum(len(r) for r in rows) > 4 * n
    bad = [t + 1 for t in range(n) if frozenset(rows[t]) != ref[t]]
    assert not bad, "half-list sets differ for tags %s" % bad[:10]
    pi = np.array([t for t in range(n) for _ in rows[t]], dtype=int)
    pj = np.array([(v & R.NEIGHMASK) - 1 for t in range(n) for v in rows[t]], dtype=int)
    fp, evdwl, vp = R.pair_lj_cut(rec["x"], L, pi, pj, np.zeros(len(pi), int), R.lj_coeffs(1.0, 1.0, 1.12246, True))
    b1, b2, bt = R.unique_bonds(rec["num_bond"], rec["bond_type"], rec["bond_atom"])
    fb, eb, vb, _ = R.bond_forces(rec["x"], L, b1, b2, bt, {1: ("fene", (30.0, 1.5, 1.0, 1.0))})
    fr = rec["f"]
    mag = np.sqrt((fr ** 2).sum(1))
    err = np.sqrt(((fp + fb - fr) ** 2).sum(1)) / np.maximum(mag, np.sqrt((mag ** 2).mean()))
    assert err.max() < 1e-11
    assert abs(evdwl - rec["energy"][0]) <= 1e-10 * abs(rec["energy"][0]) and abs(eb - rec["energy"][1]) <= 1e-10 * abs(rec["energy"][1])
    assert np.abs(vp - rec["virial_pair"]).max() <= 1e-9 * np.abs(rec["virial_pair"]).max()
    assert np.abs(vb - rec["virial_bond"]).max() <= 1e-9 * np.abs(rec["virial_bond"]).max()


from oracle.make_golden import BOND_CREATE_CFG, bond_create_trace, unpack_trace as _unpack_trace  # noqa: E402


def _bond_create_events(source):
    if source == "golden":
        return _unpack_trace(np.load(os.path.join(GOLD, "bond_create_trace_small.npz")))
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    return bond_create_trace()


@pytest.mark.parametrize("source", ["golden", "live"])
def test_bond_create_replay(source):
    """fix bond/create (src/MC/fix_bond_create.cpp, the ancestor of fix ex_load): the restatement -- fix_ex_load's
    loops without the loop-extrusion rules, bond counts kept from the first run's setup -- against recorded events of the compiled
    reference (golden: tests/golden/bond_create_trace_small.npz, made by oracle/make_golden.py; live: every event of a fresh run):
    bond rows, special lists in exact order, types, created-bond counter and Marsaglia draws"""
    cfg = BOND_CREATE_CFG
    pre, post = _bond_create_events(source)
    # FixBondCreate::setup counts once and the fix then adds only its own creations; no other fix changes bonds in these traces, so the
    # counts at any event equal a recount of its pre-state (golden: the kept events are not consecutive) -- carried along in the live run
    bc = R._bondcount(R.copy_state(pre[0]), cfg["btype"], False) if source == "live" else None
    created = 0
    for a, b in zip(pre, post):
        assert a["which"] == 3
        S = R.copy_state(a)
        rng = R.RanMars(cfg["seed"]).skip(refio.draws_consumed(a["rngc"][2]))
        cnt = R.fix_ex_load(S, rng, cfg["itype"], cfg["jtype"], cfg["rc"], cfg["btype"], cfg["prob"], cfg["iparam"][0], cfg["iparam"][1],
                            cfg["jparam"][0], cfg["jparam"][1], a["neigh_offsets"], a["neigh_entries"], ancestor=True, bc=bc)
        res = H.compare_topology(S, b)
        assert not any(res.values()), "event step %d: %s" % (a["step"], res)
        assert (S["type"] == b["type"]).all()
        assert cnt == b["counters"][2]
        assert rng.c24() == b["rngc"][2], "draw count differs at step %d" % a["step"]
        created += cnt
    assert created > (40 if source == "live" else 20), "the events must create bonds"
    assert (post[-1]["type"] == 4).sum() > (pre[0]["type"] == 4).sum(), "beads with two created bonds change type"


def test_bond_break_replay_live_reference():
    """fix bond/break (src/MC/fix_bond_break.cpp, the ancestor of fix ex_unload: the same post_integrate body on multiples of N):
    the restatement of ex_unload against every event of a run of the compiled reference that breaks stretched extruder bonds"""
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    import tempfile
    from lammps_le_b200 import systems
    cfg = dict(nevery=10, btype=2, rc=1.2, prob=0.5, seed=456456)
    s = systems.chromatin_chain(1500, 60, rho=0.2, seed=21, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    wd = tempfile.mkdtemp(prefix="le_bb_trace_")
    refio.write_data_file(os.path.join(wd, "data.le"), s)
    deck = refio.deck_header(s, "data.le") + [
        "fix 1 all nve/limit 0.05", "fix 2 all langevin 1.0 1.0 1.0 904297", "fix s0 all le/snap pre.bin pre grid",
        "fix br all bond/break %d %d %g prob %g %d" % (cfg["nevery"], cfg["btype"], cfg["rc"], cfg["prob"], cfg["seed"]),
        "fix s1 all le/snap post.bin post", "thermo_style custom step temp bonds f_br[1] f_br[2]", "thermo 100", "timestep 0.005", "run 60"]
    refio.run_reference(deck, workdir=wd)
    pre = refio.read_records(os.path.join(wd, "pre.bin"))
    post = refio.read_records(os.path.join(wd, "post.bin"))
    assert len(pre) == len(post) == 6
    broken = 0
    for a, b in zip(pre, post):
        assert a["which"] == 2
        S = R.copy_state(a)
        rng = R.RanMars(cfg["seed"]).skip(refio.draws_consumed(a["rngc"][1]))
        cnt = R.fix_ex_unload(S, rng, cfg["btype"], cfg["rc"], cfg["prob"])
        res = H.compare_topology(S, b)
        assert not any(res.values()), "event step %d: %s" % (a["step"], res)
        assert cnt == b["counters"][1]
        assert rng.c24() == b["rngc"][1], "draw count differs at step %d" % a["step"]
        broken += cnt
    assert broken > 5, "the run must break bonds"
