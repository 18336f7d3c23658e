"""GPU parity tests against the reference: golden fixtures made by oracle/make_golden.py from the compiled
reference, and -- when oracle/_ref travelled to this box -- fresh reference runs on larger inputs.

Bars (BASELINE.json north_star): neighbor lists and USER-LE bond topology bit-exact; forces and energies
within 1e-5 relative per atom.
"""
import os

import numpy as np
import pytest

from tests import lehelpers as H

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHROMATIN_BONDS = {1: ("fene", (30.0, 1.5, 1.0, 1.0)), 2: ("harmonic", (20.0, 1.3))}
FORCE_RTOL = 1e-5   # per atom, relative to max(|f_i|, rms |f|)


def _force_record_from_npz(z):
    rec = {k: z[k] for k in z.files}
    rec["bpa"], rec["maxspecial"] = int(z["bpa"]), int(z["maxspecial"])
    return rec


def check_lists(e, rec):
    off, ent = e.neighlist(half=True)
    bad = H.compare_neighlists(off, ent, rec["neigh_offsets"], rec["neigh_entries"])
    assert len(ent) == len(rec["neigh_entries"]), "half list size %d vs reference %d" % (len(ent), len(rec["neigh_entries"]))
    assert not bad, "half neighbor sets differ for %d atoms, first tags %s" % (len(bad), bad[:10])
    bl = e.bondlist()
    assert bl.shape == rec["bondlist"].shape, "bondlist length %d vs %d" % (len(bl), len(rec["bondlist"]))
    assert (bl == rec["bondlist"]).all(), "bondlist rows differ (order matters: it is the USER-LE visit order)"


def check_forces(e, rec, thermo_ref=None):
    f, th = e.compute_forces()
    fr = rec["f"]
    n = len(fr)
    mag = np.sqrt((fr ** 2).sum(1))
    rms = np.sqrt((mag ** 2).mean())
    err = np.sqrt(((f - fr) ** 2).sum(1)) / np.maximum(mag, rms)
    assert err.max() <= FORCE_RTOL, "max per-atom relative force error %.3g (atom %d)" % (err.max(), err.argmax() + 1)
    epair_ref, emol_ref = rec["energy"][0] / n, rec["energy"][1] / n
    assert abs(th["epair"] - epair_ref) <= 1e-5 * max(abs(epair_ref), 1e-3)
    assert abs(th["emol"] - emol_ref) <= 1e-5 * abs(emol_ref)
    w_ref = rec["virial_pair"] + rec["virial_bond"]
    w = np.array(th["virial"])
    assert np.abs(w - w_ref).max() <= 1e-5 * np.abs(w_ref[:3]).max()
    if thermo_ref is not None:
        assert abs(th["press"] - thermo_ref["Press"]) <= 2e-5 * max(abs(thermo_ref["Press"]), 1e-2)
        assert abs(th["temp"] - thermo_ref["Temp"]) <= 1e-6 + 1e-6 * thermo_ref["Temp"]
    return err.max()


def test_lists_and_forces_golden():
    rec = _force_record_from_npz(np.load(os.path.join(GOLD, "forces_chain.npz")))
    e = H.engine_from_record(rec, CHROMATIN_BONDS, positions="x")
    e.force_rebuild()
    check_lists(e, rec)
    t = rec["thermo"]
    check_forces(e, rec, {"Temp": t[0], "Press": t[4]})
    e.close()


def replay_events(pre, post, cfg=H.LE_DECK):
    """Replay every recorded USER-LE event from the reference's pre-state; compare the post-state."""
    problems = []
    nchecked = {1: 0, 2: 0, 3: 0}
    from oracle import refio
    for a, b in zip(pre, post):
        e = H.engine_from_record(a, CHROMATIN_BONDS, positions="xhold")
        H.define_le_fixes(e, cfg)
        e.force_rebuild()
        # lists of the last rebuild: the stale pair list ex_load scans and the bondlist the other two visit
        off, ent = e.neighlist(half=True)
        bad = H.compare_neighlists(off, ent, a["neigh_offsets"], a["neigh_entries"])
        bl = e.bondlist()
        if bad or bl.shape != a["bondlist"].shape or (bl != a["bondlist"]).any():
            problems.append(("lists", a["step"], len(bad)))
        e.set_positions(a["x"], a["image"])
        w = a["which"]
        slot = H.RNG_SLOT[w]
        e.fix_rng_reset(H.WHICH[w], cfg[H.SEED_KEY[w]]["seed"], refio.draws_consumed(a["rngc"][slot]))
        e.run_le_event(H.WHICH[w])
        got = e.topology()
        res = H.compare_topology(got, b)
        res["type"] = int((e.types() != b["type"]).sum())
        res["draws"] = int(e.fix_rng_consumed(H.WHICH[w]) != refio.draws_consumed(b["rngc"][slot]))
        st = e.stats()
        cnt = {1: st["last_extrusion_shifts"], 2: st["last_unloads"], 3: st["last_loads"]}[w]
        res["counter"] = int(cnt != b["counters"][w - 1])
        # special lists compare as tier SETS (dedup's swap-with-last order is not physical); exact order is reported
        hard = {k: v for k, v in res.items() if k != "special_exact" and v}
        if hard:
            problems.append((a["step"], w, hard))
        nchecked[w] += 1
        e.close()
    return problems, nchecked


def test_le_replay_golden():
    from oracle.make_golden import unpack_trace
    pre, post = unpack_trace(np.load(os.path.join(GOLD, "le_trace_small.npz")))
    problems, n = replay_events(pre, post)
    assert n[1] >= 3 and n[2] >= 1 and n[3] >= 1
    assert not problems, "USER-LE replay mismatches: %s" % problems[:5]


def _need_ref():
    from oracle import refio
    if not refio.have_reference():
        pytest.skip("oracle/_ref not present on this box")


def test_le_replay_live_reference():
    """a longer fresh trace: 4000 beads, 3000 steps (6 slide events, 30 loads, 30 unloads)"""
    _need_ref()
    from lammps_le_b200 import systems
    from oracle.make_golden import le_trace
    s = systems.chromatin_chain(4000, 60, rho=0.2, seed=21)
    pre, post, _ = le_trace(s, 3010, H.le_deck_lines())
    problems, n = replay_events(pre, post)
    assert n[1] >= 6
    assert not problems, "USER-LE replay mismatches: %s" % problems[:5]


@pytest.mark.parametrize("name", ["closed", "open"])
def test_le_replay_live_reference_barrier_variants(name):
    """fully closed (p_through 0, CTCF every 100 beads) and fully transparent (p_through 1) barriers, dense event cadence"""
    _need_ref()
    from lammps_le_b200 import systems
    from oracle.make_golden import le_trace
    from tests.test_oracle import LE_VARIANTS
    v = LE_VARIANTS[name]
    s = systems.chromatin_chain(v["nbeads"], v["next_"], rho=0.2, seed=v["seed"], barriers=v["barriers"],
                                p_left=0.03, p_right=0.03, p_block=0.01)
    pre, post, _ = le_trace(s, v["steps"], H.le_deck_lines(v["deck"]))
    problems, n = replay_events(pre, post, v["deck"])
    assert n[1] >= 5 and n[2] >= 5 and n[3] >= 5
    assert not problems, "USER-LE replay mismatches: %s" % problems[:5]


@pytest.mark.parametrize("source", ["golden", "live"])
def test_bond_create_replay(source):
    """fix bond/create event by event: recorded events of the compiled reference (golden fixture / every event of a fresh run)
    replayed from their pre-states on the GPU -- bond rows, special lists, types, created-bond counter and Marsaglia draws equal the
    reference's post-state (the same traces pin the restatement on the CPU, tests/test_oracle.py)"""
    from oracle import refio
    from lammps_le_b200.engine import LE_FIX_EX_LOAD
    from oracle.make_golden import BOND_CREATE_CFG as cfg, bond_create_trace, unpack_trace
    if source == "golden":
        pre, post = unpack_trace(np.load(os.path.join(GOLD, "bond_create_trace_small.npz")))
    else:
        _need_ref()
        pre, post = bond_create_trace()
    problems, created = [], 0
    for a, b in zip(pre, post):
        e = H.engine_from_record(a, CHROMATIN_BONDS, positions="xhold")
        e.fix_bond_create(cfg["nevery"], cfg["itype"], cfg["jtype"], cfg["rc"], cfg["btype"], cfg["prob"], cfg["seed"], cfg["iparam"], cfg["jparam"])
        e.force_rebuild()
        e.set_positions(a["x"], a["image"])
        e.fix_rng_reset(LE_FIX_EX_LOAD, cfg["seed"], refio.draws_consumed(a["rngc"][2]))
        e.run_le_event(LE_FIX_EX_LOAD)        # (bond counts = those of the pre-state: the trace holds no other fix that changes bonds)
        res = H.compare_topology(e.topology(), b)
        res["type"] = int((e.types() != b["type"]).sum())
        res["draws"] = int(e.fix_rng_consumed(LE_FIX_EX_LOAD) != refio.draws_consumed(b["rngc"][2]))
        res["counter"] = int(e.stats()["last_loads"] != b["counters"][2])
        created += b["counters"][2]
        hard = {k: v for k, v in res.items() if k != "special_exact" and v}
        if hard:
            problems.append((a["step"], hard))
        e.close()
    assert created > (40 if source == "live" else 20)
    assert not problems, "fix bond/create replay mismatches: %s" % problems[:5]


def test_forces_live_reference_chain_and_melt():
    _need_ref()
    from lammps_le_b200 import systems
    from oracle.make_golden import force_case
    s = systems.chromatin_chain(20000, 200, rho=0.2, seed=3)      # BASELINE config C2 size
    rec = force_case(s)
    e = H.engine_from_record(rec, CHROMATIN_BONDS, positions="x")
    e.force_rebuild()
    check_lists(e, rec)
    check_forces(e, rec, rec["thermo"])
    e.close()
    m = systems.fene_melt(40, 100, rho=0.8442)                    # dense melt, lattice start + minimise
    rec = force_case(m, velocities=True)
    e = H.engine_from_record(rec, {1: ("fene", (30.0, 1.5, 1.0, 1.0))}, positions="x")
    e.force_rebuild()
    check_lists(e, rec)
    check_forces(e, rec, rec["thermo"])
    e.close()


BENCH_BONDS = {1: ("fene", (30.0, 1.5, 1.0, 1.0)), 2: ("fene", (10.0, 4.0, 1.0, 1.0))}


def test_bench_workload_at_one_million_beads_against_the_reference():
    """The system bench.py times (10^6-bead chain, 10^4 FENE(10,4) extruders, random barriers) against the compiled
    reference at full size: half neighbor list, bondlist, step-0 forces / energies / virial from BOTH instantiations of
    the step kernel, and one replayed event of each USER-LE fix."""
    _need_ref()
    import tempfile
    from lammps_le_b200 import systems
    from oracle.make_golden import force_case, le_trace
    n = 1000000
    s = systems.chromatin_chain(n, n // 100, rho=0.2, seed=12345, barriers="random", extruder_bond=systems.EXTRUDER_FENE)
    with tempfile.TemporaryDirectory(prefix="le_1m_") as wd:
        rec = force_case(s, workdir=wd, min_args="1e-4 1e-6 30 300")
    e = H.engine_from_record(rec, BENCH_BONDS, positions="x")
    e.force_rebuild()
    off, ent = e.neighlist(half=True)
    assert len(ent) == len(rec["neigh_entries"]), "half list size %d vs reference %d" % (len(ent), len(rec["neigh_entries"]))
    assert H.neighlists_equal_as_sets(off, ent, rec["neigh_offsets"], rec["neigh_entries"]) == 0
    bl = e.bondlist()
    assert bl.shape == rec["bondlist"].shape and (bl == rec["bondlist"]).all()
    err = check_forces(e, rec, rec["thermo"])
    fp = e.compute_forces_plain()
    fr = rec["f"]
    mag = np.sqrt((fr ** 2).sum(1))
    rel = (np.sqrt(((fp - fr) ** 2).sum(1)) / np.maximum(mag, np.sqrt((mag ** 2).mean()))).max()
    assert rel <= FORCE_RTOL, "plain step kernel: max per-atom relative force error %.3g" % rel
    e.close()
    # one event of each fix at this size (steps 1, 2, 3 of a run from the minimised state), replayed from the reference's pre-state
    with tempfile.TemporaryDirectory(prefix="le_1m_") as wd:
        lines = H.le_deck_lines()
        pre, post, _ = le_trace(s, 3, lines, workdir=wd, min_args="1e-4 1e-6 30 300")
    assert sorted(a["which"] for a in pre) == [1, 2, 3]
    problems = []
    from oracle import refio
    for a, b in zip(pre, post):
        e = H.engine_from_record(a, BENCH_BONDS, positions="xhold")
        H.define_le_fixes(e)
        e.force_rebuild()
        off, ent = e.neighlist(half=True)
        if H.neighlists_equal_as_sets(off, ent, a["neigh_offsets"], a["neigh_entries"]):
            problems.append(("lists", a["step"]))
        e.set_positions(a["x"], a["image"])
        w = a["which"]
        e.fix_rng_reset(H.WHICH[w], H.LE_DECK[H.SEED_KEY[w]]["seed"], refio.draws_consumed(a["rngc"][H.RNG_SLOT[w]]))
        e.run_le_event(H.WHICH[w])
        got = e.topology()
        for key in ("num_bond", "bond_type", "bond_atom", "nspecial"):
            mask = slice(None)
            if key in ("bond_type", "bond_atom"):
                m = np.arange(b[key].shape[1])[None, :] < b["num_bond"][:, None]
                if ((got[key] != b[key]) & m).any():
                    problems.append((key, w))
            elif (got[key] != b[key]).any():
                problems.append((key, w))
        if (e.types() != b["type"]).any():
            problems.append(("type", w))
        if e.fix_rng_consumed(H.WHICH[w]) != refio.draws_consumed(b["rngc"][H.RNG_SLOT[w]]):
            problems.append(("draws", w))
        e.close()
    assert not problems, problems


@pytest.mark.gpu
def test_ranmars_state_handover_matches_the_oracle():
    """le_fix_rng_set_state / le_fix_rng_get_state (RanMars::get_state layout, src/random_mars.cpp:297-319): the device
    generator after `consumed` draws holds exactly the oracle's u[], i97, j97, c; a state set from the oracle reads back
    unchanged and continues the same stream"""
    from oracle import restate as R
    from lammps_le_b200 import systems
    from lammps_le_b200.engine import LE_FIX_EX_LOAD, LE_FIX_EX_UNLOAD
    s = systems.chromatin_chain(2000, 20, rho=0.2, seed=11)
    e = systems.make_engine(s, velocities=np.zeros((2000, 3)))

    def state_of(g):
        return np.array(list(g.u) + [g.i97, g.j97, g.c, g.cd, g.cm], dtype=np.float64)

    for seed, consumed in ((684474, 0), (456456, 5), (12345, 1234)):
        g = R.RanMars(seed)
        for _ in range(consumed):
            g.uniform()
        e.fix_rng_reset(LE_FIX_EX_LOAD, seed, consumed)
        got = e.fix_rng_get_state(LE_FIX_EX_LOAD)
        assert np.array_equal(got, state_of(g)), (seed, consumed)
        e.fix_rng_set_state(LE_FIX_EX_UNLOAD, state_of(g))
        assert np.array_equal(e.fix_rng_get_state(LE_FIX_EX_UNLOAD), state_of(g))
    bad = state_of(R.RanMars(77)); bad[99] = 5
    with pytest.raises(Exception):
        e.fix_rng_set_state(LE_FIX_EX_LOAD, bad)
    e.close()
