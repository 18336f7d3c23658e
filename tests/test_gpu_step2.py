"""GPU test: the second-generation step kernel (k_step2, csrc/le_step2.cuh) against the first (k_step).

k_step2 issues the work differently (buffer parity as an argument, batched tail rows, one survivor queue, persistent
grid) but does the same arithmetic in the same order: trajectories must agree BIT FOR BIT, through the
captured-graph path and through direct launches, with and without the thermostat, for a dilute chain (short rows)
and a dense melt (rows beyond the first batch)."""
import numpy as np
import pytest

from lammps_le_b200 import systems

pytestmark = pytest.mark.gpu

VARIANTS = [1, 3, 33, 35]   # LE_STEP_VARIANT: bit 0 k_step2, bit 1 128-thread blocks, bit 5 persistent grid (33 = the default)


def trajectory(monkeypatch, variant, system, v0, langevin, steps, dt):
    monkeypatch.setenv("LE_STEP_VARIANT", str(variant))
    e = systems.make_engine(system, velocities=v0, dt=dt)
    e.fix_nve(True)
    if langevin:
        e.fix_langevin(1.0, 1.0, 1.0, 4242)
    e.run(steps)                  # graphs of 8 timesteps + single steps
    us = e.run_timed(24)          # direct launches
    out = (e.positions(), e.velocities(), e.stats()["neigh_builds"])
    e.close()
    assert us > 0.0
    return out


def relaxed(system, n, steps):
    e = systems.make_engine(system, velocities=systems.maxwell_velocities(n, 1.0, np.ones(n), 3))
    systems.relax(e, steps=steps)
    x, im = e.positions()
    v = e.velocities()
    e.close()
    s = dict(system)
    s["x"], s["image"] = x, im
    return s, v


@pytest.mark.parametrize("langevin", [True, False])
def test_step2_matches_step_on_a_chain(monkeypatch, langevin):
    n = 6000
    s, v = relaxed(systems.chromatin_chain(n, 60, rho=0.2, seed=5), n, 600)
    ref = trajectory(monkeypatch, 0, s, v, langevin, 150, 0.005)
    assert ref[2] > 3, "the run must cross several rebuilds"
    for var in VARIANTS:
        got = trajectory(monkeypatch, var, s, v, langevin, 150, 0.005)
        assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[0][1], ref[0][1]), "positions differ, variant %d" % var
        assert np.array_equal(got[1], ref[1]), "velocities differ, variant %d" % var
        assert got[2] == ref[2]


def test_step2_matches_step_on_a_melt(monkeypatch):
    # rho = 0.8442: about ten listed neighbors per atom, so the rows beyond the first four carry most of the pairs
    s = systems.fene_melt(nchains=40, length=100)
    n = len(s["types"])
    s, v = relaxed(s, n, 400)
    ref = trajectory(monkeypatch, 0, s, v, True, 100, 0.005)
    for var in VARIANTS:
        got = trajectory(monkeypatch, var, s, v, True, 100, 0.005)
        assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[1], ref[1]), "melt trajectory differs, variant %d" % var


@pytest.mark.xfail(reason="LE_STEP_VARIANT bit 9 (reneighbor decision fused into the persistent step kernel) was written "
                          "after the last GPU call of round 1: first run pending", strict=False)
def test_fused_decide_matches_step(monkeypatch):
    n = 6000
    s, v = relaxed(systems.chromatin_chain(n, 60, rho=0.2, seed=5), n, 600)
    ref = trajectory(monkeypatch, 0, s, v, True, 150, 0.005)
    got = trajectory(monkeypatch, 1 + 32 + 512, s, v, True, 150, 0.005)
    assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]


@pytest.mark.xfail(reason="LE_STEP_VARIANT bit 10 (dynamic tile fetch, k_step2d) was written after the last GPU call of round 1: first "
                          "run pending", strict=False)
def test_dynamic_tiles_match_step(monkeypatch):
    n = 6000
    s, v = relaxed(systems.chromatin_chain(n, 60, rho=0.2, seed=5), n, 600)
    ref = trajectory(monkeypatch, 0, s, v, True, 150, 0.005)
    got = trajectory(monkeypatch, 1 + 32 + 1024, s, v, True, 150, 0.005)
    assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]


PENDING = "fp32 pair terms (LE_PAIR_FP32=1) were written after the last GPU call of round 1: first run pending"


@pytest.mark.xfail(reason=PENDING, strict=False)
def test_pair_fp32_forces_meet_the_per_atom_bar(monkeypatch):
    """pair terms in fp32, bonds in fp64: step-0 forces of the golden chain within 1e-5 per atom of the reference"""
    import os
    from tests import lehelpers as H
    from tests.test_gpu_parity import CHROMATIN_BONDS, GOLD, _force_record_from_npz, check_forces
    monkeypatch.setenv("LE_PAIR_FP32", "1")
    rec = _force_record_from_npz(np.load(os.path.join(GOLD, "forces_chain.npz")))
    e = H.engine_from_record(rec, CHROMATIN_BONDS, positions="x")
    e.force_rebuild()
    err = check_forces(e, rec)          # asserts <= 1e-5 per atom, energies and virial to 1e-5
    e.close()
    assert err is None or err <= 1e-5


@pytest.mark.xfail(reason=PENDING, strict=False)
def test_pair_fp32_step2_matches_step(monkeypatch):
    """k_step2p and k_step sum an atom's fp32 pair terms in the same order: same bits"""
    monkeypatch.setenv("LE_PAIR_FP32", "1")
    n = 6000
    s, v = relaxed(systems.chromatin_chain(n, 60, rho=0.2, seed=5), n, 600)
    ref = trajectory(monkeypatch, 0, s, v, True, 150, 0.005)
    for var in (1, 33):
        got = trajectory(monkeypatch, var, s, v, True, 150, 0.005)
        assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[1], ref[1]), "variant %d" % var
    m = systems.fene_melt(nchains=40, length=100)
    m, vm = relaxed(m, len(m["types"]), 400)
    ref = trajectory(monkeypatch, 0, m, vm, True, 100, 0.005)
    got = trajectory(monkeypatch, 33, m, vm, True, 100, 0.005)
    assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[1], ref[1]), "melt"
