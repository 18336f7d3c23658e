"""Long-run statistical parity with the reference (north_star: thermodynamic and polymer observables -- Rg, contact
probability P(s), loop-size distribution -- must agree within stated statistical error).

tests/golden/stats_chain.npz (oracle/make_stats_golden.py) holds the relaxed start state written by the compiled
reference and the per-frame observables of TWO reference runs that differ only in the Langevin seed.  The CUDA
engine starts from the same state, runs the same 60,000 steps with the same three USER-LE fixes and is sampled the
same way.  Stated error: every time-averaged observable must lie within max(3 x the seed-to-seed difference of the
two reference runs, the relative tolerance listed below) of the reference mean."""
import os

import numpy as np
import pytest

from lammps_le_b200 import systems

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = {"temp": 0.01, "epair": 0.04, "emol": 0.01, "rg": 0.04, "nloops": 0.04, "mean_loop": 0.10}
CONTACT = 1.5


def ks_distance(a, b):
    grid = np.unique(np.concatenate([a, b]))
    ca = np.searchsorted(np.sort(a), grid, side="right") / len(a)
    cb = np.searchsorted(np.sort(b), grid, side="right") / len(b)
    return np.abs(ca - cb).max()


@pytest.mark.gpu
def test_long_run_observables_match_the_reference():
    z = np.load(os.path.join(GOLD, "stats_chain.npz"))
    z = {k: z[k] for k in z.files}
    n, next_, steps, every = (int(v) for v in z["params"])
    s_list = z["s_list"]
    s = systems.chromatin_chain(n, next_, rho=0.2, seed=21, barriers="periodic", extruder_bond=systems.EXTRUDER_FENE)
    s["x"], s["image"] = z["x"], z["image"]
    e = systems.make_engine(s, velocities=z["v"], dt=0.005)
    e.fix_nve(True)
    e.fix_langevin(1.0, 1.0, 1.0, 777)
    e.fix_extrusion(200, 1, 2, 3, 0.5, 2, 4, 12345)
    e.fix_ex_load(100, 1, 1, 1.12, 2, 0.05, 684474, (1, 1), (1, 1))
    e.fix_ex_unload(100, 2, 0.5, 0.02, 456456)
    e.thermo_every(every)
    rows, sizes_late = [], []
    nfr = steps // every
    for f in range(nfr + 1):
        if f:
            e.run(every)
            th = e.thermo(-1)
        else:
            e.run(0)
            th = e.thermo(-1)
        xu, _ = e.positions(unwrap=True)
        topo = e.topology()
        nb, bt, ba = topo["num_bond"], topo["bond_type"], topo["bond_atom"]
        mask = (np.arange(bt.shape[1])[None, :] < nb[:, None]) & (bt == 2) & (ba > (np.arange(n) + 1)[:, None])
        ii, mm = np.nonzero(mask)
        sizes = (ba[ii, mm] - (ii + 1)).astype(np.float64)
        rg = np.sqrt(((xu - xu.mean(0)) ** 2).sum(1).mean())
        ps = [(np.sqrt(((xu[sp:] - xu[:-sp]) ** 2).sum(1)) < CONTACT).mean() for sp in s_list]
        rows.append([th["step"], th["temp"], th["epair"], th["emol"], rg, len(sizes), sizes.mean() if len(sizes) else 0.0] + ps)
        if f >= nfr // 2:
            sizes_late.append(sizes)
    e.close()
    ours = np.array(rows)
    a, b = z["run_a"], z["run_b"]
    assert ours.shape == a.shape and (ours[:, 0] == a[:, 0]).all()
    cols = [str(c) for c in z["columns"]]
    skip = 10                                    # frames of the initial transient
    report = []
    for k, name in enumerate(cols):
        if name == "step":
            continue
        ma, mb, mo = a[skip:, k].mean(), b[skip:, k].mean(), ours[skip:, k].mean()
        ref = 0.5 * (ma + mb)
        rel = REL_TOL.get(name, 0.05)
        tol = max(3.0 * abs(ma - mb), rel * abs(ref), 0.003 if name.startswith("P(") else 0.0)
        report.append((name, mo, ref, tol, abs(mo - ref) <= tol))
    bad = [r for r in report if not r[4]]
    assert not bad, "observables outside the stated error: %s" % [(r[0], "ours %.5g ref %.5g tol %.3g" % r[1:4]) for r in bad]
    # loop-size distribution over the second half of the run: two-sample KS distance against reference run a
    mine = np.concatenate(sizes_late)
    d_ref = ks_distance(z["sizes_a"], z["sizes_b"])
    d = ks_distance(mine, z["sizes_a"])
    assert d <= 2.0 * d_ref + 0.05, "loop-size distribution: KS %.3f vs seed-to-seed %.3f" % (d, d_ref)
