"""`velocity all create` on the host (le_host_velocity_create) against the compiled reference (oracle/_ref): the same
Park-Miller draws in the same order, momentum zeroing and rescaling -- velocities must agree to rounding."""
import os
import tempfile

import numpy as np
import pytest

from lammps_le_b200 import systems
from lammps_le_b200.engine import LeError, velocity_create
from oracle import refio


def reference_velocities(s, line):
    wd = tempfile.mkdtemp(prefix="le_vel_")
    refio.write_data_file(os.path.join(wd, "data.le"), s)
    deck = refio.deck_header(s, "data.le") + [line, "fix 1 all nve", "run 0"]
    final = os.path.join(wd, "final.bin")
    refio.run_reference(deck, workdir=wd, final=final)
    rec = refio.read_records(final)[0]
    return rec["x"], rec["v"]


@pytest.mark.parametrize("opts", ["", "dist gaussian", "dist uniform mom no", "loop local", "loop geom dist gaussian", "mom yes rot no loop all"])
def test_velocity_create_matches_reference(opts):
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    s = systems.chromatin_chain(700, 10, rho=0.2, seed=8)
    s["masses"] = np.array([1.0, 2.0, 0.5, 1.5])           # per-type masses matter for the 1/sqrt(m) factor
    x, vref = reference_velocities(s, "velocity all create 1.3 4928459 " + opts)
    kw = {"dist": "uniform", "mom": True, "loop": "all"}
    w = opts.split()
    for k in range(0, len(w), 2):
        if w[k] == "dist": kw["dist"] = w[k + 1]
        elif w[k] == "mom": kw["mom"] = w[k + 1] == "yes"
        elif w[k] == "loop": kw["loop"] = w[k + 1]
    v = velocity_create(s["types"], s["masses"], 1.3, 4928459, x=x, **kw)
    assert np.abs(v - vref).max() <= 1e-13 * np.abs(vref).max(), "velocity create %r differs from the reference" % opts
    m = s["masses"][s["types"] - 1]
    assert abs((m[:, None] * v ** 2).sum() / (3 * len(m) - 3) - 1.3) < 1e-12


def test_velocity_create_rejects_bad_arguments():
    with pytest.raises(LeError):
        velocity_create(np.ones(10, np.int32), np.ones(1), 1.0, 0)          # seed <= 0: "Illegal velocity create command"
    with pytest.raises(LeError):
        velocity_create(np.ones(10, np.int32), np.ones(1), 1.0, 5, loop="geom")   # geom needs coordinates
