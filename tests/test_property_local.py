"""`compute property/local batom1 batom2 btype` + `dump local` (the way loop sizes are taken out of a reference run):
the host helper lists the bonds in the reference's order -- checked against a `dump local` file of the compiled reference."""
import os
import tempfile

import numpy as np
import pytest

from lammps_le_b200 import systems
from lammps_le_b200.engine import property_local_bonds
from oracle import refio


def test_property_local_bonds_match_the_reference_dump_local():
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    s = systems.chromatin_chain(600, 25, rho=0.2, seed=13)
    wd = tempfile.mkdtemp(prefix="le_plocal_")
    refio.write_data_file(os.path.join(wd, "data.le"), s)
    deck = refio.deck_header(s, "data.le") + [
        "compute b all property/local batom1 batom2 btype",
        "dump d all local 1 bonds.dump index c_b[1] c_b[2] c_b[3]",
        "fix 1 all nve", "run 0"]
    final = os.path.join(wd, "final.bin")
    refio.run_reference(deck, workdir=wd, final=final)
    lines = open(os.path.join(wd, "bonds.dump")).read().splitlines()
    assert lines[0] == "ITEM: TIMESTEP" and lines[2] == "ITEM: NUMBER OF ENTRIES"
    n = int(lines[3])
    assert lines[8] == "ITEM: ENTRIES index c_b[1] c_b[2] c_b[3] "      # (the reference leaves a trailing blank)
    ref = np.array([[int(float(v)) for v in l.split()] for l in lines[9:9 + n]])
    rec = refio.read_records(final)[0]
    rows = property_local_bonds(rec["num_bond"], rec["bond_type"], rec["bond_atom"], newton_bond=0)
    assert len(rows) == n == 599 + 25
    assert (ref[:, 0] == np.arange(1, n + 1)).all()
    assert (rows == ref[:, 1:]).all()
    # loop sizes of the extruder bonds, the observable this output exists for
    loops = rows[rows[:, 2] == 2]
    assert (loops[:, 1] - loops[:, 0] == 2).all()
