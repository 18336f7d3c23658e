"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/le_b200.h declares,
fails loudly without a GPU, and the host-side helpers behave."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from lammps_le_b200 import engine
    lib = engine.load_library()
    hdr = open(os.path.join(ROOT, "include", "le_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(le_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_gpu_means_error_not_fallback(has_gpu):
    from lammps_le_b200 import engine
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.LeError) as ei:
        engine.Engine((0, 0, 0), (10, 10, 10))
    assert ei.value.code == -2


def test_product_does_not_import_oracle():
    pk = os.path.join(ROOT, "lammps_le_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".inl", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace("oracle/_ref", "").replace("oracle/", "oracle/") or "oracle" not in text, f


def test_saw_generator_is_self_avoiding_and_bonded():
    from lammps_le_b200 import systems
    s = systems.chromatin_chain(3000, 30, rho=0.2, seed=4)
    x, L = s["x"], s["box"][1][0]
    assert (x >= 0).all() and (x < L).all()
    d = x[1:] - x[:-1]
    d -= L * np.rint(d / L)
    assert np.allclose(np.sqrt((d ** 2).sum(1)), 0.97, atol=1e-9)
    # all pairs beyond the bonded neighbour keep the rejection radius (brute force on a subsample)
    idx = np.arange(0, 3000, 7)
    dd = x[idx][:, None, :] - x[None, :, :]
    dd -= L * np.rint(dd / L)
    r = np.sqrt((dd ** 2).sum(-1))
    r[np.abs(idx[:, None] - np.arange(3000)[None, :]) <= 1] = 9.0
    assert r.min() > 0.85
    bt, a1, a2 = s["bonds"]
    assert (bt == 2).sum() == 30 and ((a2 - a1)[bt == 2] == 2).all()
    assert unpacked_unwrapped_is_continuous(s)


def unpacked_unwrapped_is_continuous(s):
    from lammps_le_b200.engine import unpack_image
    L = s["box"][1][0]
    u = s["x"] + unpack_image(s["image"]) * L
    step = np.sqrt(((u[1:] - u[:-1]) ** 2).sum(1))
    return np.allclose(step, 0.97, atol=1e-9)


def test_lattice_melt_bonds_have_lattice_spacing():
    from lammps_le_b200 import systems
    m = systems.fene_melt(40, 100, 0.8442)
    x, L = m["x"], m["box"][1][0]
    bt, a1, a2 = m["bonds"]
    d = x[a2 - 1] - x[a1 - 1]
    d -= L * np.rint(d / L)
    r = np.sqrt((d ** 2).sum(1))
    assert len(bt) == 40 * 99 and r.max() < 1.2 and r.min() > 0.9


def test_le_deck_front_end_rejects_unknown_commands(tmp_path):
    """the C++ host front end stops on a command outside the path with the reference's error text (no GPU needed:
    nothing has touched the device yet)"""
    import subprocess
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    assert os.path.exists(exe), "le_deck is not built (make -C lammps_le_b200/csrc)"
    (tmp_path / "in.bad").write_text("units lj\natom_style bond\nkspace_style pppm 1e-4\n")
    r = subprocess.run([exe, "-in", "in.bad"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Unknown command: kspace_style" in r.stderr
    (tmp_path / "in.units").write_text("units real\n")
    r = subprocess.run([exe, "-in", "in.units"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "only units lj" in r.stderr


def test_le_deck_front_end_checks_thermo_style_and_velocity_before_touching_the_device(tmp_path):
    """thermo_style custom keywords and the velocity command are validated with the reference's error texts"""
    import subprocess
    exe = os.path.join(ROOT, "lammps_le_b200", "le_deck")
    cases = {
        "units lj\nthermo_style custom step temp nonsense\n": "Unknown keyword in thermo_style custom command: nonsense",
        "units lj\nthermo_style multi\n": "Illegal thermo_style command",
        "units lj\nthermo_style custom step temp pe ke etotal bonds vol\nvelocity all create 1.0 12345\n": "Velocity command before simulation box is defined",
        "units lj\nvelocity all create 1.0\n": "Illegal velocity command",
        "units lj\ncompute b all property/local batom1 foo\n": "Invalid keyword in compute property/local command",
        "units lj\ndump d all local 10 f.dump c_b[1]\n": "Could not find dump local compute ID",
    }
    for k, (deck, msg) in enumerate(cases.items()):
        f = tmp_path / ("in.%d" % k)
        f.write_text(deck)
        r = subprocess.run([exe, "-in", f.name], cwd=tmp_path, capture_output=True, text=True, timeout=60)
        assert r.returncode == 1 and msg in r.stderr, (deck, r.stderr)
