"""Drive the compiled reference (oracle/_ref) and read what oracle/ref_harness.cpp writes.

TEST INFRASTRUCTURE ONLY -- imported by tests/, oracle/make_golden.py, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by lammps_le_b200/.
"""
import os
import re
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
HARNESS = os.path.join(REF_DIR, "ref_harness")
LMP = os.path.join(REF_DIR, "lmp_ref")

_HDR = struct.Struct("<8i2q6d3i4i4x2d12dd")   # the C struct pads the 7 ints to an 8-byte boundary


def have_reference():
    return os.path.exists(HARNESS) and os.path.exists(os.path.join(REF_DIR, "liblammps_ref.so"))


def read_records(path):
    """Parse a le/snap or -final file into a list of dicts (arrays in tag order)."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos + _HDR.size <= len(data):
        h = _HDR.unpack_from(data, pos)
        pos += _HDR.size
        magic, kind, which, n, bpa, ms, nbl, has_force = h[:8]
        assert magic == 0x4C455331, "bad record magic"
        step, nneigh = h[8:10]
        rec = {"kind": kind, "which": which, "n": n, "bpa": bpa, "maxspecial": ms, "step": step,
               "boxlo": np.array(h[10:13]), "boxhi": np.array(h[13:16]), "rngc": list(h[16:19]),
               "counters": list(h[19:23]), "energy": list(h[23:25]), "virial_pair": np.array(h[25:31]),
               "virial_bond": np.array(h[31:37]), "ke2": h[37], "has_force": has_force}

        def take(dtype, count, shape=None):
            nonlocal pos
            a = np.frombuffer(data, dtype=dtype, count=count, offset=pos).copy()
            pos += a.nbytes
            return a.reshape(shape) if shape else a

        rec["x"] = take(np.float64, 3 * n, (n, 3))
        rec["xhold"] = take(np.float64, 3 * n, (n, 3))
        rec["v"] = take(np.float64, 3 * n, (n, 3))
        rec["f"] = take(np.float64, 3 * n, (n, 3))
        rec["image"] = take(np.int32, n)
        rec["type"] = take(np.int32, n)
        rec["num_bond"] = take(np.int32, n)
        rec["bond_type"] = take(np.int32, n * bpa, (n, bpa))
        rec["bond_atom"] = take(np.int32, n * bpa, (n, bpa))
        rec["nspecial"] = take(np.int32, 3 * n, (n, 3))
        rec["special"] = take(np.int32, n * ms, (n, ms))
        rec["bondlist"] = take(np.int32, 3 * nbl, (nbl, 3)) if nbl else np.zeros((0, 3), np.int32)
        if kind != 1:
            # lists are written for pre (0) and final (2) records
            rec["neigh_offsets"] = take(np.int64, n + 1)
            rec["neigh_entries"] = take(np.int32, nneigh)
        out.append(rec)
    return out


CM = 16777213
CD = 7654321
C0 = 362436


def draws_consumed(c24):
    """Number of uniform() calls a fix has made, from RanMars::c*2^24 (c_n = c_0 - n*cd mod cm), not counting
    the one in the constructor (src/random_mars.cpp:67)."""
    if c24 < 0:
        return 0
    n = ((C0 - c24) * pow(CD, -1, CM)) % CM
    return n - 1


def write_data_file(path, system, extra_bond=2, extra_special=40):
    """LAMMPS data file (atom_style bond) for a lammps_le_b200.systems dict, atoms in tag order."""
    lo, hi = system["box"]
    x, types, image = system["x"], system["types"], system["image"]
    n = len(types)
    bt, a1, a2 = system["bonds"]
    im = np.stack([(image & 1023) - 512, ((image >> 10) & 1023) - 512, ((image >> 20) & 1023) - 512], axis=1)
    with open(path, "w") as f:
        f.write("LAMMPS data file written by oracle/refio.py\n\n")
        f.write("%d atoms\n%d bonds\n\n%d atom types\n%d bond types\n\n" % (n, len(bt), system["ntypes"], system["nbondtypes"]))
        f.write("%d extra bond per atom\n%d extra special per atom\n\n" % (extra_bond, extra_special))
        for k, ax in enumerate("xyz"):
            f.write("%.17g %.17g %slo %shi\n" % (lo[k], hi[k], ax, ax))
        f.write("\nMasses\n\n")
        for k, m in enumerate(system["masses"]):
            f.write("%d %.17g\n" % (k + 1, m))
        f.write("\nAtoms # bond\n\n")
        mol = 1
        rows = ["%d %d %d %.17g %.17g %.17g %d %d %d" % (k + 1, mol, types[k], x[k, 0], x[k, 1], x[k, 2], im[k, 0], im[k, 1], im[k, 2])
                for k in range(n)]
        f.write("\n".join(rows))
        if "v" in system and system["v"] is not None:
            v = system["v"]
            f.write("\n\nVelocities\n\n")
            f.write("\n".join("%d %.17g %.17g %.17g" % (k + 1, v[k, 0], v[k, 1], v[k, 2]) for k in range(n)))
        if len(bt):
            f.write("\n\nBonds\n\n")
            f.write("\n".join("%d %d %d %d" % (k + 1, bt[k], a1[k], a2[k]) for k in range(len(bt))))
        f.write("\n")


def deck_header(system, datafile, skin=0.4, every=1, delay=1, check="yes", sort=False, comm_cutoff=5.0):
    """The reference deck of SURVEY.md Appendix B up to (not including) the fixes."""
    lines = ["units lj", "atom_style bond", "newton on off", "special_bonds fene"]
    if not sort:
        lines.append("atom_modify sort 0 0")
    lines += ["read_data %s" % datafile, "neighbor %g bin" % skin,
              "neigh_modify every %d delay %d check %s" % (every, delay, check)]
    if comm_cutoff:
        lines.append("comm_modify cutoff %g" % comm_cutoff)
    styles = sorted({s for s, _ in system["bond_coeffs"].values()})
    if len(styles) > 1:
        lines.append("bond_style hybrid " + " ".join(styles))
        for bt, (s, p) in sorted(system["bond_coeffs"].items()):
            lines.append("bond_coeff %d %s %s" % (bt, s, " ".join("%.17g" % q for q in p)))
    else:
        lines.append("bond_style " + styles[0])
        for bt, (s, p) in sorted(system["bond_coeffs"].items()):
            lines.append("bond_coeff %d %s" % (bt, " ".join("%.17g" % q for q in p)))
    lines += ["pair_style lj/cut 1.12246", "pair_modify shift yes", "pair_coeff * * 1.0 1.0 1.12246"]
    return lines


LMP_OMP = os.path.join(REF_DIR, "omp", "lmp_ref")     # the same sources + USER-OMP styles, built with -fopenmp


def have_threaded_reference():
    return os.path.exists(LMP_OMP) and os.path.exists(os.path.join(REF_DIR, "omp", "liblammps_ref.so"))


def run_reference(deck_lines, workdir=None, final=None, harness=True, timeout=3600, log=False, threads=1):
    """Run a deck through ref_harness (or lmp_ref); returns (stdout, path of the -final file or None).
    threads > 1 (harness=False only): the threaded build with `-sf omp -pk omp threads` -- pair, bond, neighbor and
    nve styles run in OpenMP threads, fix langevin and the USER-LE fixes stay serial (they have no /omp form)."""
    if not have_reference():
        raise RuntimeError("oracle/_ref is not built: run python oracle/build_ref.py")
    workdir = workdir or tempfile.mkdtemp(prefix="le_ref_")
    deck = os.path.join(workdir, "in.deck")
    with open(deck, "w") as f:
        f.write("\n".join(deck_lines) + "\n")
    omp = threads > 1 and not harness and have_threaded_reference()
    cmd = [HARNESS if harness else (LMP_OMP if omp else LMP), "-in", deck, "-log", "none" if not log else os.path.join(workdir, "log.lammps")]
    if omp:
        cmd += ["-sf", "omp", "-pk", "omp", str(threads)]
    if not log:
        cmd += ["-screen", os.path.join(workdir, "screen.txt")]
    if final and harness:
        cmd += ["-final", final]
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads) if omp else "1"
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True, timeout=timeout, env=env)
    out = r.stdout
    scr = os.path.join(workdir, "screen.txt")
    if os.path.exists(scr):
        out += open(scr).read()
    if r.returncode != 0:
        raise RuntimeError("reference run failed (%d):\n%s\n%s" % (r.returncode, out[-3000:], r.stderr[-3000:]))
    return out, final


def parse_loop_time(text):
    """'Loop time of T on P procs for S steps with N atoms' (src/finish.cpp) -> (T, P, S, N)."""
    m = re.search(r"Loop time of ([0-9.eE+-]+) on (\d+) procs for (\d+) steps with (\d+) atoms", text)
    if not m:
        return None
    return float(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))


def parse_thermo(text):
    """Thermo table(s) of a LAMMPS screen/log -> list of dict rows."""
    rows, cols = [], None
    for line in text.splitlines():
        t = line.split()
        if t and t[0] == "Step":
            cols = t
            continue
        if cols and len(t) == len(cols):
            try:
                rows.append({c: float(v) for c, v in zip(cols, t)})
                continue
            except ValueError:
                pass
        if cols and t and t[0] == "Loop":
            cols = None
    return rows
