#!/usr/bin/env python3
"""Extract the reference's own unit-test vectors for the styles on the hot path into small fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):
    python oracle/extract_ref_yaml.py
reads  unittest/force-styles/tests/{mol-pair-lj_cut,bond-fene,bond-harmonic,angle-cosine}.yaml + in.fourmol/data.fourmol
writes tests/golden/ref_yaml_<name>.npz  (inputs in tag order + the yaml's init_forces / energies / stress)
       tests/golden/ranmars_ref.npz      (first draws of RanMars for the seeds the decks use, from oracle/_ref)
The numbers are the reference's published known answers (yaml `epsilon` 5e-14 / 2.5e-13); nothing is recomputed here.
"""
import os
import subprocess
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TESTS = "/root/reference/unittest/force-styles/tests"
GOLD = os.path.join(ROOT, "tests", "golden")


def read_fourmol():
    lines = open(os.path.join(TESTS, "data.fourmol")).read().splitlines()
    box = {}
    sect = {}
    cur = None
    for ln in lines[1:]:
        t = ln.split("#")[0].split()
        if not t:
            continue
        if len(t) == 4 and t[2] in ("xlo", "ylo", "zlo"):
            box[t[2][0]] = (float(t[0]), float(t[1]))
            continue
        if t[0][0].isalpha():
            cur = " ".join(t)
            sect[cur] = []
            continue
        if cur is not None:
            sect[cur].append(t)
    atoms = sorted(sect["Atoms"], key=lambda r: int(r[0]))
    n = len(atoms)
    x = np.array([[float(r[4]), float(r[5]), float(r[6])] for r in atoms])
    typ = np.array([int(r[2]) for r in atoms], dtype=np.int32)
    bonds = [(int(r[1]), int(r[2]), int(r[3])) for r in sect["Bonds"]]
    angles = np.array([[int(r[1]), int(r[2]), int(r[3]), int(r[4])] for r in sect["Angles"]], dtype=np.int32)   # type a1 a2 a3
    bpa = 6
    nb = np.zeros(n, np.int32); bt = np.zeros((n, bpa), np.int32); ba = np.zeros((n, bpa), np.int32)
    for t, a, b in bonds:       # newton_bond off layout: every bond on both atoms (Atom::data_bonds src/atom.cpp:1261-1278)
        for p, q in ((a, b), (b, a)):
            bt[p - 1, nb[p - 1]] = t; ba[p - 1, nb[p - 1]] = q; nb[p - 1] += 1
    boxlo = np.array([box[k][0] for k in "xyz"]); boxhi = np.array([box[k][1] for k in "xyz"])
    return dict(x=x, type=typ, num_bond=nb, bond_type=bt, bond_atom=ba, boxlo=boxlo, boxhi=boxhi, angles=angles)


def block(y, key, cols):
    rows = [[float(v) for v in ln.split()] for ln in y[key].strip().splitlines()]
    a = np.array(rows)
    return a[:, -cols:] if a.shape[1] > cols else a


def main():
    os.makedirs(GOLD, exist_ok=True)
    base = read_fourmol()
    # pair lj/cut
    y = yaml.safe_load(open(os.path.join(TESTS, "mol-pair-lj_cut.yaml")))
    assert y["input_file"] == "in.fourmol" and y["pair_style"].split() == ["lj/cut", "8.0"]
    nt = int(base["type"].max())
    eps = np.zeros((nt, nt)); sig = np.zeros((nt, nt)); have = np.zeros((nt, nt), bool)
    for ln in y["pair_coeff"].strip().splitlines():
        a, b, e, s = ln.split()
        a, b = int(a) - 1, int(b) - 1
        eps[a, b] = eps[b, a] = float(e); sig[a, b] = sig[b, a] = float(s); have[a, b] = have[b, a] = True
    assert "mix arithmetic" in y["post_commands"]
    for a in range(nt):
        for b in range(nt):
            if not have[a, b]:      # Pair::mix_energy / mix_distance, arithmetic (src/pair.cpp:577-607)
                eps[a, b] = np.sqrt(eps[a, a] * eps[b, b]); sig[a, b] = 0.5 * (sig[a, a] + sig[b, b])
    np.savez_compressed(os.path.join(GOLD, "ref_yaml_mol-pair-lj_cut.npz"), **base, epsilon=eps, sigma=sig,
                        cut=np.full((nt, nt), 8.0), special_lj=np.array([1.0, 0.10, 0.25, 0.50]),   # in.fourmol special_bonds
                        init_forces=block(y, "init_forces", 3), init_vdwl=np.array(float(y["init_vdwl"])),
                        init_stress=block(y, "init_stress", 6)[0], yaml_epsilon=np.array(float(y["epsilon"])))
    for name in ("bond-fene", "bond-harmonic"):
        y = yaml.safe_load(open(os.path.join(TESTS, name + ".yaml")))
        assert y["input_file"] == "in.fourmol"
        coeff = np.array([[float(v) for v in ln.split()[1:]] for ln in y["bond_coeff"].strip().splitlines()])
        np.savez_compressed(os.path.join(GOLD, "ref_yaml_%s.npz" % name), **base, bond_coeff=coeff,
                            init_forces=block(y, "init_forces", 3), init_energy=np.array(float(y["init_energy"])),
                            init_stress=block(y, "init_stress", 6)[0], yaml_epsilon=np.array(float(y["epsilon"])))
    # angle cosine (chain stiffness)
    y = yaml.safe_load(open(os.path.join(TESTS, "angle-cosine.yaml")))
    assert y["input_file"] == "in.fourmol" and y["angle_style"] == "cosine"
    coeff = np.array([[float(v) for v in ln.split()[1:]] for ln in y["angle_coeff"].strip().splitlines()])
    np.savez_compressed(os.path.join(GOLD, "ref_yaml_angle-cosine.npz"), **base, angle_coeff=coeff,
                        init_forces=block(y, "init_forces", 3), init_energy=np.array(float(y["init_energy"])),
                        init_stress=block(y, "init_stress", 6)[0], yaml_epsilon=np.array(float(y["epsilon"])))
    # RanMars known answers from the compiled reference
    harness = os.path.join(HERE, "_ref", "ref_harness")
    seeds = [12345, 684474, 456456, 904297, 1, 900000000]
    draws = []
    for s in seeds:
        out = subprocess.run([harness, "-ranmars", str(s), "300", "-log", "none", "-screen", "none"], capture_output=True, text=True, check=True).stdout
        draws.append([float.fromhex(ln.split()[1]) for ln in out.splitlines() if ln.startswith("RANMARS")])
    np.savez_compressed(os.path.join(GOLD, "ranmars_ref.npz"), seeds=np.array(seeds), draws=np.array(draws))
    print("wrote", sorted(f for f in os.listdir(GOLD) if f.startswith("ref_yaml") or f.startswith("ranmars")))


if __name__ == "__main__":
    sys.exit(main())
