#!/usr/bin/env python3
"""BASELINE.json configs[0] as a fixture: the reference's own bench/data.chain (32,000-bead FENE melt) packed into
tests/golden/bench_chain.npz, plus the thermo lines of the reference's published log for bench/in.chain
(bench/log.6Oct16.chain.fixed.icc.1: step 0 and step 100) -- the golden numbers for this configuration.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):  python oracle/extract_bench_chain.py
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BENCH = "/root/reference/bench"


def main():
    lines = open(os.path.join(BENCH, "data.chain")).read().splitlines()
    n = int([l for l in lines if l.endswith("atoms")][0].split()[0])
    nb = int([l for l in lines if l.endswith("bonds")][0].split()[0])
    box = [tuple(float(v) for v in l.split()[:2]) for l in lines if l.endswith(("xlo xhi", "ylo yhi", "zlo zhi"))]
    ia, iv, ib = lines.index("Atoms"), lines.index("Velocities"), lines.index("Bonds")
    atoms = np.array([l.split() for l in lines[ia + 2:ia + 2 + n]], dtype=np.float64)
    vel = np.array([l.split() for l in lines[iv + 2:iv + 2 + n]], dtype=np.float64)
    bonds = np.array([l.split() for l in lines[ib + 2:ib + 2 + nb]], dtype=np.int64)
    o = np.argsort(atoms[:, 0]); atoms = atoms[o]
    vel = vel[np.argsort(vel[:, 0])]
    log = open(os.path.join(BENCH, "log.6Oct16.chain.fixed.icc.1")).read()
    m = re.search(r"Step Temp E_pair E_mol TotEng Press \n(.*?)\nLoop time", log, re.S)
    thermo = np.array([[float(v) for v in row.split()] for row in m.group(1).splitlines()])
    builds = int(re.search(r"Neighbor list builds = (\d+)", log).group(1))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bench_chain.npz"),
                        mol=atoms[:, 1].astype(np.int32), type=atoms[:, 2].astype(np.int32), x=atoms[:, 3:6], image=atoms[:, 6:9].astype(np.int32),
                        v=vel[:, 1:4], bonds=bonds[:, 1:4].astype(np.int32), boxlo=np.array([b[0] for b in box]), boxhi=np.array([b[1] for b in box]),
                        ref_thermo=thermo, ref_builds=np.array(builds))
    print("bench_chain.npz: %d atoms, %d bonds, reference thermo rows:\n%s" % (n, nb, thermo))


if __name__ == "__main__":
    main()
