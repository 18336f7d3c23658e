"""Shared helpers of the parity tests: build an engine from a reference record, compare states."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from lammps_le_b200.engine import Engine, LE_FIX_EXTRUSION, LE_FIX_EX_LOAD, LE_FIX_EX_UNLOAD  # noqa: E402

NEIGHMASK = 0x3FFFFFFF

# the USER-LE settings of the reference deck used for traces
LE_DECK = {
    "extrusion": dict(nevery=500, neutral=1, left=2, right=3, p_through=0.5, btype=2, roadblock=4, seed=12345),
    "ex_load": dict(nevery=100, itype=1, jtype=1, rc=1.12, btype=2, prob=0.02, seed=684474, iparam=(1, 1), jparam=(1, 1)),
    "ex_unload": dict(nevery=100, btype=2, rc=0.5, prob=0.05, seed=456456),
}


def le_deck_lines(cfg=LE_DECK, pre="pre.bin", post="post.bin"):
    x, l, u = cfg["extrusion"], cfg["ex_load"], cfg["ex_unload"]
    return [
        "fix s0 all le/snap %s pre grid" % pre,
        "fix loop all extrusion %d %d %d %d %g %d %d" % (x["nevery"], x["neutral"], x["left"], x["right"], x["p_through"], x["btype"], x["roadblock"]),
        "fix loading all ex_load %d %d %d %g %d prob %g %d iparam %d %d jparam %d %d" % (
            l["nevery"], l["itype"], l["jtype"], l["rc"], l["btype"], l["prob"], l["seed"], l["iparam"][0], l["iparam"][1], l["jparam"][0], l["jparam"][1]),
        "fix unloading all ex_unload %d %d %g prob %g %d" % (u["nevery"], u["btype"], u["rc"], u["prob"], u["seed"]),
        "fix s1 all le/snap %s post" % post,
    ]


def engine_from_record(rec, bond_coeffs, skin=0.4, every=1, delay=1, check=1, positions="xhold", device=0):
    """Engine holding exactly the reference state of a harness record (positions = xhold or x)."""
    e = Engine(rec["boxlo"], rec["boxhi"], (1, 1, 1), device)
    ntypes = int(max(4, rec["type"].max()))
    nbt = max(bond_coeffs)
    e.set_types(np.ones(ntypes), nbt)
    e.set_pair_lj(1.0, 1.0, 1.12246, shift=True)
    for bt, (style, params) in bond_coeffs.items():
        e.set_bond(bt, style, params)
    e.set_special((0.0, 1.0, 1.0))
    e.set_newton(1, 0)
    e.set_neighbor(skin, every, delay, check)
    e.set_capacity(rec["bpa"], rec["maxspecial"])
    e.upload_atoms(rec["type"], rec[positions], rec.get("v"), rec["image"])
    e.upload_topology(rec["num_bond"], rec["bond_type"], rec["bond_atom"], rec["nspecial"], rec["special"])
    return e


def define_le_fixes(e, cfg=LE_DECK):
    x, l, u = cfg["extrusion"], cfg["ex_load"], cfg["ex_unload"]
    e.fix_extrusion(x["nevery"], x["neutral"], x["left"], x["right"], x["p_through"], x["btype"], x["roadblock"], x["seed"])
    e.fix_ex_load(l["nevery"], l["itype"], l["jtype"], l["rc"], l["btype"], l["prob"], l["seed"], l["iparam"], l["jparam"])
    e.fix_ex_unload(u["nevery"], u["btype"], u["rc"], u["prob"], u["seed"])


def neigh_sets(offsets, entries):
    """per-atom frozenset of (partner tag, special bits) from a CSR list"""
    return [frozenset(int(v) for v in entries[offsets[t]:offsets[t + 1]]) for t in range(len(offsets) - 1)]


def compare_neighlists(off_a, ent_a, off_b, ent_b):
    """number of atoms whose neighbor sets differ + total entries"""
    a, b = neigh_sets(off_a, ent_a), neigh_sets(off_b, ent_b)
    bad = [t + 1 for t in range(len(a)) if a[t] != b[t]]
    return bad


def neighlists_equal_as_sets(off_a, ent_a, off_b, ent_b):
    """vectorised form of compare_neighlists for large systems: per-atom entry SETS are equal iff the (row, entry) pairs
    agree after sorting (a list never holds an entry twice); returns the number of differing pairs"""
    def pairs(off, ent):
        off = np.asarray(off, dtype=np.int64)
        rows = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off))
        key = rows * (1 << 32) + (np.asarray(ent).astype(np.int64) & 0xFFFFFFFF)
        return np.sort(key)
    a, b = pairs(off_a, ent_a), pairs(off_b, ent_b)
    if len(a) != len(b):
        return abs(len(a) - len(b)) + int(len(np.setxor1d(a, b)))
    return int((a != b).sum())


def special_tiers(nspecial, special):
    out = []
    for t in range(len(nspecial)):
        n1, n2, n3 = nspecial[t]
        row = special[t]
        out.append((frozenset(row[:n1].tolist()), frozenset(row[n1:n2].tolist()), frozenset(row[n2:n3].tolist())))
    return out


def compare_topology(got, ref):
    """dict of mismatch counts between an engine topology() dict and a reference record"""
    res = {}
    res["num_bond"] = int((got["num_bond"] != ref["num_bond"]).sum())
    bpa = ref["bond_type"].shape[1]
    mask = np.arange(bpa)[None, :] < ref["num_bond"][:, None]
    res["bond_type"] = int(((got["bond_type"] != ref["bond_type"]) & mask).sum())
    res["bond_atom"] = int(((got["bond_atom"] != ref["bond_atom"]) & mask).sum())
    res["nspecial"] = int((got["nspecial"] != ref["nspecial"]).any(axis=1).sum())
    ta, tb = special_tiers(got["nspecial"], got["special"]), special_tiers(ref["nspecial"], ref["special"])
    res["special_tiers"] = sum(1 for a, b in zip(ta, tb) if a != b)
    ms = ref["special"].shape[1]
    smask = np.arange(ms)[None, :] < ref["nspecial"][:, 2][:, None]
    res["special_exact"] = int(((got["special"] != ref["special"]) & smask).any(axis=1).sum())
    return res


WHICH = {1: LE_FIX_EXTRUSION, 2: LE_FIX_EX_UNLOAD, 3: LE_FIX_EX_LOAD}
RNG_SLOT = {1: 0, 2: 1, 3: 2}
SEED_KEY = {1: "extrusion", 2: "ex_unload", 3: "ex_load"}
