"""Synthetic workloads of BASELINE.json and the standard engine set-up for them.

A "system" is a plain dict:
  box (lo, hi), types[N], x[N,3], image[N], bonds = (btype[], a[], b[]), masses, nbondtypes
plus the force-field settings of the chromatin model:
  pair lj/cut 1.12246 (WCA, shift yes), bond 1 = FENE(30, 1.5, 1, 1), bond 2 = harmonic(20, 1.3),
  special_bonds fene, newton on off, neighbor 0.4 bin.
"""
import numpy as np

from .engine import Engine, gen_lattice_melt, gen_saw_chains

WCA_CUT = 1.12246
NEUTRAL, LEFT, RIGHT, ROADBLOCK = 1, 2, 3, 4


EXTRUDER_HARMONIC = ("harmonic", (20.0, 1.3))        # the deck of SURVEY.md Appendix B (parity fixtures)
EXTRUDER_FENE = ("fene", (10.0, 4.0, 1.0, 1.0))        # same WCA core as the pair potential: no overlap while
                                                       # bonded, so unloading never switches WCA on at small r


def chromatin_chain(n, n_extruders, rho=0.2, seed=12345, barriers="periodic", nchains=1,
                    p_left=0.005, p_right=0.005, p_block=0.001, extruder_bond=EXTRUDER_HARMONIC):
    """One (or nchains) self-avoiding chain(s) of n beads at number density rho with pre-placed extruders.

    barriers="periodic": CTCF every 100 beads, left at i%100==50, right at i%100==75 (config C2);
    barriers="random":   i.i.d. per bead P(left)=P(right)=0.005, P(roadblock)=0.001, seed 2024 (config C4).
    Extruder bonds (type 2) join beads (i, i+2) at non-overlapping random i.
    """
    L = (n / rho) ** (1.0 / 3.0)
    x, image = gen_saw_chains(n, nchains, L, 0.97, 0.9, seed)
    types = np.full(n, NEUTRAL, dtype=np.int32)
    idx = np.arange(n)
    if barriers == "periodic":
        types[idx % 100 == 50] = LEFT
        types[idx % 100 == 75] = RIGHT
    elif barriers == "random":
        r = np.random.default_rng(2024).random(n)
        types[r < p_left] = LEFT
        types[(r >= p_left) & (r < p_left + p_right)] = RIGHT
        types[(r >= p_left + p_right) & (r < p_left + p_right + p_block)] = ROADBLOCK
    clen = n // nchains
    a = np.arange(1, n + 1)
    keep = (a % clen) != 0          # no bond across chain ends
    b1 = a[keep][a[keep] < n]
    btype = [np.ones(len(b1), dtype=np.int32)]
    at1, at2 = [b1.astype(np.int32)], [(b1 + 1).astype(np.int32)]
    if n_extruders > 0:
        rng = np.random.default_rng(seed)
        # anchors i (1-based tag) with i, i+1, i+2 interior beads of one chain, spaced >= 4 apart
        slots = np.arange(2, n - 3, 4)
        pos_in_chain = (slots - 1) % clen
        slots = slots[(pos_in_chain >= 1) & (pos_in_chain <= clen - 4)]
        pick = np.sort(rng.choice(slots, size=min(n_extruders, len(slots)), replace=False))
        btype.append(np.full(len(pick), 2, dtype=np.int32))
        at1.append(pick.astype(np.int32))
        at2.append((pick + 2).astype(np.int32))
    return {
        "name": "chromatin_%d" % n, "box": (np.zeros(3), np.full(3, L)), "types": types, "x": x, "image": image,
        "bonds": (np.concatenate(btype), np.concatenate(at1), np.concatenate(at2)),
        "masses": np.ones(4), "nbondtypes": 2, "ntypes": 4,
        "bond_coeffs": {1: ("fene", (30.0, 1.5, 1.0, 1.0)), 2: extruder_bond},
        "bond_per_atom": 4, "maxspecial": 46,
    }


def fene_melt(nchains=320, length=100, rho=0.8442):
    """bench/in.chain-like melt (320 x 100 at rho* = 0.8442) started from a lattice snake path."""
    x, image, L = gen_lattice_melt(nchains, length, rho)
    n = nchains * length
    a = np.arange(1, n + 1)
    b1 = a[(a % length) != 0]
    return {
        "name": "melt_%d" % n, "box": (np.zeros(3), np.full(3, L)), "types": np.ones(n, dtype=np.int32), "x": x,
        "image": image, "bonds": (np.ones(len(b1), dtype=np.int32), b1.astype(np.int32), (b1 + 1).astype(np.int32)),
        "masses": np.ones(1), "nbondtypes": 1, "ntypes": 1,
        "bond_coeffs": {1: ("fene", (30.0, 1.5, 1.0, 1.0))},
        "bond_per_atom": 2, "maxspecial": 8,
    }


def maxwell_velocities(n, temp, masses_per_atom, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 3)) * np.sqrt(temp / masses_per_atom)[:, None]
    v -= v.mean(axis=0)
    return v


def make_engine(system, device=0, skin=0.4, every=1, delay=1, check=1, dt=0.005, maxneigh=None, velocities=None, dd=None):
    """Engine configured like the reference deck of SURVEY.md Appendix B for `system`.
    dd = dict(rank=, world=, halo=, group=): one slab of a multi-GPU run (engine_dd.DDEngine)."""
    lo, hi = system["box"]
    if dd:
        from .engine_dd import DDEngine
        e = DDEngine(lo, hi, (1, 1, 1), device, **dd)
    else:
        e = Engine(lo, hi, (1, 1, 1), device)
    e.set_types(system["masses"], system["nbondtypes"])
    e.set_pair_lj(1.0, 1.0, WCA_CUT, shift=True)
    for bt, (style, params) in system["bond_coeffs"].items():
        e.set_bond(bt, style, params)
    e.set_special((0.0, 1.0, 1.0))          # special_bonds fene
    e.set_newton(1, 0)                      # newton on off
    e.set_neighbor(skin, every, delay, check)
    e.set_capacity(system["bond_per_atom"], system["maxspecial"])
    if maxneigh:
        e.set_neighbor_capacity(maxneigh)
    e.set_timestep(dt)
    e.upload_atoms(system["types"], system["x"], velocities, system["image"])
    bt, a1, a2 = system["bonds"]
    e.upload_bonds(bt, a1, a2)
    return e


def relax(e, steps=2000, xmax=0.05, temp=1.0, damp=1.0, seed=4711):
    """Push-off run (fix nve/limit + fix langevin), the usual preparation of a generated polymer start."""
    e.fix_nve_limit(xmax)
    e.fix_langevin(temp, temp, damp, seed)
    e.run(steps)
    e.fix_nve(True)


def replicate(system, nx, ny, nz):
    """`replicate nx ny nz` (src/replicate.cpp) of a bonded system: images of the box side by side, bonds copied with
    the atom ids shifted; used for the weak-scaling melt of bench/in.chain.scaled"""
    lo, hi = (np.asarray(a, dtype=np.float64) for a in system["box"])
    L = hi - lo
    n = len(system["types"])
    bt, a1, a2 = system["bonds"]
    xs, ts, ims, bts, b1s, b2s = [], [], [], [], [], []
    from .engine import unpack_image, pack_image
    img = unpack_image(system["image"])
    k = 0
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                # a bond may cross the original box: unwrap, shift, wrap into the big box
                xu = system["x"] + img * L + np.array([ix, iy, iz]) * L
                big = L * np.array([nx, ny, nz])
                w = np.floor((xu - lo) / big)
                xs.append(xu - w * big)
                ims.append(pack_image(w.astype(np.int64)))
                ts.append(system["types"])
                bts.append(bt); b1s.append(a1 + k * n); b2s.append(a2 + k * n)
                k += 1
    out = dict(system)
    out.update(name="%s_x%d" % (system.get("name", "sys"), nx * ny * nz), box=(lo, lo + L * np.array([nx, ny, nz])),
               types=np.concatenate(ts), x=np.concatenate(xs), image=np.concatenate(ims),
               bonds=(np.concatenate(bts), np.concatenate(b1s).astype(np.int32), np.concatenate(b2s).astype(np.int32)))
    if system.get("v") is not None:
        out["v"] = np.tile(system["v"], (nx * ny * nz, 1))
    return out
