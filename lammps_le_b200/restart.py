"""The reference's binary restart files for this path: `read_restart` / `write_restart` (src/read_restart.cpp, src/write_restart.cpp)
over the host-only entry points le_host_restart_* (csrc/le_restart.cpp, in libleb200_host.so: no GPU needed to read or write).

    info, atoms = read_restart("chain.restart")        # header (dict) + per-atom arrays in TAG order
    e = engine_from_restart(info, atoms)                 # a configured Engine with the state uploaded (special lists rebuilt)
    write_restart("out.restart", e)                      # a file the reference's read_restart accepts
"""
import ctypes as C

import numpy as np

from .engine import LeError, _pd, _pi, load_host_library

MAXT = 8


class RestartHeader(C.Structure):
    _fields_ = [("ntimestep", C.c_int64), ("natoms", C.c_int64), ("nbonds", C.c_int64),
                ("ntypes", C.c_int), ("nbondtypes", C.c_int), ("bond_per_atom", C.c_int), ("extra_bond_per_atom", C.c_int), ("maxspecial", C.c_int),
                ("newton_pair", C.c_int), ("newton_bond", C.c_int), ("periodic", C.c_int * 3), ("nprocs_file", C.c_int), ("atom_sortfreq", C.c_int),
                ("boxlo", C.c_double * 3), ("boxhi", C.c_double * 3), ("special_lj", C.c_double * 3), ("dt", C.c_double), ("comm_cutoff", C.c_double),
                ("mass", C.c_double * MAXT),
                ("version", C.c_char * 32), ("units", C.c_char * 16), ("atom_style", C.c_char * 16), ("pair_style", C.c_char * 32), ("bond_style", C.c_char * 32),
                ("cut_global", C.c_double), ("offset_flag", C.c_int), ("mix_flag", C.c_int), ("tail_flag", C.c_int),
                ("pair_setflag", C.c_int * (MAXT * MAXT)),
                ("pair_eps", C.c_double * (MAXT * MAXT)), ("pair_sigma", C.c_double * (MAXT * MAXT)), ("pair_cut", C.c_double * (MAXT * MAXT)),
                ("bond_coeffs_stored", C.c_int),
                ("bond_k", C.c_double * MAXT), ("bond_r0", C.c_double * MAXT), ("bond_eps", C.c_double * MAXT), ("bond_sigma", C.c_double * MAXT),
                ("nhybrid", C.c_int), ("hybrid_styles", (C.c_char * 16) * 4)]


def _lib():
    lib = load_host_library()
    if not hasattr(lib, "_restart_ready"):
        pi, pd, cs = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_char_p
        lib.le_host_restart_read_header.argtypes = [cs, C.POINTER(RestartHeader), cs, C.c_int]
        lib.le_host_restart_read_atoms.argtypes = [cs, pi, pi, pi, pi, pd, pd, pi, pi, pi, cs, C.c_int]
        lib.le_host_restart_write.argtypes = [cs, C.POINTER(RestartHeader), pi, pi, pi, pi, pd, pd, pi, pi, pi, cs, C.c_int]
        lib._restart_ready = True
    return lib


def _header_dict(h):
    nt, nbt = h.ntypes, h.nbondtypes
    d = {k: getattr(h, k) for k in ("ntimestep", "natoms", "nbonds", "ntypes", "nbondtypes", "bond_per_atom", "extra_bond_per_atom", "maxspecial",
                                     "newton_pair", "newton_bond", "dt", "comm_cutoff", "cut_global", "offset_flag", "mix_flag", "tail_flag",
                                     "bond_coeffs_stored", "atom_sortfreq", "nprocs_file")}
    for k in ("version", "units", "atom_style", "pair_style", "bond_style"):
        d[k] = getattr(h, k).decode()
    d["periodic"] = list(h.periodic)
    d["boxlo"], d["boxhi"], d["special_lj"] = np.array(h.boxlo), np.array(h.boxhi), np.array(h.special_lj)
    d["mass"] = np.array(h.mass[:nt])
    M = lambda a: np.array(a[:]).reshape(MAXT, MAXT)[:nt, :nt] if False else np.array([[a[i * nt + j] for j in range(nt)] for i in range(nt)])
    d["pair_setflag"], d["pair_eps"], d["pair_sigma"], d["pair_cut"] = M(h.pair_setflag), M(h.pair_eps), M(h.pair_sigma), M(h.pair_cut)
    d["bond_k"], d["bond_r0"] = np.array(h.bond_k[:nbt]), np.array(h.bond_r0[:nbt])
    d["bond_eps"], d["bond_sigma"] = np.array(h.bond_eps[:nbt]), np.array(h.bond_sigma[:nbt])
    d["hybrid_styles"] = [bytes(h.hybrid_styles[m]).split(b"\0")[0].decode() for m in range(h.nhybrid)]
    return d


def read_restart(path):
    """(header dict, atoms dict in tag order: tag type image molecule x v num_bond bond_type bond_atom)"""
    lib = _lib()
    h = RestartHeader()
    err = C.create_string_buffer(256)
    if lib.le_host_restart_read_header(path.encode(), C.byref(h), err, 256):
        raise LeError(-1, err.value.decode())
    n, bpa = int(h.natoms), int(h.bond_per_atom)
    a = {"tag": np.zeros(n, np.int32), "type": np.zeros(n, np.int32), "image": np.zeros(n, np.int32), "molecule": np.zeros(n, np.int32),
         "x": np.zeros((n, 3)), "v": np.zeros((n, 3)), "num_bond": np.zeros(n, np.int32),
         "bond_type": np.zeros((n, max(bpa, 1)), np.int32), "bond_atom": np.zeros((n, max(bpa, 1)), np.int32)}
    if lib.le_host_restart_read_atoms(path.encode(), _pi(a["tag"]), _pi(a["type"]), _pi(a["image"]), _pi(a["molecule"]), _pd(a["x"]), _pd(a["v"]),
                                      _pi(a["num_bond"]), _pi(a["bond_type"]), _pi(a["bond_atom"]), err, 256):
        raise LeError(-1, err.value.decode())
    order = np.argsort(a["tag"], kind="stable")
    if not np.array_equal(a["tag"][order], np.arange(1, n + 1)):
        raise LeError(-1, "restart file: atom ids are not 1..N")
    return _header_dict(h), {k: np.ascontiguousarray(v[order]) for k, v in a.items()}


def header_for(info):
    """RestartHeader from a header dict (the keys read_restart returns)"""
    h = RestartHeader()
    nt, nbt = int(info["ntypes"]), int(info["nbondtypes"])
    for k in ("ntimestep", "natoms", "nbonds", "ntypes", "nbondtypes", "bond_per_atom", "extra_bond_per_atom", "maxspecial", "newton_pair", "newton_bond",
              "dt", "comm_cutoff", "cut_global", "offset_flag", "mix_flag", "tail_flag", "atom_sortfreq"):
        setattr(h, k, type(getattr(h, k))(info.get(k, 0)))
    for k in ("version", "units", "atom_style", "pair_style", "bond_style"):
        setattr(h, k, str(info.get(k, "")).encode())
    for q in range(3):
        h.periodic[q] = int(info.get("periodic", (1, 1, 1))[q]); h.boxlo[q] = float(info["boxlo"][q]); h.boxhi[q] = float(info["boxhi"][q])
        h.special_lj[q] = float(info["special_lj"][q])
    for t in range(nt):
        h.mass[t] = float(info["mass"][t])
        for u in range(nt):
            k = t * nt + u
            h.pair_setflag[k] = int(info["pair_setflag"][t][u]); h.pair_eps[k] = float(info["pair_eps"][t][u])
            h.pair_sigma[k] = float(info["pair_sigma"][t][u]); h.pair_cut[k] = float(info["pair_cut"][t][u])
    for t in range(nbt):
        h.bond_k[t], h.bond_r0[t] = float(info["bond_k"][t]), float(info["bond_r0"][t])
        h.bond_eps[t], h.bond_sigma[t] = float(info["bond_eps"][t]), float(info["bond_sigma"][t])
    hs = info.get("hybrid_styles", [])
    h.nhybrid = len(hs)
    for m, name in enumerate(hs):
        h.hybrid_styles[m].value = name.encode()
    return h


def write_restart_arrays(path, info, atoms):
    """write header dict + per-atom arrays (any order; each bond on both atoms under newton_bond off, as Atom holds them)"""
    lib = _lib()
    h = header_for(info)
    err = C.create_string_buffer(256)
    bpa = int(info["bond_per_atom"])
    i32 = lambda k: np.ascontiguousarray(atoms[k], dtype=np.int32)
    bt = np.zeros((len(atoms["tag"]), max(bpa, 1)), np.int32); ba = np.zeros_like(bt)
    bt[:, :atoms["bond_type"].shape[1]] = atoms["bond_type"][:, :bpa]; ba[:, :atoms["bond_atom"].shape[1]] = atoms["bond_atom"][:, :bpa]
    x, v = np.ascontiguousarray(atoms["x"], dtype=np.float64), np.ascontiguousarray(atoms["v"], dtype=np.float64)
    tag, typ, img, mol, nb = i32("tag"), i32("type"), i32("image"), i32("molecule"), i32("num_bond")
    if lib.le_host_restart_write(path.encode(), C.byref(h), _pi(tag), _pi(typ), _pi(img), _pi(mol), _pd(x), _pd(v), _pi(nb), _pi(bt), _pi(ba), err, 256):
        raise LeError(-1, err.value.decode())


def engine_from_restart(info, atoms, device=0, bond_coeffs=None, skin=0.4, every=1, delay=0, check=1):
    """Engine holding the restart's state.  bond_coeffs = {btype: (style, params)} is needed when the file is a `bond_style hybrid`
    one (the reference stores only the sub-style names and asks for bond_coeff again, BondHybrid::write_restart)."""
    from .engine import Engine
    e = Engine(info["boxlo"], info["boxhi"], tuple(info["periodic"]), device)
    e.set_types(info["mass"], info["nbondtypes"])
    nt = info["ntypes"]
    # PairLJCut::init_one mixes the pairs that were not set explicitly (geometric by default, src/pair.cpp:205-240)
    eps, sig, cut = (np.array(info[k], dtype=np.float64) for k in ("pair_eps", "pair_sigma", "pair_cut"))
    for i in range(nt):
        for j in range(i + 1, nt):
            if not info["pair_setflag"][i][j]:
                eps[i, j], sig[i, j], cut[i, j] = np.sqrt(eps[i, i] * eps[j, j]), np.sqrt(sig[i, i] * sig[j, j]), max(cut[i, i], cut[j, j])
            eps[j, i], sig[j, i], cut[j, i] = eps[i, j], sig[i, j], cut[i, j]
    e.set_pair_lj(eps, sig, cut, shift=bool(info["offset_flag"]))
    for t in range(info["nbondtypes"]):
        if bond_coeffs and (t + 1) in bond_coeffs:
            e.set_bond(t + 1, *bond_coeffs[t + 1])
        elif info["bond_style"] == "fene":
            e.set_bond(t + 1, "fene", (info["bond_k"][t], info["bond_r0"][t], info["bond_eps"][t], info["bond_sigma"][t]))
        elif info["bond_style"] == "harmonic":
            e.set_bond(t + 1, "harmonic", (info["bond_k"][t], info["bond_r0"][t]))
        else:
            raise LeError(-1, "All bond coeffs are not set")          # the reference's message for a hybrid restart without bond_coeff
    e.set_special(tuple(info["special_lj"]))
    e.set_newton(info["newton_pair"], info["newton_bond"])
    e.set_neighbor(skin, every, delay, check)
    e.set_capacity(info["bond_per_atom"], max(info["maxspecial"], 1))
    e.set_timestep(info["dt"])
    e.upload_atoms(atoms["type"], atoms["x"], atoms["v"], atoms["image"])
    if info["newton_bond"]:
        t, m = np.nonzero(np.arange(atoms["bond_type"].shape[1])[None, :] < atoms["num_bond"][:, None])
        e.upload_bonds(atoms["bond_type"][t, m], (t + 1).astype(np.int32), atoms["bond_atom"][t, m])
    else:
        e.lib.le_upload_topology(e._h, _pi(atoms["num_bond"]), _pi(np.ascontiguousarray(atoms["bond_type"])),
                                 _pi(np.ascontiguousarray(atoms["bond_atom"])), None, None) and e._ck(-1)
    e.reset_timestep(info["ntimestep"])
    return e
