"""ctypes binding of include/le_b200.h -- the Python mirror of the C ABI.

Every method maps one-to-one onto an `le_*` entry point (same names, argument meaning and error
behaviour); errors are raised as `LeError` carrying `le_last_error()`, the way the reference raises
`LAMMPSException` (src/library.h:236-237).  There is no fallback: if the CUDA library is missing or
no GPU is present the constructor raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libleb200.so")

LE_FIX_EXTRUSION, LE_FIX_EX_UNLOAD, LE_FIX_EX_LOAD = 1, 2, 3
LE_BOND_NONE, LE_BOND_FENE, LE_BOND_HARMONIC = 0, 1, 2
SBBITS = 30
NEIGHMASK = 0x3FFFFFFF


class LeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("le_b200 error %d: %s" % (code, msg))
        self.code = code


class Thermo(C.Structure):
    _fields_ = [("step", C.c_int64), ("temp", C.c_double), ("epair", C.c_double), ("emol", C.c_double),
                ("etotal", C.c_double), ("press", C.c_double), ("ke", C.c_double), ("virial", C.c_double * 6),
                ("nbonds", C.c_int64), ("fene_warnings", C.c_int64), ("le_f1", C.c_int64 * 3), ("le_f2", C.c_int64 * 3), ("eangle", C.c_double)]

    def as_dict(self):
        return {"step": self.step, "temp": self.temp, "epair": self.epair, "emol": self.emol,
                "etotal": self.etotal, "press": self.press, "ke": self.ke, "virial": list(self.virial),
                "nbonds": self.nbonds, "fene_warnings": self.fene_warnings, "le_f1": list(self.le_f1), "le_f2": list(self.le_f2),
                "eangle": self.eangle}


class MinResult(C.Structure):
    _fields_ = [("stop", C.c_int), ("niter", C.c_int), ("neval", C.c_int), ("einitial", C.c_double), ("eprevious", C.c_double),
                ("efinal", C.c_double), ("fnorm2_init", C.c_double), ("fnorm2_final", C.c_double), ("fnorminf_init", C.c_double),
                ("fnorminf_final", C.c_double), ("alpha_final", C.c_double)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("steps", "neigh_builds", "dangerous_builds", "half_pairs", "full_entries", "kernel_launches",
                 "extrusion_shifts", "loads", "unloads", "last_extrusion_shifts", "last_loads", "last_unloads")] + \
               [("last_run_gpu_ms", C.c_double), ("last_run_le_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def load_library():
    """dlopen the CUDA library; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("lammps_le_b200: %s is missing -- the CUDA extension is not built; "
                          "run `make -C lammps_le_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P, I, D, I64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "le_create": [C.POINTER(P), I, pd, pd, pi], "le_set_types": [P, I, pd, I],
        "le_set_pair_lj": [P, I, pd, pd, pd, I], "le_set_bond": [P, I, I, pd], "le_set_special": [P, pd], "le_set_angle_types": [P, I], "le_set_angle": [P, I, I, pd], "le_upload_angles": [P, I, pi, pi, pi, pi],
        "le_set_neighbor": [P, D, I, I, I], "le_set_neighbor_capacity": [P, I], "le_set_newton": [P, I, I],
        "le_set_capacity": [P, I, I], "le_set_timestep": [P, D], "le_reset_timestep": [P, I64],
        "le_thermo_every": [P, I], "le_fix_nve": [P, I], "le_fix_nve_limit": [P, D],
        "le_fix_langevin": [P, D, D, D, I], "le_fix_extrusion": [P, I, I, I, I, D, I, I, I],
        "le_fix_ex_load": [P, I, I, I, D, I, D, I, I, I, I, I], "le_fix_ex_unload": [P, I, I, D, D, I], "le_fix_bond_break": [P, I, I, D, D, I],
        "le_fix_bond_create": [P, I, I, I, D, I, D, I, I, I, I, I],
        "le_unfix": [P, I], "le_upload_atoms": [P, I, pi, pi, pd, pd, pi], "le_upload_bonds": [P, I, pi, pi, pi],
        "le_upload_topology": [P, pi, pi, pi, pi, pi], "le_set_positions": [P, pd, pi], "le_set_velocities": [P, pd],
        "le_run": [P, I64], "le_set_run_span": [P, I64, I64], "le_run_timed": [P, I64, pd], "le_force_rebuild": [P], "le_run_le_event": [P, I],
        "le_fix_rng_reset": [P, I, I, I64], "le_fix_rng_consumed": [P, I, C.POINTER(I64)],
        "le_fix_rng_set_state": [P, I, pd], "le_fix_rng_get_state": [P, I, pd],
        "le_minimize": [P, D, D, I, I, C.POINTER(MinResult)],
        "le_compute_forces": [P, pd, C.POINTER(Thermo)], "le_compute_forces_plain": [P, pd], "le_natoms": [P], "le_download_x": [P, pd, pi],
        "le_download_v": [P, pd], "le_download_types": [P, pi], "le_download_topology": [P, pi, pi, pi, pi, pi],
        "le_download_neighlist": [P, I, C.POINTER(I64), pi, C.POINTER(I64)],
        "le_download_bondlist": [P, pi, C.POINTER(I64)], "le_thermo_count": [P],
        "le_get_thermo": [P, I, C.POINTER(Thermo)], "le_get_stats": [P, C.POINTER(Stats)], "le_compute_rg": [P, pd],
        "le_gen_saw_chains": [I, I, D, D, D, C.c_uint64, pd, pi], "le_gen_lattice_melt": [I, I, D, pd, pd, pi],
        "le_observables": [P, I, pi, D, I, I, I, pd, C.POINTER(I64), C.POINTER(I64)],
        "le_local_capacity": [P], "le_download_owned": [P, pi, pi, pd, pi, pd], "le_upload_owned": [P, I, pi, pd, pi, pd],
        "le_dd_init": [P, I, I, D], "le_dd_balance": [P, I], "le_dd_get_handle": [P, C.c_void_p], "le_dd_connect": [P, C.c_void_p],
        "le_get_thermo_sums": [P, I, pd], "le_get_force_sums": [P, pd],
        "le_host_velocity_create": [I, pi, pd, pd, D, I, I, I, I, pd],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = I
    lib.le_destroy.argtypes = [P]
    lib.le_destroy.restype = None
    lib.le_last_error.argtypes = [P]
    lib.le_last_error.restype = C.c_char_p
    lib.le_version.restype = C.c_char_p
    lib.le_host_property_local_bonds.argtypes = [I, I, pi, pi, pi, I, pi]
    lib.le_host_property_local_bonds.restype = I64
    lib.le_min_stop_string.argtypes = [I]
    lib.le_min_stop_string.restype = C.c_char_p
    lib.le_step_kernel_name.argtypes = [P]
    lib.le_step_kernel_name.restype = C.c_char_p
    lib.le_timestep.argtypes = [P]
    lib.le_timestep.restype = I64
    _lib = lib
    return lib


_hostlib = None


def load_host_library():
    """the host-only helpers (le_gen_*, le_host_*): libleb200_host.so (plain g++, no CUDA) so that making synthetic inputs
    does not load the CUDA engine; falls back to the full library when only that has been built"""
    global _hostlib
    if _hostlib is not None:
        return _hostlib
    path = os.path.join(_HERE, "libleb200_host.so")
    if not os.path.exists(path):
        _hostlib = load_library()
        return _hostlib
    lib = C.CDLL(path)
    I, D, I64 = C.c_int, C.c_double, C.c_int64
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.le_gen_saw_chains.argtypes = [I, I, D, D, D, C.c_uint64, pd, pi]
    lib.le_gen_lattice_melt.argtypes = [I, I, D, pd, pd, pi]
    lib.le_host_velocity_create.argtypes = [I, pi, pd, pd, D, I, I, I, I, pd]
    lib.le_host_property_local_bonds.argtypes = [I, I, pi, pi, pi, I, pi]
    lib.le_host_property_local_bonds.restype = I64
    for n in ("le_gen_saw_chains", "le_gen_lattice_melt", "le_host_velocity_create"):
        getattr(lib, n).restype = I
    _hostlib = lib
    return lib


def _pd(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _journaled(fn):
    """record a settings / fix call so that the context can be rebuilt with the same script (DDEngine.rebalance)"""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        out = fn(self, *args, **kwargs)
        self._journal.append((fn.__name__, args, kwargs))
        return out
    return wrapper


class Engine:
    """One simulation context on one GPU (the reference's `LAMMPS` object for this path)."""

    def __init__(self, boxlo, boxhi, periodic=(1, 1, 1), device=0):
        self.lib = load_library()
        self._h = C.c_void_p()
        lo, hi, per = _f64(boxlo), _f64(boxhi), _i32(periodic)
        rc = self.lib.le_create(C.byref(self._h), device, _pd(lo), _pd(hi), _pi(per))
        if rc != 0:
            msg = self.lib.le_last_error(self._h).decode() if self._h else "no CUDA device (there is no CPU fallback)"
            if self._h:
                self.lib.le_destroy(self._h)
                self._h = C.c_void_p()
            raise LeError(rc, msg)
        self.boxlo, self.boxhi = np.array(lo), np.array(hi)
        self.bpa, self.maxspecial, self.ntypes = 4, 16, 1
        self._journal = getattr(self, "_journal", [])

    def close(self):
        if getattr(self, "_h", None):
            self.lib.le_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise LeError(rc, self.lib.le_last_error(self._h).decode())

    # ---- settings ----
    @_journaled
    def set_types(self, masses, nbondtypes):
        m = _f64(masses)
        self.ntypes = len(m)
        self._ck(self.lib.le_set_types(self._h, len(m), _pd(m), nbondtypes))

    @_journaled
    def set_pair_lj(self, epsilon, sigma, cut, shift=True):
        """epsilon/sigma/cut: scalars (all pairs) or ntypes x ntypes matrices."""
        nt = self.ntypes
        e, s, c = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (nt, nt)))
                   for a in (epsilon, sigma, cut)]
        self._ck(self.lib.le_set_pair_lj(self._h, nt, _pd(e), _pd(s), _pd(c), 1 if shift else 0))

    @_journaled
    def set_bond(self, btype, style, params):
        style = {"fene": LE_BOND_FENE, "harmonic": LE_BOND_HARMONIC, "none": LE_BOND_NONE}.get(style, style)
        p = np.zeros(4)
        p[:len(params)] = params
        self._ck(self.lib.le_set_bond(self._h, btype, style, _pd(p)))

    @_journaled
    def set_angle_types(self, nangletypes):
        self._ck(self.lib.le_set_angle_types(self._h, nangletypes))

    @_journaled
    def set_angle(self, atype, style, params):
        """angle_coeff: style "cosine" with params (K,)"""
        style = {"cosine": 1, "none": 0}.get(style, style)
        p = np.zeros(4)
        p[:len(params)] = params
        self._ck(self.lib.le_set_angle(self._h, atype, style, _pd(p)))

    @_journaled
    def set_special(self, lj):
        a = _f64(lj)
        self._ck(self.lib.le_set_special(self._h, _pd(a)))

    @_journaled
    def set_neighbor(self, skin, every=1, delay=10, check=1):
        self._ck(self.lib.le_set_neighbor(self._h, skin, every, delay, check))

    @_journaled
    def set_neighbor_capacity(self, n):
        self._ck(self.lib.le_set_neighbor_capacity(self._h, n))

    @_journaled
    def set_newton(self, pair=1, bond=0):
        self._ck(self.lib.le_set_newton(self._h, pair, bond))

    @_journaled
    def set_capacity(self, bond_per_atom, maxspecial):
        self._ck(self.lib.le_set_capacity(self._h, bond_per_atom, maxspecial))
        self.bpa, self.maxspecial = bond_per_atom, maxspecial

    @_journaled
    def set_timestep(self, dt):
        self._ck(self.lib.le_set_timestep(self._h, dt))

    def reset_timestep(self, step):
        self._ck(self.lib.le_reset_timestep(self._h, step))

    @_journaled
    def thermo_every(self, n):
        self._ck(self.lib.le_thermo_every(self._h, n))

    # ---- fixes ----
    @_journaled
    def fix_nve(self, enable=True):
        self._ck(self.lib.le_fix_nve(self._h, 1 if enable else 0))

    @_journaled
    def fix_nve_limit(self, xmax):
        self._ck(self.lib.le_fix_nve_limit(self._h, xmax))

    @_journaled
    def fix_langevin(self, t_start, t_stop, damp, seed):
        self._ck(self.lib.le_fix_langevin(self._h, t_start, t_stop, damp, seed))

    @_journaled
    def fix_extrusion(self, nevery, neutral, left, right, p_through, btype, roadblock=-1, seed=12345):
        self._ck(self.lib.le_fix_extrusion(self._h, nevery, neutral, left, right, p_through, btype, roadblock, seed))

    @_journaled
    def fix_ex_load(self, nevery, itype, jtype, rc, btype, prob=1.0, seed=12345, iparam=(0, 0), jparam=(0, 0)):
        self._ck(self.lib.le_fix_ex_load(self._h, nevery, itype, jtype, rc, btype, prob, seed,
                                         iparam[0], iparam[1], jparam[0], jparam[1]))

    @_journaled
    def fix_ex_unload(self, nevery, btype, rc, prob=1.0, seed=12345):
        self._ck(self.lib.le_fix_ex_unload(self._h, nevery, btype, rc, prob, seed))

    @_journaled
    def fix_bond_break(self, nevery, btype, rmax, prob=1.0, seed=12345):
        """fix bond/break (src/MC/fix_bond_break.cpp): fix ex_unload's body on steps that are multiples of nevery"""
        self._ck(self.lib.le_fix_bond_break(self._h, nevery, btype, rmax, prob, seed))

    @_journaled
    def fix_bond_create(self, nevery, itype, jtype, rmin, btype, prob=1.0, seed=12345, iparam=(0, 0), jparam=(0, 0)):
        """fix bond/create (src/MC/fix_bond_create.cpp): fix ex_load's ancestor -- the closest eligible listed neighbor within
        rmin, no loop-extrusion rules, events on steps that are multiples of nevery"""
        self._ck(self.lib.le_fix_bond_create(self._h, nevery, itype, jtype, rmin, btype, prob, seed,
                                             iparam[0], iparam[1], jparam[0], jparam[1]))

    @_journaled
    def unfix(self, which):
        self._ck(self.lib.le_unfix(self._h, which))

    # ---- atoms ----
    def upload_atoms(self, types, x, v=None, image=None, tags=None):
        t, xx, vv, im, tg = _i32(types), _f64(x), _f64(v), _i32(image), _i32(tags)
        self._ck(self.lib.le_upload_atoms(self._h, len(t), _pi(tg), _pi(t), _pd(xx), _pd(vv), _pi(im)))

    def upload_bonds(self, btype, atom1, atom2):
        b, a1, a2 = _i32(btype), _i32(atom1), _i32(atom2)
        self._ck(self.lib.le_upload_bonds(self._h, len(b), _pi(b), _pi(a1), _pi(a2)))

    def upload_angles(self, atype, a1, a2, a3):
        """the data file's Angles section: type, end, centre, end (1-based tags), every angle once"""
        t, x1, x2, x3 = _i32(atype), _i32(a1), _i32(a2), _i32(a3)
        self._ck(self.lib.le_upload_angles(self._h, len(t), _pi(t), _pi(x1), _pi(x2), _pi(x3)))

    def upload_topology(self, num_bond, bond_type, bond_atom, nspecial, special):
        a = [_i32(q) for q in (num_bond, bond_type, bond_atom, nspecial, special)]
        self._ck(self.lib.le_upload_topology(self._h, *[_pi(q) for q in a]))

    def set_positions(self, x, image=None):
        xx, im = _f64(x), _i32(image)
        self._ck(self.lib.le_set_positions(self._h, _pd(xx), _pi(im)))

    def set_velocities(self, v):
        vv = _f64(v)
        self._ck(self.lib.le_set_velocities(self._h, _pd(vv)))

    # ---- run ----
    def run(self, nsteps):
        self._ck(self.lib.le_run(self._h, nsteps))

    def set_run_span(self, start, stop):
        """`run N start S stop E`: the next run() calls are segments of one run S..E (span of the Langevin ramp); stop <= start: off"""
        self._ck(self.lib.le_set_run_span(self._h, start, stop))

    def run_timed(self, nsteps):
        """run with direct launches; returns the average duration (us) of the plain step kernel"""
        us = C.c_double()
        self._ck(self.lib.le_run_timed(self._h, nsteps, C.byref(us)))
        return us.value

    def step_kernel_name(self):
        """the plain step kernel le_run launches in this configuration"""
        return self.lib.le_step_kernel_name(self._h).decode()

    def force_rebuild(self):
        self._ck(self.lib.le_force_rebuild(self._h))

    def run_le_event(self, which):
        self._ck(self.lib.le_run_le_event(self._h, which))

    def fix_rng_reset(self, which, seed, consumed=0):
        self._ck(self.lib.le_fix_rng_reset(self._h, which, seed, consumed))

    def fix_rng_consumed(self, which):
        n = C.c_int64()
        self._ck(self.lib.le_fix_rng_consumed(self._h, which, C.byref(n)))
        return n.value

    def fix_rng_set_state(self, which, state103):
        """RanMars::set_state layout (u[0..97], i97, j97, c, cd, cm)"""
        st = np.ascontiguousarray(state103, dtype=np.float64)
        assert st.shape == (103,)
        self._ck(self.lib.le_fix_rng_set_state(self._h, which, _pd(st)))

    def fix_rng_get_state(self, which):
        st = np.zeros(103)
        self._ck(self.lib.le_fix_rng_get_state(self._h, which, _pd(st)))
        return st

    def compute_forces(self):
        n = self.natoms
        f = np.zeros((n, 3))
        t = Thermo()
        self._ck(self.lib.le_compute_forces(self._h, _pd(f), C.byref(t)))
        return f, t.as_dict()

    def minimize(self, etol, ftol, maxiter, maxeval):
        """`minimize etol ftol maxiter maxeval` (min_style cg, quadratic line search); returns the "Minimization stats" """
        r = MinResult()
        self._ck(self.lib.le_minimize(self._h, float(etol), float(ftol), int(maxiter), int(maxeval), C.byref(r)))
        out = {n: getattr(r, n) for n, _ in r._fields_}
        out["stop_string"] = self.lib.le_min_stop_string(r.stop).decode()
        return out

    def compute_forces_plain(self):
        """forces from the plain instantiation of the step kernel (the one production timesteps run)"""
        f = np.zeros((self.natoms, 3))
        self._ck(self.lib.le_compute_forces_plain(self._h, _pd(f)))
        return f

    # ---- results ----
    @property
    def natoms(self):
        return self.lib.le_natoms(self._h)

    @property
    def timestep(self):
        return self.lib.le_timestep(self._h)

    def positions(self, unwrap=False):
        n = self.natoms
        x = np.zeros((n, 3))
        im = np.zeros(n, dtype=np.int32)
        self._ck(self.lib.le_download_x(self._h, _pd(x), _pi(im)))
        if unwrap:
            x = x + unpack_image(im) * (self.boxhi - self.boxlo)
        return x, im

    def owned_buffers(self, pinned=True):
        """host buffers for download_owned / upload_owned: (tag[cap], x[cap,3], image[cap], v[cap,3])"""
        cap = self.lib.le_local_capacity(self._h)
        bufs = [np.zeros(cap, np.int32), np.zeros((cap, 3)), np.zeros(cap, np.int32), np.zeros((cap, 3))]
        if pinned:
            import torch
            bufs = [torch.from_numpy(b).pin_memory().numpy() for b in bufs]
        return tuple(bufs)

    def download_owned(self, bufs):
        """fill bufs (owned_buffers()) with the atoms this GPU owns; returns their number"""
        tag, x, im, v = bufs
        n = C.c_int()
        self._ck(self.lib.le_download_owned(self._h, C.byref(n), _pi(tag), _pd(x), _pi(im), _pd(v)))
        return n.value

    def upload_owned(self, n, bufs):
        tag, x, im, v = bufs
        self._ck(self.lib.le_upload_owned(self._h, n, _pi(tag), _pd(x), _pi(im), _pd(v)))

    def velocities(self):
        v = np.zeros((self.natoms, 3))
        self._ck(self.lib.le_download_v(self._h, _pd(v)))
        return v

    def types(self):
        t = np.zeros(self.natoms, dtype=np.int32)
        self._ck(self.lib.le_download_types(self._h, _pi(t)))
        return t

    def topology(self):
        n = self.natoms
        nb = np.zeros(n, dtype=np.int32)
        bt = np.zeros((n, self.bpa), dtype=np.int32)
        ba = np.zeros((n, self.bpa), dtype=np.int32)
        ns = np.zeros((n, 3), dtype=np.int32)
        sp = np.zeros((n, self.maxspecial), dtype=np.int32)
        self._ck(self.lib.le_download_topology(self._h, _pi(nb), _pi(bt), _pi(ba), _pi(ns), _pi(sp)))
        return {"num_bond": nb, "bond_type": bt, "bond_atom": ba, "nspecial": ns, "special": sp}

    def neighlist(self, half=True):
        """CSR over tags: (offsets[N+1], entries) with entries = partner tag | which << 30."""
        n = self.natoms
        tot = C.c_int64()
        self._ck(self.lib.le_download_neighlist(self._h, 1 if half else 0, None, None, C.byref(tot)))
        off = np.zeros(n + 1, dtype=np.int64)
        ent = np.zeros(max(tot.value, 1), dtype=np.int32)
        self._ck(self.lib.le_download_neighlist(self._h, 1 if half else 0,
                                                off.ctypes.data_as(C.POINTER(C.c_int64)), _pi(ent), C.byref(tot)))
        return off, ent[:tot.value]

    def bondlist(self):
        n = C.c_int64()
        self._ck(self.lib.le_download_bondlist(self._h, None, C.byref(n)))
        rows = np.zeros((max(n.value, 1), 3), dtype=np.int32)
        self._ck(self.lib.le_download_bondlist(self._h, _pi(rows), C.byref(n)))
        return rows[:n.value]

    def thermo(self, index=None):
        if index is None:
            out = []
            for k in range(self.lib.le_thermo_count(self._h)):
                t = Thermo()
                self._ck(self.lib.le_get_thermo(self._h, k, C.byref(t)))
                out.append(t.as_dict())
            return out
        t = Thermo()
        self._ck(self.lib.le_get_thermo(self._h, index, C.byref(t)))
        return t.as_dict()

    def stats(self):
        s = Stats()
        self._ck(self.lib.le_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def observables_raw(self, s_list, rc=1.5, btype=2, nbins=64, bin_width=8):
        """device tallies over the atoms this GPU owns: (rg_sums[5], contacts[len(s_list)], loop_hist[nbins])"""
        sl = _i32(s_list)
        sums = np.zeros(5)
        cont = np.zeros(len(sl), dtype=np.int64)
        hist = np.zeros(max(nbins, 1), dtype=np.int64)
        self._ck(self.lib.le_observables(self._h, len(sl), _pi(sl), rc, btype, nbins, bin_width, _pd(sums),
                                         cont.ctypes.data_as(C.POINTER(C.c_int64)), hist.ctypes.data_as(C.POINTER(C.c_int64))))
        return sums, cont, hist[:nbins]

    def observables(self, s_list, rc=1.5, btype=2, nbins=64, bin_width=8):
        """Rg, contact probability P(s) and the loop-size histogram, computed on the GPU"""
        sums, cont, hist = self._sum_over_ranks(*self.observables_raw(s_list, rc, btype, nbins, bin_width))
        n = self.natoms
        com = sums[1:4] / sums[0]
        rg = float(np.sqrt(max(sums[4] / sums[0] - (com ** 2).sum(), 0.0)))
        ps = cont / np.maximum(n - np.asarray(s_list), 1)
        return {"rg": rg, "ps": ps, "loop_hist": hist, "nloops": int(hist.sum())}

    def _sum_over_ranks(self, *arrays):
        return arrays

    def rg(self):
        r = C.c_double()
        self._ck(self.lib.le_compute_rg(self._h, C.byref(r)))
        return r.value


def unpack_image(im):
    im = np.asarray(im)
    return np.stack([(im & 1023) - 512, ((im >> 10) & 1023) - 512, ((im >> 20) & 1023) - 512], axis=-1)


def pack_image(ixyz):
    a = np.asarray(ixyz, dtype=np.int64)
    return (((a[..., 0] + 512) & 1023) | (((a[..., 1] + 512) & 1023) << 10) | (((a[..., 2] + 512) & 1023) << 20)).astype(np.int32)


def property_local_bonds(num_bond, bond_type, bond_atom, newton_bond=0):
    """rows (batom1, batom2, btype) of `compute property/local batom1 batom2 btype` in the reference's order"""
    lib = load_host_library()
    nb = np.ascontiguousarray(num_bond, dtype=np.int32)
    bt = np.ascontiguousarray(bond_type, dtype=np.int32)
    ba = np.ascontiguousarray(bond_atom, dtype=np.int32)
    n, bpa = bt.shape
    m = lib.le_host_property_local_bonds(n, bpa, _pi(nb), _pi(bt), _pi(ba), int(newton_bond), None)
    rows = np.zeros((max(m, 1), 3), dtype=np.int32)
    lib.le_host_property_local_bonds(n, bpa, _pi(nb), _pi(bt), _pi(ba), int(newton_bond), _pi(rows))
    return rows[:m]


def velocity_create(types, masses, t_desired, seed, dist="uniform", mom=True, loop="all", x=None):
    """`velocity all create T seed dist ... mom ... loop ...` (Velocity::create, src/velocity.cpp:162-401) on the host:
    returns v[n,3] in tag order for Engine.set_velocities."""
    lib = load_host_library()
    types = np.ascontiguousarray(types, dtype=np.int32)
    masses = np.ascontiguousarray(masses, dtype=np.float64)
    n = len(types)
    v = np.zeros((n, 3))
    xx = np.ascontiguousarray(x, dtype=np.float64) if x is not None else None
    rc = lib.le_host_velocity_create(n, _pi(types), _pd(masses), _pd(xx) if xx is not None else None, float(t_desired), int(seed),
                                     {"uniform": 0, "gaussian": 1}[dist], 1 if mom else 0, {"all": 0, "local": 1, "geom": 2}[loop], _pd(v))
    if rc:
        raise LeError(rc, "velocity create: illegal arguments")
    return v


def gen_saw_chains(n, nchains, L, step=0.97, rmin=0.9, seed=12345):
    """Self-avoiding walk(s) wrapped into [0,L)^3: returns x[n,3], image[n]."""
    lib = load_host_library()
    x = np.zeros((n, 3))
    im = np.zeros(n, dtype=np.int32)
    rc = lib.le_gen_saw_chains(n, nchains, L, step, rmin, seed, _pd(x), _pi(im))
    if rc:
        raise LeError(rc, "le_gen_saw_chains: bad arguments")
    return x, im


def gen_lattice_melt(nchains, length, rho=0.8442):
    lib = load_host_library()
    n = nchains * length
    x = np.zeros((n, 3))
    im = np.zeros(n, dtype=np.int32)
    L = C.c_double()
    rc = lib.le_gen_lattice_melt(nchains, length, rho, C.byref(L), _pd(x), _pi(im))
    if rc:
        raise LeError(rc, "le_gen_lattice_melt: bad arguments")
    return x, im, L.value
