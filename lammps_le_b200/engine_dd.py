"""Several GPUs of one box: one process and one `le_ctx` per GPU, x-slab domain decomposition.

`DDEngine` is `Engine` plus the hand-shakes the C ABI leaves to the caller (include/le_b200.h, "several GPUs"):
  * the CUDA IPC handles of the peer arenas are all-gathered with torch.distributed and handed to le_dd_connect;
  * uploads take the whole system on every rank (as every LAMMPS rank reads the data file);
  * downloads fill the entries of the atoms a GPU owns: the wrapper zero-fills and sums over ranks;
  * thermo tallies are summed over ranks before they are normalised (ComputeTemp / ComputePressure do an
    MPI_Allreduce there, src/compute_temp.cpp:85, src/compute_pressure.cpp:180).
The data path itself (halo update, migration, reneighbor decision) never goes through this file: it is peer-memory
stores and flags issued by the kernels (csrc/le_md.cuh).  torch.distributed is plumbing only.
"""
import ctypes as C

import numpy as np

from .engine import Engine, Thermo, _pd, _pi  # noqa: F401


def slab_of_cells(ncx, world):
    """[X0, X1) of every rank for EQUAL-WIDTH slabs: the split le_engine.cu:setup_cells uses when no cuts by atom count exist
    (le_dd_balance(ctx, 0) / DDEngine(balance=False)); the default places the cuts by the cumulative atom count instead"""
    return [(r * ncx // world, (r + 1) * ncx // world) for r in range(world)]


def finalize_thermo(sums, natoms, volume, step=0, nbonds=0):
    """thermo record from summed tallies (le_engine.cu:thermo_from_slot; ComputeTemp, ComputePressure)"""
    n = float(natoms)
    dof = 3.0 * n - 3.0
    return {"step": step, "ke": 0.5 * sums[0], "temp": sums[0] / dof if dof > 0 else 0.0, "epair": sums[1] / n,
            "emol": (sums[2] + sums[10]) / n, "eangle": sums[10] / n, "etotal": (0.5 * sums[0] + sums[1] + sums[2] + sums[10]) / n, "virial": list(sums[3:9]),
            "press": (sums[0] + sums[3] + sums[4] + sums[5]) / (3.0 * volume), "fene_warnings": int(round(sums[9])),
            "nbonds": nbonds}


def merge_csr(parts):
    """merge per-rank CSR neighbor lists whose rows are empty for atoms a rank does not own"""
    n = len(parts[0][0]) - 1
    counts = sum(np.diff(off) for off, _ in parts)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(counts)
    ent = np.zeros(int(off[-1]), dtype=np.int32)
    fill = off[:-1].copy()
    for o, e in parts:
        c = np.diff(o)
        rows = np.nonzero(c)[0]
        for t in rows:
            ent[fill[t]:fill[t] + c[t]] = e[o[t]:o[t + 1]]
            fill[t] += c[t]
    return off, ent


class DDEngine(Engine):
    def __init__(self, boxlo, boxhi, periodic=(1, 1, 1), device=0, rank=0, world=1, halo=0.0, group=None, balance=True):
        super().__init__(boxlo, boxhi, periodic, device)
        self.rank, self.world, self.group = rank, world, group
        self._ctor = (boxlo, boxhi, periodic, device, float(halo))
        self._ck(self.lib.le_dd_init(self._h, rank, world, float(halo)))
        if not balance:
            self._ck(self.lib.le_dd_balance(self._h, 0))     # equal-width slabs, as the reference without a balance command

    # ---- load balance ----
    def owned_counts(self):
        """atoms every GPU owns right now (length `world`)"""
        n = np.zeros(self.world, dtype=np.int64)
        n[self.rank] = self.download_owned(self.owned_buffers(pinned=False))
        return self._allreduce(n)

    def imbalance(self):
        """max / mean of the owned counts: the `imbalance factor` of the reference's balance command (src/balance.cpp:245-300)"""
        c = self.owned_counts().astype(np.float64)
        return float(c.max() / c.mean())

    def rebalance(self, thresh=1.0):
        """Dynamic load balance in the sense of `fix balance N thresh shift x ...` (src/fix_balance.cpp:191-270, src/balance.cpp):
        when the imbalance factor exceeds `thresh`, the slab cuts are placed anew where the cumulative atom count of the CURRENT
        configuration crosses r N / P.  The cuts are part of the device layout (slab widths size the peer arenas and the ghost
        regions), so the context is rebuilt from the recorded settings / fixes and the state is carried over: positions, image
        flags, velocities, types, the per-atom bond and special tables (extruder bonds included), the Marsaglia state of each
        USER-LE fix and the timestep.  Collective; returns (imbalance before, imbalance after) -- equal when nothing was done."""
        before = self.imbalance()
        if self.world == 1 or before <= thresh:
            return before, before
        x, im = self.positions()
        v, ty, topo, step = self.velocities(), self.types(), self.topology(), int(self.timestep)
        rng = {}
        for which in (1, 2, 3):
            try:
                rng[which] = self.fix_rng_get_state(which)
            except Exception:
                pass
        journal = list(self._journal)
        self.barrier()
        self.close()
        boxlo, boxhi, periodic, device, halo = self._ctor
        self._journal = []
        DDEngine.__init__(self, boxlo, boxhi, periodic, device, self.rank, self.world, halo, self.group)
        for name, args, kwargs in journal:
            getattr(self, name)(*args, **kwargs)
        self.upload_atoms(ty, x, v, im)
        self.upload_topology(topo["num_bond"], topo["bond_type"], topo["bond_atom"], topo["nspecial"], topo["special"])
        for which, st in rng.items():
            self.fix_rng_set_state(which, st)
        self.reset_timestep(step)
        return before, self.imbalance()

    # ---- plumbing ----
    def _allreduce(self, a):
        if self.world == 1:
            return a
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(a)
        dist.all_reduce(t, group=self.group)
        return a

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def upload_atoms(self, types, x, v=None, image=None, tags=None):
        super().upload_atoms(types, x, v, image, tags)
        if self.world > 1:
            import torch
            import torch.distributed as dist
            h = (C.c_ubyte * 64)()
            self._ck(self.lib.le_dd_get_handle(self._h, C.cast(h, C.c_void_p)))
            mine = torch.tensor(list(h), dtype=torch.uint8)
            allh = [torch.zeros(64, dtype=torch.uint8) for _ in range(self.world)]
            dist.all_gather(allh, mine, group=self.group)
            buf = (C.c_ubyte * (64 * self.world))(*[int(b) for t in allh for b in t.tolist()])
            self._ck(self.lib.le_dd_connect(self._h, C.cast(buf, C.c_void_p)))
            self.barrier()

    # ---- collective calls: every rank enters together (the kernels spin on peer flags) ----
    def run(self, nsteps):
        self.barrier()
        super().run(nsteps)

    def run_timed(self, nsteps):
        self.barrier()
        return super().run_timed(nsteps)

    def force_rebuild(self):
        self.barrier()
        super().force_rebuild()

    def run_le_event(self, which):
        self.barrier()
        super().run_le_event(which)

    def compute_forces(self):
        self.barrier()
        n = self.natoms
        f = np.zeros((n, 3))
        t = Thermo()
        self._ck(self.lib.le_compute_forces(self._h, _pd(f), C.byref(t)))
        if self.world == 1:
            return f, t.as_dict()
        s = np.zeros(16)
        self._ck(self.lib.le_get_force_sums(self._h, _pd(s)))
        self._allreduce(f)
        self._allreduce(s)
        return f, finalize_thermo(s, n, float(np.prod(self.boxhi - self.boxlo)), t.step, t.nbonds)

    # ---- results: owned entries summed over ranks ----
    def positions(self, unwrap=False):
        x, im = super().positions(False)
        self._allreduce(x)
        self._allreduce(im)
        if unwrap:
            from .engine import unpack_image
            x = x + unpack_image(im) * (self.boxhi - self.boxlo)
        return x, im

    def velocities(self):
        return self._allreduce(super().velocities())

    def types(self):
        return self._allreduce(super().types())

    def neighlist(self, half=True):
        off, ent = super().neighlist(half)
        if self.world == 1:
            return off, ent
        import torch
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, (off, ent), group=self.group)
        return merge_csr(parts)

    def bondlist(self):
        rows = super().bondlist()
        if self.world == 1:
            return rows
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, rows, group=self.group)
        allr = np.concatenate(parts, axis=0)
        # the reference's order: ascending owner tag, then slot order (kept inside a rank: stable sort on the owner)
        return allr[np.argsort(allr[:, 0], kind="stable")]

    def thermo(self, index=None):
        if self.world == 1:
            return super().thermo(index)
        n = self.lib.le_thermo_count(self._h)
        idx = range(n) if index is None else [index if index >= 0 else n + index]
        out = []
        vol = float(np.prod(self.boxhi - self.boxlo))
        for k in idx:
            t = Thermo()
            self._ck(self.lib.le_get_thermo(self._h, k, C.byref(t)))
            s = np.zeros(16)
            self._ck(self.lib.le_get_thermo_sums(self._h, k, _pd(s)))
            self._allreduce(s)
            out.append(finalize_thermo(s, self.natoms, vol, t.step, t.nbonds))
        return out if index is None else out[0]

    def _sum_over_ranks(self, *arrays):
        return tuple(self._allreduce(np.ascontiguousarray(a)) for a in arrays)

    def stats(self):
        s = super().stats()
        if self.world > 1:
            a = np.array([s["half_pairs"], s["full_entries"]], dtype=np.int64)
            self._allreduce(a)
            s["half_pairs"], s["full_entries"] = int(a[0]), int(a[1])
        return s


def init_process_group():
    """(rank, world, local_rank, gloo group) from the torchrun environment; (0, 1, 0, None) when run plainly.
    LE_DD_SHARE_GPU=1 puts every rank on device 0 (development on a one-GPU box: the slabs then time-slice one GPU and
    talk through CUDA IPC exactly as they would over NVLink; NCCL refuses two ranks per device, so gloo only)."""
    import os
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return 0, 1, local, None
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    share = os.environ.get("LE_DD_SHARE_GPU", "0") == "1"
    if share:
        local = 0
    if not dist.is_initialized():
        if torch.cuda.is_available() and not share:
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    group = dist.new_group(backend="gloo")        # small host-side exchanges (IPC handles, result sums)
    return rank, world, local, group
