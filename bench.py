#!/usr/bin/env python3
"""bench.py -- headline benchmark: atom-steps/s of the chromatin + loop-extrusion timestep.

Workload (BASELINE.json configs[3] on one GPU): a 1,000,000-bead self-avoiding chromatin chain at number
density 0.2 with 10,000 extruders, random left/right/roadblock barriers, WCA + FENE/harmonic + Langevin/NVE,
`fix extrusion 500 / ex_load 100 / ex_unload 100` (SURVEY.md section 8d cadence for throughput runs).
One bench "step" = `--md-steps` MD timesteps (default 500 = one full USER-LE cycle) issued by ONE le_run call.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our engine (CUDA, through the C ABI)
  python bench.py --impl reference ...                           the reference's CPU code (oracle/_ref)

value      whole-job atom-steps/s with all state resident in HBM when the timed region starts
e2e        same metric through the C ABI with HOST buffers: every step uploads positions+velocities from
           pinned host memory, runs, and downloads positions
roofline   fused step kernel (k_step4, le_step4.cuh): algorithmic bytes (72.1 + 4 nbar per atom-step, nbar measured)
           / its live CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the compiled reference (oracle/_ref, threaded USER-OMP build when present, up to 16 host threads) on a
              bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "chromatin chain 1,000,000 beads rho=0.2, 10,000 extruders, random CTCF/roadblock barriers, " \
           "WCA + FENE backbone + FENE(10,4) extruder bonds + Langevin/NVE, extrusion 500 / ex_load 100 (p 0.01) / ex_unload 100 (p 0.05), dt 0.005"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


def build_system(n_beads, n_ext, seed):
    from lammps_le_b200 import systems
    # extruder bonds are FENE(10, 4.0) with the same WCA core as the pair potential: the harmonic(20, 1.3) extruder of
    # the parity fixtures lets (i, i+2) overlap and aborts with "Bad FENE bond" within ~1000 steps at 1M beads
    # (in the reference too)
    return systems.chromatin_chain(n_beads, n_ext, rho=0.2, seed=seed, barriers="random", extruder_bond=systems.EXTRUDER_FENE)


def le_fixes(e):
    e.fix_extrusion(500, 1, 2, 3, 0.5, 2, 4, 12345)
    e.fix_ex_load(100, 1, 1, 1.12, 2, 0.01, 684474, (1, 1), (1, 1))
    e.fix_ex_unload(100, 2, 0.5, 0.05, 456456)


REF_LE_LINES = ["fix loop all extrusion 500 1 2 3 0.5 2 4",
                "fix loading all ex_load 100 1 1 1.12 2 prob 0.01 684474 iparam 1 1 jparam 1 1",
                "fix unloading all ex_unload 100 2 0.5 prob 0.05 456456"]


# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_step4<0,0,1> launch at 1M beads from the committed ncu --set full
# capture (profiles/r02_ncu_full_kstep4_kbuild3.txt): an OFFLINE figure of this kernel on this workload (ncu cannot run inside
# the bench; it replays every launch with a flushed L2, so this is cold-cache traffic).  Printed only for the configuration
# it was captured on (1 GPU, 1M beads, this kernel), null otherwise.
NCU_TRAFFIC = {"kernel": "k_step4<0,0,1>", "bytes": 55.9e6, "source": "profiles/r02_ncu_full_kstep4_8blocks.txt"}
LE_HALO = 6.0   # ghost shell for USER-LE on several GPUs: longest extruder bond (FENE R0 = 4) + one backbone bond (1.5) + skin (4 cell layers here)


def prepared_engine(n_beads, n_ext, seed, device, relax_steps, dd=None):
    from lammps_le_b200 import systems
    s = build_system(n_beads, n_ext, seed)
    v = systems.maxwell_velocities(n_beads, 1.0, np.ones(n_beads), seed)
    e = systems.make_engine(s, device=device, velocities=v, dt=0.005, dd=dd)
    systems.relax(e, steps=relax_steps)
    e.fix_langevin(1.0, 1.0, 1.0, 904297)
    le_fixes(e)
    e.reset_timestep(0)
    return s, e


def reference_deck_tail(md_steps_per_seg, nseg_warm, nseg_timed):
    """fixes + run segments of the reference deck; the window is placed so that the timed segments hold a `fix extrusion`
    event (step 501) next to the ex_unload / ex_load events (502/503, 602/603, ...), like one bench step of our arm"""
    start = 490 - nseg_warm * md_steps_per_seg
    deck = ["reset_timestep %d" % max(start, 0), "fix 1 all nve", "fix 2 all langevin 1.0 1.0 1.0 904297"] + REF_LE_LINES
    deck += ["thermo_style custom step temp epair emol bonds", "thermo 1000000", "timestep 0.005"]
    deck += ["run %d" % md_steps_per_seg] * (nseg_warm + nseg_timed)
    return deck


def parse_segments(out, n, nseg_skip, nseg_timed):
    import re
    loops = [(float(a), int(b)) for a, b in re.findall(r"Loop time of ([0-9.eE+-]+) on \d+ procs for (\d+) steps", out)]
    timed = loops[-nseg_timed:] if nseg_timed else loops[nseg_skip:]
    tsum = sum(t for t, _ in timed)
    steps = sum(k for _, k in timed)
    return n * steps / tsum, tsum / max(len(timed), 1), steps


def reference_rate(s, x, image, v, topo, md_steps_per_seg, nseg_warm, nseg_timed, workdir=None, threads=1):
    """atom-steps/s of the compiled reference (oracle/_ref) started from a GIVEN state (the cpu_baseline leg of our arm:
    the same relaxed state the GPU is running).  threads > 1: the threaded build (oracle/_ref/omp, USER-OMP styles)."""
    from oracle import refio
    wd = workdir or tempfile.mkdtemp(prefix="le_bench_ref_")
    n = len(s["types"])
    s2 = dict(s)
    s2["x"], s2["image"], s2["v"] = x, image, v
    # bonds from the live topology (each bond once, lower tag first, backbone before extruders like the input)
    nb, bt, ba = topo["num_bond"], topo["bond_type"], topo["bond_atom"]
    mask = (np.arange(bt.shape[1])[None, :] < nb[:, None]) & (ba > (np.arange(n) + 1)[:, None])
    ii, mm = np.nonzero(mask)
    rows = np.stack([bt[ii, mm], ii + 1, ba[ii, mm]], axis=1)
    rows = rows[np.lexsort((rows[:, 1], rows[:, 0]))]
    rows = np.asarray(rows)
    s2["bonds"] = (rows[:, 0].astype(np.int32), rows[:, 1].astype(np.int32), rows[:, 2].astype(np.int32))
    refio.write_data_file(os.path.join(wd, "data.le"), s2)
    deck = refio.deck_header(s2, "data.le", sort=True) + reference_deck_tail(md_steps_per_seg, nseg_warm, nseg_timed)
    t0 = time.time()
    out, _ = refio.run_reference(deck, workdir=wd, harness=False, timeout=3000, threads=threads)
    wall = time.time() - t0
    rate, t_seg, steps = parse_segments(out, n, nseg_warm, nseg_timed)
    return rate, t_seg, wall, steps


def reference_rate_own_start(s, md_steps_per_seg, nseg_warm, nseg_timed, relax_steps, workdir=None, threads=1):
    """the reference arm proper: generator output -> the reference's OWN relaxation (minimize, then a push-off run with
    fix nve/limit + fix langevin) -> the timed segments, all inside one lmp_ref process; nothing of the CUDA engine is
    loaded or run"""
    from oracle import refio
    wd = workdir or tempfile.mkdtemp(prefix="le_bench_ref_")
    n = len(s["types"])
    refio.write_data_file(os.path.join(wd, "data.le"), s)
    deck = refio.deck_header(s, "data.le", sort=True)
    deck += ["thermo 1000000", "minimize 1e-4 1e-6 60 600", "reset_timestep 0", "velocity all create 1.0 12345",
             "fix r1 all nve/limit 0.05", "fix r2 all langevin 1.0 1.0 1.0 4711", "timestep 0.005", "run %d" % relax_steps,
             "unfix r1", "unfix r2"]
    deck += reference_deck_tail(md_steps_per_seg, nseg_warm, nseg_timed)
    t0 = time.time()
    out, _ = refio.run_reference(deck, workdir=wd, harness=False, timeout=3000, threads=threads)
    wall = time.time() - t0
    rate, t_seg, steps = parse_segments(out, n, 0, nseg_timed)
    return rate, t_seg, wall, steps


def reference_threads():
    """host threads for the reference: all the cores of the box when the threaded build is there (one MPI rank in any
    case: the USER-LE fixes are only defined on one rank), else 1"""
    from oracle import refio
    if not refio.have_threaded_reference():
        return 1
    if os.environ.get("LE_REF_THREADS"):
        return max(1, int(os.environ["LE_REF_THREADS"]))
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    # one thread per physical core (logical CPUs / 2), at most 16: beyond that the threaded part (pair, bond, neighbor,
    # nve) no longer shrinks the step -- fix langevin and the USER-LE fixes are serial.  Measured in the build container
    # (8 logical CPUs) at 1M beads: 2.2 M atom-steps/s with 1 thread, 3.5 M with 4, 2.9 M with 8.
    return max(1, min(cores // 2, 16))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from lammps_le_b200.engine import Engine
    from lammps_le_b200.engine_dd import init_process_group
    rank, world, local, group = init_process_group()
    torch.cuda.set_device(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_beads, n_ext, md = args.beads, args.extruders, args.md_steps
    # weak scaling (default): ONE system of world x n_beads beads (world x n_ext extruders), cut into x-slabs, one per GPU;
    # strong scaling (--scaling strong, BASELINE configs[3]): ONE system of n_beads beads over all GPUs.
    # halo positions travel as peer stores from the integrator, atoms migrate at every reneighboring
    strong = args.scaling == "strong"
    if strong:
        n_beads, n_ext = n_beads // world, n_ext // world          # per GPU; the system holds world x that
    dd = dict(rank=rank, world=world, halo=LE_HALO, group=group) if world > 1 else None
    s, e = prepared_engine(n_beads * world, n_ext * world, 12345, local, args.relax, dd)
    for _ in range(args.warmup):
        e.run(md)
    st0 = e.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    gpu_ms = le_ms = 0.0
    for _ in range(args.steps):
        e.run(md)
        sk = Engine.stats(e)
        gpu_ms += sk["last_run_gpu_ms"]
        le_ms += sk["last_run_le_ms"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    st1 = e.stats()
    tmax = torch.tensor([gpu_ms / 1e3, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    t_gpu, t_wall = float(tmax[0]), float(tmax[1])
    total_atom_steps = world * n_beads * md * args.steps
    value = total_atom_steps / t_wall          # wall clock between the two barriers (max over ranks); t_gpu = the engines' own CUDA-event time

    # ---- live duration of the dominant kernel: a short run with direct launches and CUDA events around k_step<0>
    # on the engine's own stream (steps without USER-LE event) ----
    kstep_us = e.run_timed(40)
    e.run(md - 40)            # complete the cycle so that the next segment starts at the same USER-LE phase

    # ---- end to end through the C ABI with HOST buffers: every step uploads positions + velocities of the atoms the GPU
    # owns from pinned host memory (le_upload_owned), runs, and downloads them again (le_download_owned) ----
    bufs = e.owned_buffers(pinned=True)
    nloc = e.download_owned(bufs)
    barrier()
    t0 = time.perf_counter()
    moved = 0
    for _ in range(max(1, args.steps // 2)):
        e.upload_owned(nloc, bufs)
        e.run(md)
        nloc = e.download_owned(bufs)
        moved += nloc
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e, float(moved)], dtype=torch.float64, device="cuda")
    if world > 1:
        tsum = te.clone()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        moved = float(tsum[1])
    e2e_value = world * n_beads * md * max(1, args.steps // 2) / float(te[0])
    per_step_atoms = moved / max(1, args.steps // 2)
    h2d = int(per_step_atoms * (4 + 24 + 4 + 24))
    d2h = int(per_step_atoms * (4 + 24 + 4 + 24))

    if rank != 0:
        if world > 1:
            dist.barrier()
        e.close()
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ----
    nbar_full = st1["full_entries"] / (n_beads * world)
    nbar_half = st1["half_pairs"] / (n_beads * world)
    builds = st1["neigh_builds"] - st0["neigh_builds"]
    steps_timed = md * args.steps
    kint = steps_timed / max(builds, 1)
    bytes_step = 72.1 + 4.0 * nbar_half                      # SURVEY.md 8(d) algorithmic bytes per atom-step
    bytes_amort = bytes_step + (40.0 + 4.0 * nbar_half) / kint
    peak, how = measured_peak_gbs()
    whole_step = bytes_amort * n_beads * steps_timed / t_wall / 1e9       # per GPU, everything amortised (rebuilds, USER-LE)
    achieved = bytes_step * n_beads / (kstep_us * 1e-6) / 1e9             # the step kernel alone: algorithmic bytes / its duration
    line = {
        "metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_wall / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64 pair + bond terms on exact 32-bit fixed-point differences / f32 thermostat + integration", "data": "synthetic",
        "config": {"workload": (WORKLOAD if n_beads * (world if strong else 1) == 1000000 else "same model, %d beads, %d extruders" % (n_beads, n_ext))
                   + (" -- ONE system over %d GPUs (BASELINE configs[3], strong scaling)" % world if strong else ""),
                   "beads_per_gpu": n_beads, "md_steps_per_step": md, "nbar_half": nbar_half, "nbar_full": nbar_full,
                   "steps_per_rebuild": kint, "l2_note": "working set (%.0f MB per GPU) vs 126 MB L2: inputs %s L2" % (
                       n_beads * (32 + 16 + 16 + 4 * nbar_full + 12 + 4) / 1e6, "exceed" if n_beads >= 1000000 else "fit in"),
                   "parallelism": ("one %d-bead system in %d x-slabs (one per GPU): halo stores over NVLink peer memory fused into "
                                   "the integrator, flag hand-shakes, migration at every rebuild, replicated USER-LE logic" % (n_beads * world, world))
                   if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(st1["kernel_launches"] - st0["kernel_launches"]),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC["bytes"] if (n_beads == 1000000 and world == 1 and Engine.step_kernel_name(e) == NCU_TRAFFIC["kernel"]) else None,
                     "peak_source": how,
                     "kernel": Engine.step_kernel_name(e), "kernel_us": kstep_us, "bytes_per_atom_step": bytes_step,
                     "traffic_source": "offline: dram__bytes_read.sum + dram__bytes_write.sum of one %s launch under ncu --set full (flushed L2), %s; null when the run is not that configuration" % (NCU_TRAFFIC["kernel"], NCU_TRAFFIC["source"]),
                     "whole_step": {"achieved": whole_step, "frac": whole_step / peak, "bytes_per_atom_step": bytes_amort,
                                    "note": "72.1 + 4 nbar + (40 + 4 nbar)/K bytes per atom-step over the whole timed loop (rebuilds and USER-LE included)"}},
        "wall_s": t_wall, "gpu_event_s": t_gpu, "user_le_ms_per_md_step": le_ms / steps_timed,
        "e2e_note": "one bench step = %d MD timesteps in ONE le_run call; the e2e copies (le_upload_owned / le_download_owned of all owned atoms, pinned host memory) happen once per bench step, i.e. once per %d MD steps" % (md, md),
        "le_events": {"shifts": st1["extrusion_shifts"] - st0["extrusion_shifts"], "loads": st1["loads"] - st0["loads"],
                      "unloads": st1["unloads"] - st0["unloads"]},
    }
    # ---- CPU baseline: the compiled reference on a bounded sample of the same state ----
    if world == 1 and not args.no_cpu:
        try:
            from oracle import refio
            if refio.have_reference():
                topo = e.topology()
                seg = max(4, int(2.0e6 / n_beads * 8))          # ~16 MD steps per segment at 1M beads
                xo, im = e.positions()
                nthr = reference_threads()
                rate, _, wall_ref, nst = reference_rate(s, xo, im, e.velocities(), topo, seg, 1, 5, threads=nthr)
                line["cpu_baseline"] = {"value": rate, "unit": "atom-steps/s", "cores": nthr, "kind": "reference",
                                        "sample": "%d MD steps of the same relaxed state in the compiled reference (oracle/_ref, 1 MPI rank: USER-LE "
                                                  "is only defined on one rank; %s), window holds one fix extrusion event and the ex_unload / ex_load events next to it, %.0f s wall "
                                                  "incl. setup" % (nst, "%d OpenMP threads, USER-OMP pair/bond/neighbor/nve styles, fix langevin and "
                                                  "the USER-LE fixes serial" % nthr if nthr > 1 else "serial build", wall_ref)}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "atom-steps/s", "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref missing on this box"}
        except Exception as ex:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": "atom-steps/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % str(ex)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    e.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refio
    n_beads, n_ext = args.beads, args.extruders
    if not refio.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
        return
    # the generator is host-only (libleb200_host.so, plain g++): the CUDA engine is neither loaded nor run in this arm; the start
    # state is relaxed by the reference's own minimize + push-off run inside the same lmp_ref process that is then timed
    s = build_system(n_beads, n_ext, 12345)
    seg = max(4, int(2.0e6 / n_beads * 8))
    nthr = reference_threads()
    rate, t_seg, wall, nst = reference_rate_own_start(s, seg, args.warmup, args.steps, max(100, args.relax // 5), threads=nthr)
    line = {"impl": "reference", "metric": "atom-steps/s", "value": rate, "unit": "atom-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_seg, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if n_beads == 1000000 else "same model, %d beads, %d extruders" % (n_beads, n_ext),
                       "md_steps_per_step": seg, "parallelism": "1 MPI rank (serial stubs), %d OpenMP thread(s), g++ -O3" % nthr},
            "cpu_baseline": {"value": rate, "unit": "atom-steps/s", "cores": nthr, "kind": "reference",
                             "sample": "%d MD steps in %d `run` segments of the compiled reference (oracle/_ref/omp: -O3, USER-OMP pair/bond/"
                                       "neighbor/nve styles; fix langevin and the USER-LE fixes are serial), start state relaxed by the reference "
                                       "itself (minimize + %d steps of fix nve/limit); the timed window holds one fix extrusion event and the "
                                       "ex_unload / ex_load events next to it; Loop time of each segment as LAMMPS prints it; %.0f s wall in all"
                                       % (nst, args.steps, max(100, args.relax // 5), wall)},
            "e2e": {"value": rate, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--beads", type=int, default=1000000)
    ap.add_argument("--extruders", type=int, default=10000)
    ap.add_argument("--md-steps", type=int, default=500)
    ap.add_argument("--relax", type=int, default=1500)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --beads per GPU (one system of N x beads); strong: --beads in all (BASELINE configs[3]: one 1M-bead chain over N GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
