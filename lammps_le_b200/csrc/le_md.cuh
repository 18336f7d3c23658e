// le_md.cuh -- peer flags, the reneighbor decision, migration / ghost exchange of a rebuild, host <-> device exchange,
// observables.  (The fused step kernel is in le_step3.cuh, the cell sort in le_sort.cuh, the list build in le_build3.cuh.)
//
// Launch structure of one timestep (le_engine.cu): k_step3 -> k_decide -> [conditional graph node:
// k_cell_count .. k_build3].  The step number and the position-buffer parity live in the device-side
// control block, so a captured CUDA graph of several steps is replayed unchanged for the whole run
// and the rebuild kernels are only launched on the steps that need them.
#pragma once
#include "le_common.cuh"

struct StepArgs {
  int do_final;       // second half of velocity Verlet for the step whose forces are computed here
  int do_initial;     // first half of the next step (v += dtf f/m; x += dt v)
  int slot;           // thermo slot to tally energy / virial / kinetic energy into (EV kernels)
  int write_force;    // store the conservative force of every atom in fout (tag order)
  int langevin;       // add drag + noise
  int rdp1;           // 1 + position buffer holding the current coordinates when the host knows it, 0 = read Ctrl::cur
  int angles;         // add the angle forces k_angle left in Dev::fang
};

// ------------------------------------------------------------------------------------------------
// peer flags (multi-GPU): system-scope release stores / acquire loads on words in a peer's arena
// ------------------------------------------------------------------------------------------------
// The poster issues ONE __threadfence_system() (all its earlier stores, and -- by stream order -- those of the
// kernels before it, are then visible system-wide) followed by relaxed flag stores to each peer; the waiter polls
// with relaxed loads and fences once after the last flag arrived.  (A release store / acquire load per flag costs
// a full system fence each.)
__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
#define LE_PEER_TIMEOUT_CYCLES 40000000000LL   // ~20 s: a peer that is this late is gone
// spin until (word >> shift) >= want; gives up (and makes every later wait return at once) on timeout / error
__device__ __noinline__ unsigned long long le_wait_flag(Ctrl *c, const unsigned long long *p, unsigned long long want, int shift) {
  const long long t0 = clock64();
  for (;;) {
    const unsigned long long v = ld_sys(p);
    if ((v >> shift) >= want) return v;
    if (*(volatile int *)&c->err) return 0;
    if (clock64() - t0 > LE_PEER_TIMEOUT_CYCLES) { le_raise(c, LE_DERR_PEER_TIMEOUT, (int)want, (int)(v >> shift), shift); return 0; }
    __nanosleep(64);
  }
}
__device__ __forceinline__ int left_rank(const Dev &d) { return (d.rank + d.nranks - 1) % d.nranks; }
__device__ __forceinline__ int right_rank(const Dev &d) { return (d.rank + 1) % d.nranks; }

// close a force-evaluation epoch (one thread, in the kernel that follows k_step in stream order): tell every peer
// that this GPU's k_step -- and with it the halo stores into the peer's ghost slots -- is complete, together with
// this GPU's "an atom moved half the skin" bit; then wait for the same word from every peer and fold the bits, so
// that the reneighbor decision is identical everywhere (the MPI_Allreduce of Neighbor::decide,
// src/neighbor.cpp:1944).  Words of consecutive epochs alternate between two slots: a peer can run at most one
// epoch ahead.
__device__ __forceinline__ void close_epoch(const Dev &d) {
  Ctrl *c = d.ctrl;
  const unsigned long long e = (unsigned long long)c->epoch;
  if (d.nranks > 1) {
    __threadfence_system();
    int moved = c->moved;
    const unsigned long long word = (e << 1) | (unsigned long long)(moved != 0);
    const int slot = FLAG_STEP + (int)(e & 1) * LE_MAXRANKS;
    for (int p = 0; p < d.nranks; p++)
      if (p != d.rank) st_sys(&d.peer[p].flags[slot + d.rank], word);
    for (int p = 0; p < d.nranks; p++) {
      if (p == d.rank) continue;
      const unsigned long long v = le_wait_flag(c, &d.flags[slot + p], e, 1);
      moved |= (int)(v & 1);
    }
    __threadfence_system();
    c->moved = moved;
  }
  c->epoch = (long long)e + 1;
}

// ------------------------------------------------------------------------------------------------
// Neighbor::decide (src/neighbor.cpp:1933-1948): one thread.  With advance != 0 it first closes the
// timestep (ntimestep++, swap the position buffers, wait for the peers' halos).  When it decides to
// rebuild it also does the bookkeeping of Neighbor::build (ago = 0, ncalls++) because the rebuild
// kernels that follow are a conditional graph node switched by cudaGraphSetConditional.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void decide_body(const Dev &d, cudaGraphConditionalHandle handle, int advance, int use_handle) {
  Ctrl *c = d.ctrl;
  if (advance) { c->step++; c->cur ^= 1; close_epoch(d); }
  // moved / forced / rebuild_now / ago are the first 16 bytes of the control block: one load, one store
  int4 w;
  asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(c) : "memory");
  int moved = w.x, forced = w.y, ago = w.w;
  int r = 0;
  if (forced) r = 1;
  else {
    ago++;
    if (ago >= c_P.delay && ago % c_P.every == 0) {
      if (!c_P.check) r = 1;
      else if (moved) {
        r = 1;
        const int mx = c_P.every > c_P.delay ? c_P.every : c_P.delay;
        if (ago == mx) c->ndanger++;
      }
    }
  }
  if (r) { ago = 0; c->nbuilds++; }
  // `moved` is the state of THIS step (Neighbor::check_distance looks at the displacements only on the steps where
  // delay / every allow a rebuild, neighbor.cpp:1943-1947): the step kernel re-tests every atom whose bound has reached
  // skin/2 on every step, so the flag is cleared here and an atom that moved out and back in does not trigger later
  *reinterpret_cast<int4 *>(c) = make_int4(0, 0, r, ago);
  if (use_handle) cudaGraphSetConditional(handle, r ? 1u : 0u);
}

__global__ void k_decide(Dev d, cudaGraphConditionalHandle handle, int advance, int use_handle) { decide_body(d, handle, advance, use_handle); }

// close a timestep without deciding (the USER-LE fixes of the new step run before Neighbor::decide)
__global__ void k_advance(Dev d) { d.ctrl->step++; d.ctrl->cur ^= 1; close_epoch(d); }

// bookkeeping of an unconditional rebuild (Verlet::setup, le_force_rebuild)
__global__ void k_after_build(Dev d) {
  d.ctrl->moved = 0;
  d.ctrl->forced = 0;
  d.ctrl->ago = 0;
  d.ctrl->nbuilds++;
}

// ------------------------------------------------------------------------------------------------
// rebuild, part 1: migration + cell sort of the owned atoms.
//   Cells are at least one neighbor cutoff wide; the fixed-point coordinate gives the cell by one
//   multiply-high.  Within a cell atoms are ordered by tag so that the local order -- and with it every
//   floating-point sum downstream -- is reproducible from run to run.  An atom whose cell has left
//   this GPU's slab is written straight into the inbox of the neighbor GPU (CommBrick::exchange,
//   src/comm_brick.cpp:601-722); bond and special rows are replicated by tag, so the extruder bonds
//   of a migrating bead arrive with it.
// ------------------------------------------------------------------------------------------------
struct RbScratch { int out_count[2]; int in_count[2]; };

__global__ void k_cell_count(Dev d, RbScratch *rb) {
  const int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  const int lo = d.own0, hi = d.own0 + d.ctrl->nown;
  if (d.nranks > 1) {
    // forget the ghosts of the previous list (their tags may be anywhere after this rebuild).  ghost_tag is this
    // GPU's private record of its ghosts: the peers may already be overwriting pos_hold's ghost slots
    const int nl = d.ctrl->nghl, nr = d.ctrl->nghr;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < nl + nr; g += gridDim.x * blockDim.x) d.map[d.ghost_tag[g] - 1] = -1;
  }
  for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
    const int4 p = pos[i];
    const int cx = __umulhi((unsigned)p.x, (unsigned)d.ncell[0]);
    const int cy = __umulhi((unsigned)p.y, (unsigned)d.ncell[1]);
    const int cz = __umulhi((unsigned)p.z, (unsigned)d.ncell[2]);
    const int lx = local_layer(d, cx);
    if (lx >= d.halo && lx < d.nlx - d.halo) {
      const int c = cell_slot(d, lx, cy, cz);
      d.cellid[i] = c;
      d.slot[i] = atomicAdd(&d.cell_count[c], 1);
    } else {
      d.cellid[i] = -1;
      d.map[(p.w >> 3) - 1] = -1;
      if (lx < 0) { le_raise(d.ctrl, LE_DERR_LOCAL_OVERFLOW, p.w >> 3, cx, 1); continue; }
      const int side = lx < d.halo ? 0 : 1;                 // 0: to the left neighbor
      const int k = atomicAdd(&rb->out_count[side], 1);
      if (k >= d.inbox_cap) { le_raise(d.ctrl, LE_DERR_LOCAL_OVERFLOW, k, d.inbox_cap, 2); continue; }
      const PeerView &pv = d.peer[side == 0 ? left_rank(d) : right_rank(d)];
      const size_t o = (size_t)(side == 0 ? 1 : 0) * d.inbox_cap + k;   // it arrives "from the right" at the left neighbor
      pv.in_pos[o] = p; pv.in_vel[o] = d.vel[i]; pv.in_img[o] = d.img[i];
    }
  }
}

// tell the neighbors how many migrants were written into their inboxes, wait for ours
__global__ void k_rb_post_inbox(Dev d, RbScratch *rb) {
  Ctrl *c = d.ctrl;
  __threadfence_system();
  const unsigned long long e = (unsigned long long)(++c->rebuild_epoch);
  const int par = (int)(e & 1) * 2;
  st_sys(&d.peer[left_rank(d)].flags[FLAG_INBOX + par + 1], (e << 24) | (unsigned)rb->out_count[0]);
  st_sys(&d.peer[right_rank(d)].flags[FLAG_INBOX + par + 0], (e << 24) | (unsigned)rb->out_count[1]);
  rb->out_count[0] = rb->out_count[1] = 0;        // for the next rebuild
  for (int side = 0; side < 2; side++) {
    const unsigned long long v = le_wait_flag(c, &d.flags[FLAG_INBOX + par + side], e, 24);
    rb->in_count[side] = (int)(v & 0xffffffu);
  }
  __threadfence_system();
  c->nown_unsorted = c->nown + rb->in_count[0] + rb->in_count[1];
  if (d.own0 + c->nown_unsorted > d.gr0) le_raise(c, LE_DERR_LOCAL_OVERFLOW, c->nown_unsorted, d.gr0 - d.own0, 3);
}

// append the arrived atoms behind the owned ones and count them into their cells
__global__ void k_inbox(Dev d, RbScratch *rb) {
  const int n0 = rb->in_count[0], n1 = rb->in_count[1];
  const int cur = d.ctrl->cur;
  const int base = d.own0 + d.ctrl->nown;
  if (base + n0 + n1 > d.gr0) return;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n0 + n1; k += gridDim.x * blockDim.x) {
    const size_t o = k < n0 ? (size_t)k : (size_t)d.inbox_cap + (k - n0);
    const int4 p = d.in_pos[o];
    const int i = base + k;
    d.pos[cur][i] = p; d.vel[i] = d.in_vel[o]; d.img[i] = d.in_img[o];
    const int cx = __umulhi((unsigned)p.x, (unsigned)d.ncell[0]);
    const int cy = __umulhi((unsigned)p.y, (unsigned)d.ncell[1]);
    const int cz = __umulhi((unsigned)p.z, (unsigned)d.ncell[2]);
    const int lx = local_layer(d, cx);
    if (lx < d.halo || lx >= d.nlx - d.halo) { le_raise(d.ctrl, LE_DERR_LOCAL_OVERFLOW, p.w >> 3, cx, 4); d.cellid[i] = -1; continue; }
    const int c = cell_slot(d, lx, cy, cz);
    d.cellid[i] = c;
    d.slot[i] = atomicAdd(&d.cell_count[c], 1);
  }
}

// ------------------------------------------------------------------------------------------------
// rebuild, part 2 (multi-GPU): ghost creation (CommBrick::borders, src/comm_brick.cpp:727-876).
//   Because x is the slowest index of the local order, the atoms of the slab's first / last `halo`
//   layers are one contiguous slice each; it is stored verbatim into the neighbor's ghost slots
//   together with the matching cell_start entries, so the neighbor needs neither a sort nor a scan
//   for its ghosts and the per-step halo update of k_step is slot-for-slot.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int peer_cell_slot(const Dev &d, int nlx_peer, int lx, int cy, int cz) {
  return (lx * d.ncell[1] + cy) * d.ncell[2] + cz + (lx >= d.halo) + (lx >= nlx_peer - d.halo);
}

__global__ void k_push_ghosts(Dev d) {
  Ctrl *c = d.ctrl;
  const int H = d.halo, ncy = d.ncell[1], ncz = d.ncell[2];
  const int layer = ncy * ncz;
  const int own_lo = d.own0, own_hi = d.own0 + c->nown;
  const int sl_end = d.cell_start[cell_slot(d, 2 * H - 1, ncy - 1, ncz - 1) + 1];   // end of my first H layers
  const int sr_beg = d.cell_start[cell_slot(d, d.nlx - 2 * H, 0, 0)];               // start of my last H layers
  const PeerView &L = d.peer[left_rank(d)], &R = d.peer[right_rank(d)];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (tid == 0) {
    c->send_l_end = sl_end; c->send_r_beg = sr_beg;
    if (sl_end - own_lo > d.own0 || own_hi - sr_beg > d.own0) le_raise(c, LE_DERR_LOCAL_OVERFLOW, sl_end - own_lo, own_hi - sr_beg, 5);
  }
  if (sl_end - own_lo > d.own0 || own_hi - sr_beg > d.own0) return;
  for (int k = own_lo + tid; k < sl_end; k += nth) {          // -> right ghosts of the left neighbor
    L.pos_hold[d.gr0 + (k - own_lo)] = d.pos_hold[k];   // the receiver copies it into its live buffer (k_ghost_map)
  }
  for (int k = sr_beg + tid; k < own_hi; k += nth) {          // -> left ghosts of the right neighbor
    R.pos_hold[k - sr_beg] = d.pos_hold[k];
  }
  for (int q = tid; q <= H * layer; q += nth) {               // cell_start of those layers (+ the closing sentinel)
    const int lx = q / layer, rem = q - lx * layer;
    const int cy = rem / ncz, cz = rem - cy * ncz;
    if (q < H * layer) {
      L.cell_start[peer_cell_slot(d, d.nlx_left, d.nlx_left - H + lx, cy, cz)] = d.gr0 + (d.cell_start[cell_slot(d, H + lx, cy, cz)] - own_lo);
      R.cell_start[peer_cell_slot(d, d.nlx_right, lx, cy, cz)] = d.cell_start[cell_slot(d, d.nlx - 2 * H + lx, cy, cz)] - sr_beg;
    } else {
      L.cell_start[d.nlx_left * layer + 2] = d.gr0 + (sl_end - own_lo);
      R.cell_start[H * layer] = own_hi - sr_beg;
    }
  }
}

__global__ void k_rb_post_ghosts(Dev d) {
  Ctrl *c = d.ctrl;
  __threadfence_system();
  const unsigned long long e = (unsigned long long)c->rebuild_epoch;
  const int par = (int)(e & 1) * 2;
  // my first layers are the left neighbor's RIGHT ghosts (side 1 there); my last layers the right neighbor's LEFT ghosts
  st_sys(&d.peer[left_rank(d)].flags[FLAG_GHOST + par + 1], (e << 24) | (unsigned)(c->send_l_end - d.own0));
  st_sys(&d.peer[right_rank(d)].flags[FLAG_GHOST + par + 0], (e << 24) | (unsigned)(d.own0 + c->nown - c->send_r_beg));
  const unsigned long long v0 = le_wait_flag(c, &d.flags[FLAG_GHOST + par + 0], e, 24);
  const unsigned long long v1 = le_wait_flag(c, &d.flags[FLAG_GHOST + par + 1], e, 24);
  __threadfence_system();
  c->nghl = (int)(v0 & 0xffffffu);
  c->nghr = (int)(v1 & 0xffffffu);
}

__global__ void k_ghost_map(Dev d) {
  const int nl = d.ctrl->nghl, nr = d.ctrl->nghr;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < nl + nr; g += gridDim.x * blockDim.x) {
    const int k = g < nl ? g : d.gr0 + (g - nl);
    const int4 p = d.pos_hold[k];
    const int tag = p.w >> 3;
    d.pos[d.ctrl->cur][k] = p;
    d.ghost_tag[g] = tag;
    d.map[tag - 1] = k;
  }
}

// ------------------------------------------------------------------------------------------------
// host <-> device exchange of the atoms this GPU owns (le_download_owned / le_upload_owned): conversion between
// the reference's doubles and the fixed-point / fp32 device state happens here, so the host only moves flat
// buffers.  The quantisation is the one le_upload_atoms does on the host (nearest grid point; a coordinate
// outside the box is wrapped and the wrap goes into the image flags, Domain::remap src/domain.cpp:1050-1110).
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_owned(Dev d, int *tag, double *x, int *image, double *v) {
  const int n = d.ctrl->nown;
  const int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int k = d.own0 + j;
    const int4 p = pos[k];
    if (tag) tag[j] = p.w >> 3;
    if (x) {
      x[3 * j] = le_deq((unsigned)p.x, 0); x[3 * j + 1] = le_deq((unsigned)p.y, 1); x[3 * j + 2] = le_deq((unsigned)p.z, 2);
    }
    if (image) image[j] = d.img[k];
    if (v) { const float4 vv = d.vel[k]; v[3 * j] = vv.x; v[3 * j + 1] = vv.y; v[3 * j + 2] = vv.z; }
  }
}

__global__ void k_unpack_owned(Dev d, int n, const int *tag, const double *x, const int *image, const double *v) {
  int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  const int own_end = d.own0 + d.ctrl->nown;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int t = tag[j];
    const int k = (t >= 1 && t <= d.N) ? d.map[t - 1] : -1;
    if (k < d.own0 || k >= own_end) { le_raise(d.ctrl, LE_DERR_MISSING_ATOM, t, k, 4); continue; }
    if (x) {
      int4 p = pos[k];
      int w[3];
      unsigned u[3];
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const double f = __ddiv_rn(__dsub_rn(x[3 * j + q], c_P.lo[q]), c_P.L[q]);
        const double fl = floor(f);
        double uu = rint(__dmul_rn(__dsub_rn(f, fl), 4294967296.0));
        int ww = (int)fl;
        if (uu >= 4294967296.0) { uu -= 4294967296.0; ww += 1; }
        u[q] = (unsigned)uu; w[q] = ww;
      }
      int ix, iy, iz;
      if (image) {
        const int im = image[j];
        ix = (im & 1023) - 512 + w[0]; iy = ((im >> 10) & 1023) - 512 + w[1]; iz = ((im >> 20) & 1023) - 512 + w[2];
      } else {
        // no image flags given: keep the unwrapped trajectory continuous (nearest-image move)
        const int im = d.img[k];
        ix = (im & 1023) - 512 + le_image_shift((unsigned)p.x, u[0]);
        iy = ((im >> 10) & 1023) - 512 + le_image_shift((unsigned)p.y, u[1]);
        iz = ((im >> 20) & 1023) - 512 + le_image_shift((unsigned)p.z, u[2]);
      }
      d.img[k] = ((ix + 512) & 1023) | (((iy + 512) & 1023) << 10) | (((iz + 512) & 1023) << 20);
      p.x = (int)u[0]; p.y = (int)u[1]; p.z = (int)u[2];
      pos[k] = p;
    }
    if (v) {
      float4 vv = d.vel[k];
      vv.x = (float)v[3 * j]; vv.y = (float)v[3 * j + 1]; vv.z = (float)v[3 * j + 2];
      d.vel[k] = vv;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// polymer observables on the device: radius of gyration from unwrapped coordinates
// (ComputeGyration, src/compute_gyration.cpp:60-100), contact counts of bead pairs (t, t+s) for a list of chain
// separations s (-> contact probability P(s)), and the histogram of extruder loop sizes |b - a| over the bonds of
// one type (the loops; compute property/local batom1 batom2 btype, src/compute_property_local.cpp:104-117).
// One pass over the owned atoms; every GPU tallies its own atoms, the caller sums over GPUs.
//   out_d[0..4] = sum m, sum m x, sum m y, sum m z, sum m |x|^2      out_i[0..ns) contacts, then nbins histogram bins
// ------------------------------------------------------------------------------------------------
#define OBS_MAXS 32
struct ObsArgs { int ns, btype, nbins, bin_width; float rcsq; int s[OBS_MAXS]; };

__global__ void k_observables(Dev d, ObsArgs A, double *out_d, unsigned long long *out_i) {
  const int n = d.ctrl->nown;
  const int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int k = d.own0 + j;
    const int4 p = pos[k];
    const int t = p.w >> 3;
    const int im = d.img[k];
    const double m = (double)c_P.mass[p.w & 7];
    const double x = le_deq((unsigned)p.x, 0) + (double)((im & 1023) - 512) * c_P.L[0];
    const double y = le_deq((unsigned)p.y, 1) + (double)(((im >> 10) & 1023) - 512) * c_P.L[1];
    const double z = le_deq((unsigned)p.z, 2) + (double)(((im >> 20) & 1023) - 512) * c_P.L[2];
    acc[0] += m; acc[1] += m * x; acc[2] += m * y; acc[3] += m * z; acc[4] += m * (x * x + y * y + z * z);
    for (int q = 0; q < A.ns; q++) {
      const int t2 = t + A.s[q];
      if (t2 > d.N) continue;
      const int k2 = d.map[t2 - 1];
      if (k2 < 0) continue;                    // not on this GPU: farther away than the halo, hence than rc
      const int4 p2 = pos[k2];
      const float dx = (float)(int)((unsigned)p2.x - (unsigned)p.x) * c_P.fscale[0];
      const float dy = (float)(int)((unsigned)p2.y - (unsigned)p.y) * c_P.fscale[1];
      const float dz = (float)(int)((unsigned)p2.z - (unsigned)p.z) * c_P.fscale[2];
      if (dx * dx + dy * dy + dz * dz < A.rcsq) atomicAdd(&out_i[q], 1ull);
    }
    if (A.nbins > 0) {
      const int nb = d.num_bond[t - 1];
      for (int mth = 0; mth < nb; mth++) {
        if (d.bond_type[(size_t)(t - 1) * d.bpa + mth] != A.btype) continue;
        const int b = d.bond_atom[(size_t)(t - 1) * d.bpa + mth];
        if (b > t) atomicAdd(&out_i[A.ns + min((b - t) / A.bin_width, A.nbins - 1)], 1ull);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 5; q++) {
    const double sres = warp_sum(acc[q]);
    if ((threadIdx.x & 31) == 0 && sres != 0.0) atomicAdd(&out_d[q], sres);
  }
}
