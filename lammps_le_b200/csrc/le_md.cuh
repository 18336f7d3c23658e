// le_md.cuh -- the molecular-dynamics kernels: cell sort, neighbor/bond list build, fused step.
//
// Every phase is a __device__ function with a grid-stride loop plus a thin __global__ wrapper,
// so the same code runs as separate launches (profilable one by one) or inside one persistent
// cooperative kernel.
#pragma once
#include "le_common.cuh"

struct StepArgs {
  int rd;             // position buffer to read (the other one is written)
  int do_final;       // second half of velocity Verlet for the step whose forces are computed here
  int do_initial;     // first half of the next step (v += dtf f/m; x += dt v)
  int ev;             // tally energy / virial / kinetic energy into thermo slot `slot`
  int slot;
  int write_force;    // store the conservative force of every atom in fout (tag order)
  int langevin;       // add drag + noise
  unsigned step_lo, step_hi;   // timestep of this force evaluation (noise counter)
  float tsqrt;        // sqrt(target temperature) at this step (FixLangevin::compute_target)
};

// ------------------------------------------------------------------------------------------------
// fused step: WCA pair force over the full ELL rows + FENE/harmonic bond rows + Langevin + NVE.
//   reference: PairLJCut::compute src/pair_lj_cut.cpp:68-140, BondFENE::compute
//   src/MOLECULE/bond_fene.cpp:52-128, BondHarmonic::compute bond_harmonic.cpp:48-100,
//   FixLangevin::post_force_templated src/fix_langevin.cpp:587-777 (uniform noise, :672-675),
//   FixNVE::initial/final_integrate src/fix_nve.cpp:64-140, Neighbor::check_distance
//   src/neighbor.cpp:1962-2014.
// One thread owns one atom: it gathers its neighbors (each pair is evaluated from both sides, so
// no atomics and no force array), accumulates the force in fp64, finishes the velocity update of
// this step and starts the next one.  fp32 pair math, fixed-point minimum image for free.
// ------------------------------------------------------------------------------------------------
template <int EV>
__device__ __forceinline__ void step_phase(const Dev &d, const StepArgs &a) {
  const int N = d.N;
  const int4 *__restrict__ posr = d.pos[a.rd];
  int4 *__restrict__ posw = d.pos[a.rd ^ 1];
  const unsigned *__restrict__ neigh = d.neigh;
  const unsigned *__restrict__ bondrow = d.bondrow;
  const float sx = c_P.fscale[0], sy = c_P.fscale[1], sz = c_P.fscale[2];
  const int nt = c_P.ntypes;

  double acc[10];
  if (EV) {
#pragma unroll
    for (int q = 0; q < 10; q++) acc[q] = 0.0;
  }

  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int4 pi = posr[i];
    float4 vi = d.vel[i];
    const unsigned cnt = d.counts[i];
    const int ti = pi.w & 0xff;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    double evdwl = 0.0, ebond = 0.0;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0, v5 = 0.0;

    // ---- pair ----
    // fp32 distance screen on the exact fixed-point differences; the (few) pairs inside the force cutoff are
    // evaluated in fp64: r^-14 amplifies a 1e-7 error of r^2 sevenfold and the WCA/FENE terms of bonded
    // neighbours cancel to ~10% of their size, so fp32 pair math cannot meet the 1e-5 per-atom bar
    const int nn = cnt & 0xff;
#pragma unroll 4
    for (int k = 0; k < nn; k++) {
      const unsigned e = __ldg(&neigh[(size_t)k * N + i]);
      const int j = e & NEIGH_IDX_MASK;
      const int4 pj = __ldg(&posr[j]);
      const int idx = (int)((unsigned)pi.x - (unsigned)pj.x);
      const int idy = (int)((unsigned)pi.y - (unsigned)pj.y);
      const int idz = (int)((unsigned)pi.z - (unsigned)pj.z);
      const float dxf = (float)idx * sx, dyf = (float)idy * sy, dzf = (float)idz * sz;
      const float rsqf = dxf * dxf + dyf * dyf + dzf * dzf;
      const int tp = c_P.pair_uniform ? 0 : ti * nt + (pj.w & 0xff);
      if (rsqf < c_P.cutsq_screen[tp]) {
        const double dx = (double)idx * c_P.scale[0], dy = (double)idy * c_P.scale[1], dz = (double)idz * c_P.scale[2];
        const double rsq = dx * dx + dy * dy + dz * dz;
        if (rsq < c_P.cutsq_d[tp]) {
          double r2inv = (double)(1.0f / (float)rsq);
          r2inv = r2inv * (2.0 - rsq * r2inv);            // one Newton step: full double accuracy
          const double r6inv = r2inv * r2inv * r2inv;
          const double factor = (double)c_P.special_lj[e >> 30];
          const double fpair = factor * r6inv * (c_P.lj1_d[tp] * r6inv - c_P.lj2_d[tp]) * r2inv;
          fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
          if (EV) {
            evdwl += factor * (r6inv * (c_P.lj3_d[tp] * r6inv - c_P.lj4_d[tp]) - c_P.offset_d[tp]);
            v0 += dx * dx * fpair; v1 += dy * dy * fpair; v2 += dz * dz * fpair;
            v3 += dx * dy * fpair; v4 += dx * dz * fpair; v5 += dy * dz * fpair;
          }
        }
      }
    }
    double pv0 = v0, pv1 = v1, pv2 = v2, pv3 = v3, pv4 = v4, pv5 = v5;  // pair part (counted twice)

    // ---- bonds (fp64: see above) ----
    const int nb = (cnt >> 16) & 0xff;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0, b4 = 0.0, b5 = 0.0;
    for (int m = 0; m < nb; m++) {
      const unsigned e = __ldg(&bondrow[(size_t)m * N + i]);
      const int j = e & BOND_IDX_MASK;
      const int bt = e >> 28;
      const int4 pj = __ldg(&posr[j]);
      const double dx = (double)(int)((unsigned)pi.x - (unsigned)pj.x) * c_P.scale[0];
      const double dy = (double)(int)((unsigned)pi.y - (unsigned)pj.y) * c_P.scale[1];
      const double dz = (double)(int)((unsigned)pi.z - (unsigned)pj.z) * c_P.scale[2];
      const double rsq = dx * dx + dy * dy + dz * dz;
      double fbond;
      if (c_P.bstyle[bt] == 1) {  // FENE
        const double r0sq = c_P.br0_d[bt] * c_P.br0_d[bt];
        double rlogarg = 1.0 - rsq / r0sq;
        if (rlogarg < 0.1) {
          if (EV) acc[9] += 0.5;  // each long bond is seen from both ends
          if (rlogarg <= -3.0) le_raise(d.ctrl, LE_DERR_BAD_FENE, __float_as_int(vi.w), j);
          rlogarg = 0.1;
        }
        fbond = -c_P.bk_d[bt] / rlogarg;
        const double sig2 = c_P.bsig_d[bt] * c_P.bsig_d[bt];
        double sr6 = 0.0;
        const bool core = rsq < 1.2599210498948732 * sig2;   // TWO_1_3
        if (core) {
          const double sr2 = sig2 / rsq;
          sr6 = sr2 * sr2 * sr2;
          fbond += 48.0 * c_P.beps_d[bt] * sr6 * (sr6 - 0.5) / rsq;
        }
        if (EV) {
          double eb = -0.5 * c_P.bk_d[bt] * r0sq * log(rlogarg);
          if (core) eb += 4.0 * c_P.beps_d[bt] * sr6 * (sr6 - 1.0) + c_P.beps_d[bt];
          ebond += eb;
        }
      } else if (c_P.bstyle[bt] == 2) {  // harmonic
        const double r = sqrt(rsq);
        const double dr = r - c_P.br0_d[bt];
        const double rk = c_P.bk_d[bt] * dr;
        fbond = (r > 0.0) ? -2.0 * rk / r : 0.0;
        if (EV) ebond += rk * dr;
      } else {
        fbond = 0.0;
      }
      fx += dx * fbond; fy += dy * fbond; fz += dz * fbond;
      if (EV) {
        b0 += dx * dx * fbond; b1 += dy * dy * fbond; b2 += dz * dz * fbond;
        b3 += dx * dy * fbond; b4 += dx * dz * fbond; b5 += dy * dz * fbond;
      }
    }

    const int tag = __float_as_int(vi.w);
    if (a.write_force) {
      double *fo = d.fout + (size_t)(tag - 1) * 3;
      fo[0] = fx; fo[1] = fy; fo[2] = fz;
    }

    // ---- Langevin drag + uniform noise (post_force) ----
    if (a.langevin) {
      unsigned r[4];
      philox4x32_10((unsigned)tag, a.step_lo, a.step_hi, 0x4c45u, c_P.seed_lo, c_P.seed_hi, r);
      const float g1 = c_P.gfac1[ti], g2 = c_P.gfac2[ti] * a.tsqrt;
      const float u0 = (float)(r[0] >> 8) * 5.9604644775390625e-8f - 0.5f;
      const float u1 = (float)(r[1] >> 8) * 5.9604644775390625e-8f - 0.5f;
      const float u2 = (float)(r[2] >> 8) * 5.9604644775390625e-8f - 0.5f;
      fx += (double)(g1 * vi.x + g2 * u0);
      fy += (double)(g1 * vi.y + g2 * u1);
      fz += (double)(g1 * vi.z + g2 * u2);
    }

    // ---- velocity Verlet ----
    const float m = c_P.mass[ti];
    const float dtfm = c_P.dtf / m;
    const float ffx = (float)fx, ffy = (float)fy, ffz = (float)fz;
    if (a.do_final) {
      vi.x += dtfm * ffx; vi.y += dtfm * ffy; vi.z += dtfm * ffz;
      if (c_P.vlimitsq > 0.0f) {   // FixNVELimit::final_integrate
        const float vsq = vi.x * vi.x + vi.y * vi.y + vi.z * vi.z;
        if (vsq > c_P.vlimitsq) { const float sc = sqrtf(c_P.vlimitsq / vsq); vi.x *= sc; vi.y *= sc; vi.z *= sc; }
      }
    }
    if (EV) {
      acc[0] += (double)m * ((double)vi.x * vi.x + (double)vi.y * vi.y + (double)vi.z * vi.z);
      acc[1] += 0.5 * evdwl;
      acc[2] += 0.5 * ebond;
      acc[3] += 0.5 * (pv0 + b0); acc[4] += 0.5 * (pv1 + b1); acc[5] += 0.5 * (pv2 + b2);
      acc[6] += 0.5 * (pv3 + b3); acc[7] += 0.5 * (pv4 + b4); acc[8] += 0.5 * (pv5 + b5);
    }
    if (a.do_initial) {
      vi.x += dtfm * ffx; vi.y += dtfm * ffy; vi.z += dtfm * ffz;
      if (c_P.vlimitsq > 0.0f) {   // FixNVELimit::initial_integrate
        const float vsq = vi.x * vi.x + vi.y * vi.y + vi.z * vi.z;
        if (vsq > c_P.vlimitsq) { const float sc = sqrtf(c_P.vlimitsq / vsq); vi.x *= sc; vi.y *= sc; vi.z *= sc; }
      }
      const int dux = __float2int_rn(c_P.dt * vi.x * c_P.inv_fscale[0]);
      const int duy = __float2int_rn(c_P.dt * vi.y * c_P.inv_fscale[1]);
      const int duz = __float2int_rn(c_P.dt * vi.z * c_P.inv_fscale[2]);
      const unsigned nx = (unsigned)pi.x + (unsigned)dux;
      const unsigned ny = (unsigned)pi.y + (unsigned)duy;
      const unsigned nz = (unsigned)pi.z + (unsigned)duz;
      // image flags: a wrap of the 32-bit coordinate is a periodic crossing (Domain::remap)
      int wx = 0, wy = 0, wz = 0;
      if (dux > 0 && nx < (unsigned)pi.x) wx = 1; else if (dux < 0 && nx > (unsigned)pi.x) wx = -1;
      if (duy > 0 && ny < (unsigned)pi.y) wy = 1; else if (duy < 0 && ny > (unsigned)pi.y) wy = -1;
      if (duz > 0 && nz < (unsigned)pi.z) wz = 1; else if (duz < 0 && nz > (unsigned)pi.z) wz = -1;
      if (wx | wy | wz) {
        const int im = d.img[i];
        int ix = (im & 1023) - 512 + wx;
        int iy = ((im >> 10) & 1023) - 512 + wy;
        int iz = ((im >> 20) & 1023) - 512 + wz;  // 10+10+10 packing of LAMMPS_SMALLBIG (src/lmptype.h)
        d.img[i] = ((ix + 512) & 1023) | (((iy + 512) & 1023) << 10) | (((iz + 512) & 1023) << 20);
      }
      posw[i] = make_int4((int)nx, (int)ny, (int)nz, pi.w);
      // displacement since the last rebuild
      const int4 ph = d.pos_hold[i];
      const float hx = (float)(int)(nx - (unsigned)ph.x) * sx;
      const float hy = (float)(int)(ny - (unsigned)ph.y) * sy;
      const float hz = (float)(int)(nz - (unsigned)ph.z) * sz;
      if (hx * hx + hy * hy + hz * hz > c_P.triggersq) d.ctrl->moved = 1;
    }
    d.vel[i] = vi;
  }

  if (EV) {
    __shared__ double red[32][10];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 10; q++) {
      double s = warp_sum(acc[q]);
      if (lane == 0) red[warp][q] = s;
    }
    __syncthreads();
    if (warp == 0) {
      const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
      for (int q = 0; q < 10; q++) {
        double s = (lane < nw) ? red[lane][q] : 0.0;
        s = warp_sum(s);
        if (lane == 0 && s != 0.0) atomicAdd(&d.thermo[(size_t)a.slot * LE_THERMO_W + q], s);
      }
    }
  }
}

template <int EV>
__global__ void __launch_bounds__(256) k_step(Dev d, StepArgs a) { step_phase<EV>(d, a); }

// ------------------------------------------------------------------------------------------------
// Neighbor::decide (src/neighbor.cpp:1933-1948): one thread.
// ------------------------------------------------------------------------------------------------
__global__ void k_decide(Dev d) {
  Ctrl *c = d.ctrl;
  int r = 0;
  if (c->forced) r = 1;
  else {
    c->ago++;
    if (c->ago >= c_P.delay && c->ago % c_P.every == 0) {
      if (!c_P.check) r = 1;
      else if (c->moved) {
        r = 1;
        const int mx = c_P.every > c_P.delay ? c_P.every : c_P.delay;
        if (c->ago == mx) c->ndanger++;
      }
    }
  }
  c->rebuild_now = r;
  c->forced = 0;
}

// ------------------------------------------------------------------------------------------------
// cell sort.  Cells are at least one neighbor cutoff wide; the fixed-point coordinate gives the
// cell by one multiply-high.  Within a cell atoms are ordered by tag so that the sorted order --
// and with it every floating-point sum downstream -- is reproducible from run to run.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_of(const Dev &d, int4 p) {
  const int cx = __umulhi((unsigned)p.x, (unsigned)d.ncell[0]);
  const int cy = __umulhi((unsigned)p.y, (unsigned)d.ncell[1]);
  const int cz = __umulhi((unsigned)p.z, (unsigned)d.ncell[2]);
  return (cz * d.ncell[1] + cy) * d.ncell[0] + cx;
}

__global__ void k_cell_count(Dev d, int cur, int gated) {
  if (gated && !d.ctrl->rebuild_now) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const int c = cell_of(d, d.pos[cur][i]);
    d.cellid[i] = c;
    d.slot[i] = atomicAdd(&d.cell_count[c], 1);
  }
}

#define SCAN_BLOCK 1024
// exclusive scan of cell_count -> cell_start in three launches; also re-zeroes cell_count
__global__ void k_scan_partial(Dev d, int gated) {
  if (gated && !d.ctrl->rebuild_now) return;
  __shared__ int sh[32];
  const int base = blockIdx.x * SCAN_BLOCK;
  const int idx = base + threadIdx.x;
  int v = (idx < d.ncells) ? d.cell_count[idx] : 0;
  int s = v;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = sh[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) d.blocksum[blockIdx.x] = t;
  }
}

__global__ void k_scan_blocks(Dev d, int gated) {  // one block
  if (gated && !d.ctrl->rebuild_now) return;
  __shared__ int sh[SCAN_BLOCK];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < d.nscanblocks; base += SCAN_BLOCK) {
    const int idx = base + threadIdx.x;
    const int v = (idx < d.nscanblocks) ? d.blocksum[idx] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
      int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (idx < d.nscanblocks) d.blocksum[idx] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[SCAN_BLOCK - 1];
    __syncthreads();
  }
}

__global__ void k_scan_apply(Dev d, int gated) {
  if (gated && !d.ctrl->rebuild_now) return;
  __shared__ int sh[SCAN_BLOCK];
  const int idx = blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int v = (idx < d.ncells) ? d.cell_count[idx] : 0;
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
    int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  if (idx < d.ncells) {
    d.cell_start[idx] = d.blocksum[blockIdx.x] + sh[threadIdx.x] - v;
    d.cell_count[idx] = 0;
  }
  if (idx == d.ncells - 1) d.cell_start[d.ncells] = d.N;
}

__global__ void k_cell_scatter(Dev d, int gated) {
  if (gated && !d.ctrl->rebuild_now) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x)
    d.order[d.cell_start[d.cellid[i]] + d.slot[i]] = i;
}

// order each cell's members by tag (insertion sort; cells hold a handful of atoms)
__global__ void k_cell_sort(Dev d, int gated) {
  if (gated && !d.ctrl->rebuild_now) return;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d.ncells; c += gridDim.x * blockDim.x) {
    const int s = d.cell_start[c], e = d.cell_start[c + 1];
    for (int a = s + 1; a < e; a++) {
      const int ia = d.order[a];
      const int ta = __float_as_int(d.vel[ia].w);
      int b = a - 1;
      while (b >= s) {
        const int ib = d.order[b];
        if (__float_as_int(d.vel[ib].w) <= ta) break;
        d.order[b + 1] = ib;
        b--;
      }
      d.order[b + 1] = ia;
    }
  }
}

// gather into sorted order: pos_hold (= new xhold), vel_tmp, img_hold; refresh the tag map
__global__ void k_gather(Dev d, int cur, int gated) {
  if (gated && !d.ctrl->rebuild_now) return;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < d.N; k += gridDim.x * blockDim.x) {
    const int i = d.order[k];
    const float4 v = d.vel[i];
    d.pos_hold[k] = d.pos[cur][i];
    d.vel_tmp[k] = v;
    d.img_hold[k] = d.img[i];
    d.map[__float_as_int(v.w) - 1] = k;
    d.ex13[k] = 0;
  }
}

// NBin::coord2bin for one dimension (src/nbin.cpp:120-150)
__device__ __forceinline__ int ref_bin(double x, int dim) {
  const double lo = c_P.lo[dim], hi = c_P.hi[dim], inv = c_P.bininv[dim];
  const int nb = c_P.nbin[dim];
  int ix;
  if (x >= hi) ix = (int)__dmul_rn(__dsub_rn(x, hi), inv) + nb;
  else if (x >= lo) { ix = (int)__dmul_rn(__dsub_rn(x, lo), inv); ix = min(ix, nb - 1); }
  else ix = (int)__dmul_rn(__dsub_rn(x, lo), inv) - 1;
  return ix;
}

// NPair::find_special (src/npair.h:112-136)
__device__ __forceinline__ int find_special(const int *__restrict__ row, int n1, int n2, int n3, int nscan, int tag) {
  for (int k = 0; k < nscan; k++) {
    if (row[k] == tag) {
      const int tier = (k < n1) ? 1 : (k < n2) ? 2 : 3;
      const int f = c_P.special_flag[tier];
      if (f == 0) return -1;
      if (f == 1) return 0;
      return tier;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// neighbor + bond list build, one thread per atom of the sorted order.
//   Distance test in fp64 on the dequantised coordinates with the reference's operation order
//   (delx = xi - xj', rsq = dx*dx+dy*dy+dz*dz, rsq <= cutneighsq; npair_half_bin_newton.cpp:98-103)
//   so the list is bit-identical to NPairHalfBinNewton::build on the same coordinates.  Every
//   accepted pair goes into the full row; whether the reference would have stored it on THIS atom
//   (same-bin rule :84-91, upper-half stencil nstencil_half_bin_3d_newton.cpp:26-38) decides if it
//   sits in the leading "half" part of the row.
//   Also: bond partner rows (NTopoBondAll::build, src/ntopo_bond_all.cpp:39-86), periodic-crossing
//   flags, the (t,t+2) half-list membership used by fix ex_load, and the copy of the sorted state
//   back into the live arrays.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_build(Dev d, int cur, int gated) {
  if (gated && !d.ctrl->rebuild_now) return;
  const int N = d.N;
  const int4 *__restrict__ ph = d.pos_hold;
  const int nt = c_P.ntypes;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int4 pi = ph[i];
    const float4 vi = d.vel_tmp[i];
    d.pos[cur][i] = pi;
    d.vel[i] = vi;
    d.img[i] = d.img_hold[i];
    const int tagi = __float_as_int(vi.w);
    const int ti = pi.w & 0xff;
    const double xi = le_deq((unsigned)pi.x, 0), yi = le_deq((unsigned)pi.y, 1), zi = le_deq((unsigned)pi.z, 2);
    const int bix = ref_bin(xi, 0), biy = ref_bin(yi, 1), biz = ref_bin(zi, 2);
    const int *__restrict__ srow = d.special + (size_t)(tagi - 1) * d.maxspecial;
    const int n1 = d.nspecial[(tagi - 1) * 3], n2 = d.nspecial[(tagi - 1) * 3 + 1], n3 = d.nspecial[(tagi - 1) * 3 + 2];
    const int nscan = c_P.nscan_tier == 0 ? 0 : c_P.nscan_tier == 1 ? n1 : c_P.nscan_tier == 2 ? n2 : n3;

    const int cx = __umulhi((unsigned)pi.x, (unsigned)d.ncell[0]);
    const int cy = __umulhi((unsigned)pi.y, (unsigned)d.ncell[1]);
    const int cz = __umulhi((unsigned)pi.z, (unsigned)d.ncell[2]);
    int nh = 0, nbk = 0;
    const int maxn = d.maxneigh;
    unsigned *__restrict__ row = d.neigh + i;

    for (int oz = 0; oz < d.cell_span[2]; oz++) {
      int zc = d.cell_abs[2] ? oz : cz - d.cell_rad[2] + oz;   // a small box visits every cell once
      if (zc < 0) zc += d.ncell[2]; else if (zc >= d.ncell[2]) zc -= d.ncell[2];
      for (int oy = 0; oy < d.cell_span[1]; oy++) {
        int yc = d.cell_abs[1] ? oy : cy - d.cell_rad[1] + oy;
        if (yc < 0) yc += d.ncell[1]; else if (yc >= d.ncell[1]) yc -= d.ncell[1];
        for (int ox = 0; ox < d.cell_span[0]; ox++) {
          int xc = d.cell_abs[0] ? ox : cx - d.cell_rad[0] + ox;
          if (xc < 0) xc += d.ncell[0]; else if (xc >= d.ncell[0]) xc -= d.ncell[0];
          const int c = (zc * d.ncell[1] + yc) * d.ncell[0] + xc;
          const int js = d.cell_start[c], je = d.cell_start[c + 1];
          for (int j = js; j < je; j++) {
            if (j == i) continue;
            const int4 pj = __ldg(&ph[j]);
            // xj - xi as the minimum-image fixed-point difference
            const int idx = (int)((unsigned)pj.x - (unsigned)pi.x);
            const int idy = (int)((unsigned)pj.y - (unsigned)pi.y);
            const int idz = (int)((unsigned)pj.z - (unsigned)pi.z);
            const float fx = (float)idx * c_P.fscale[0], fy = (float)idy * c_P.fscale[1], fz = (float)idz * c_P.fscale[2];
            if (fx * fx + fy * fy + fz * fz > c_P.cutneighmaxsq_f) continue;   // coarse reject with margin
            // periodic shift of j's image: (xj - xi)_wrapped - (xj - xi)_raw = s * 2^32
            const int sxs = (int)(((long long)idx - ((long long)(unsigned)pj.x - (long long)(unsigned)pi.x)) >> 32);
            const int sys = (int)(((long long)idy - ((long long)(unsigned)pj.y - (long long)(unsigned)pi.y)) >> 32);
            const int szs = (int)(((long long)idz - ((long long)(unsigned)pj.z - (long long)(unsigned)pi.z)) >> 32);
            double xj = le_deq((unsigned)pj.x, 0), yj = le_deq((unsigned)pj.y, 1), zj = le_deq((unsigned)pj.z, 2);
            if (sxs) xj = __dadd_rn(xj, (double)sxs * c_P.L[0]);   // ghost coordinate (AtomVec::pack_border, x + pbc*prd)
            if (sys) yj = __dadd_rn(yj, (double)sys * c_P.L[1]);
            if (szs) zj = __dadd_rn(zj, (double)szs * c_P.L[2]);
            const double delx = __dsub_rn(xi, xj), dely = __dsub_rn(yi, yj), delz = __dsub_rn(zi, zj);
            const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(delx, delx), __dmul_rn(dely, dely)), __dmul_rn(delz, delz));
            const int tj = pj.w & 0xff;
            if (!(rsq <= c_P.cutneighsq[ti * nt + tj])) continue;
            const int tagj = __float_as_int(d.vel_tmp[j].w);
            int which = find_special(srow, n1, n2, n3, nscan, tagj);
            if (which != 0) {  // Domain::minimum_image_check (src/domain.h:156-161)
              if ((c_P.periodic[0] && fabs(delx) > c_P.half[0]) || (c_P.periodic[1] && fabs(dely) > c_P.half[1]) ||
                  (c_P.periodic[2] && fabs(delz) > c_P.half[2])) which = 0;
            }
            if (which < 0) continue;
            // would the reference store this pair on atom i?
            const int dbx = ref_bin(xj, 0) - bix, dby = ref_bin(yj, 1) - biy, dbz = ref_bin(zj, 2) - biz;
            bool mine;
            if ((dbx | dby | dbz) == 0) {
              if ((sxs | sys | szs) == 0) mine = tagj > tagi;       // owned j later in the bin's list
              else mine = !(zj < zi || (zj == zi && (yj < yi || (yj == yi && xj < xi))));  // ghost j
            } else {
              mine = dbz > 0 || (dbz == 0 && (dby > 0 || (dby == 0 && dbx > 0)));
            }
            if (nh + nbk >= maxn) { le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, maxn); continue; }
            const unsigned e = (unsigned)j | ((unsigned)which << 30);
            if (mine) {
              row[(size_t)nh * N] = e; nh++;
              const int dt = tagj - tagi;
              const int gh = (sxs | sys | szs) ? 4 : 0;   // stored neighbor is a periodic ghost
              if (dt == 2) d.ex13[tagi - 1] = 3 | gh;          // in list, stored on the lower tag
              else if (dt == -2) d.ex13[tagj - 1] = 1 | gh;    // in list, stored on the upper tag
            } else {
              row[(size_t)(maxn - 1 - nbk) * N] = e; nbk++;
            }
          }
        }
      }
    }
    for (int t = 0; t < nbk; t++) row[(size_t)(nh + t) * N] = row[(size_t)(maxn - 1 - t) * N];

    // bond partner rows
    const int nb = d.num_bond[tagi - 1];
    for (int m = 0; m < nb; m++) {
      const int pt = d.bond_atom[(size_t)(tagi - 1) * d.bpa + m];
      const int bt = d.bond_type[(size_t)(tagi - 1) * d.bpa + m];
      const int j = d.map[pt - 1];
      d.bondrow[(size_t)m * N + i] = (unsigned)j | ((unsigned)(bt - 1) << 28);
      const int4 pj = ph[j];
      // image of the partner closest to this atom, as shifts -1/0/+1 per dimension packed 2 bits each
      // (+1 bias; 21 = same image): Domain::closest_image picks a ghost exactly when a shift is non-zero
      const int s0 = (int)(((long long)(int)((unsigned)pj.x - (unsigned)pi.x) - ((long long)(unsigned)pj.x - (long long)(unsigned)pi.x)) >> 32);
      const int s1 = (int)(((long long)(int)((unsigned)pj.y - (unsigned)pi.y) - ((long long)(unsigned)pj.y - (long long)(unsigned)pi.y)) >> 32);
      const int s2 = (int)(((long long)(int)((unsigned)pj.z - (unsigned)pi.z) - ((long long)(unsigned)pj.z - (long long)(unsigned)pi.z)) >> 32);
      d.bond_cross[(size_t)(tagi - 1) * d.bpa + m] = (unsigned char)((s0 + 1) | ((s1 + 1) << 2) | ((s2 + 1) << 4));
    }
    d.counts[i] = (unsigned)(nh + nbk) | ((unsigned)nh << 8) | ((unsigned)nb << 16);
  }
}

// reset after a rebuild (gated): displacement flag, age
__global__ void k_after_build(Dev d, int gated) {
  if (gated && !d.ctrl->rebuild_now) return;
  d.ctrl->moved = 0;
  d.ctrl->forced = 0;
  d.ctrl->ago = 0;
  d.ctrl->nbuilds++;
}

// list statistics on demand
__global__ void k_count_pairs(Dev d, unsigned long long *out) {
  unsigned long long h = 0, f = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const unsigned c = d.counts[i];
    f += c & 0xff;
    h += (c >> 8) & 0xff;
  }
  for (int o = 16; o > 0; o >>= 1) {
    h += __shfl_xor_sync(0xffffffffu, h, o);
    f += __shfl_xor_sync(0xffffffffu, f, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], h); atomicAdd(&out[1], f); }
}
