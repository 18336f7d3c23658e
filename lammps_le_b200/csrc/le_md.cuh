// le_md.cuh -- the molecular-dynamics kernels: fused step, reneighbor decision, cell sort, list build.
//
// Launch structure of one timestep (le_engine.cu): k_step -> k_decide -> [conditional graph node:
// k_cell_count .. k_build].  The step number and the position-buffer parity live in the device-side
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
  int rdp1;           // 1 + position buffer holding the current coordinates when the host knows it, 0 = read Ctrl::cur (k_step2)
  int fuse;           // k_step2p inside the steady-state graph: the last block to finish also closes the timestep and takes the
                      // reneighbor decision of the next one (k_decide's work), switching the conditional node `handle`
  unsigned long long handle;
  int skip;           // development only (LE_STEP_SKIP): 1 no gathers, 2 no pair evaluation, 4 no bonds, 8 no neighbor rows, 32 no boundary-first block order
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
// this step and starts the next one.  Distances come from exact differences of the 32-bit
// fixed-point coordinates (minimum image for free); the first four neighbor and three bond slots
// are fetched as one batch of independent loads so the index -> position dependency is paid once.
// ------------------------------------------------------------------------------------------------
struct ForceAcc {
  double fx, fy, fz;
  float px, py, pz;     // fp32 partial sum of the pair forces (Params::pair32)
  double evdwl, ebond;
  double pv[6], bv[6];
  double warn;
};

// 1/x to full double accuracy: MUFU.RCP64H seed (rcp.approx.ftz.f64, ~20 bits) + two Newton steps
__device__ __forceinline__ double le_rcp(double x) {
  double t;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(x));
  t = t * (2.0 - x * t);
  t = t * (2.0 - x * t);
  return t;
}

// harmonic bond (bond_harmonic.cpp:71-80), kept out of line: sqrt and a division in fp64 are long code and
// only the extruder bonds of some decks use the style
__device__ __noinline__ double harmonic_fbond(double rsq, double k, double r0, double *eb) {
  const double r = sqrt(rsq);
  const double dr = r - r0;
  const double rk = k * dr;
  *eb = rk * dr;
  return (r > 0.0) ? -2.0 * rk / r : 0.0;
}

// fp32 screen of one listed pair on the exact fixed-point differences: inside (a hair more than) the force cutoff?
template <int UNI>
__device__ __forceinline__ bool pair_screen(const int4 pi, const int4 pj, int ti, int nt, float sx, float sy, float sz) {
  const int idx = (int)((unsigned)pi.x - (unsigned)pj.x);
  const int idy = (int)((unsigned)pi.y - (unsigned)pj.y);
  const int idz = (int)((unsigned)pi.z - (unsigned)pj.z);
  const float dxf = (float)idx * sx, dyf = (float)idy * sy, dzf = (float)idz * sz;
  const float rsqf = dxf * dxf + dyf * dyf + dzf * dzf;
  const int tp = UNI ? 0 : ti * nt + (pj.w & 7);      // UNI: one coefficient set for all type pairs -> immediate constant operands
  return rsqf < c_P.cutsq_screen[tp];
}

// the (few) pairs inside the force cutoff are evaluated in fp64: r^-14 amplifies a 1e-7 error of r^2 sevenfold and
// the WCA/FENE terms of bonded neighbours cancel to ~10% of their size, so fp32 pair math cannot meet the 1e-5
// per-atom bar
template <int EV, int UNI>
__device__ __forceinline__ void pair_term(ForceAcc &A, const int4 pi, const int4 pj, unsigned e, int ti, int nt) {
  const int idx = (int)((unsigned)pi.x - (unsigned)pj.x);
  const int idy = (int)((unsigned)pi.y - (unsigned)pj.y);
  const int idz = (int)((unsigned)pi.z - (unsigned)pj.z);
  const int tp = UNI ? 0 : ti * nt + (pj.w & 7);
  const double dx = (double)idx * c_P.scale[0], dy = (double)idy * c_P.scale[1], dz = (double)idz * c_P.scale[2];
  const double rsq = dx * dx + dy * dy + dz * dz;
  if (rsq < c_P.cutsq_d[tp]) {
    const double r2inv = le_rcp(rsq);
    const double r6inv = r2inv * r2inv * r2inv;
    const double factor = (double)c_P.special_lj[e >> 30];
    const double fpair = factor * r6inv * (c_P.lj1_d[tp] * r6inv - c_P.lj2_d[tp]) * r2inv;
    A.fx += dx * fpair; A.fy += dy * fpair; A.fz += dz * fpair;
    if (EV) {
      A.evdwl += factor * (r6inv * (c_P.lj3_d[tp] * r6inv - c_P.lj4_d[tp]) - c_P.offset_d[tp]);
      A.pv[0] += dx * dx * fpair; A.pv[1] += dy * dy * fpair; A.pv[2] += dz * dz * fpair;
      A.pv[3] += dx * dy * fpair; A.pv[4] += dx * dz * fpair; A.pv[5] += dy * dz * fpair;
    }
  }
}

// The same pair term in fp32 (Params::pair32; "fp32 pair math, fp64 accumulation"): listed pairs are never bonded
// (special_bonds 0 x x drops the 1-2 pairs), so there is no FENE/WCA cancellation to protect and r^-14 turns the 1e-7 of an
// fp32 r^2 into ~1e-6 of the pair force.  An atom's pair terms are summed in fp32 in row order (rows 0, 1, 2, ...), the
// sum then joins the fp64 bond terms.  Shared by k_step and k_step2 so that both give the same bits.
template <int EV, int UNI>
__device__ __forceinline__ void pair_term32(ForceAcc &A, const int4 pi, const int4 pj, unsigned e, int ti, int nt, float sx, float sy, float sz) {
  const float dxf = __fmul_rn((float)(int)((unsigned)pi.x - (unsigned)pj.x), sx);
  const float dyf = __fmul_rn((float)(int)((unsigned)pi.y - (unsigned)pj.y), sy);
  const float dzf = __fmul_rn((float)(int)((unsigned)pi.z - (unsigned)pj.z), sz);
  const float rsqf = __fmaf_rn(dzf, dzf, __fmaf_rn(dxf, dxf, __fmul_rn(dyf, dyf)));
  const int tp = UNI ? 0 : ti * nt + (pj.w & 7);
  if (rsqf < c_P.cutsq[tp]) {
    const float r2inv = __frcp_rn(rsqf);
    const float r6inv = __fmul_rn(__fmul_rn(r2inv, r2inv), r2inv);
    const float factor = c_P.special_lj[e >> 30];
    const float fpair = __fmul_rn(__fmul_rn(__fmul_rn(factor, r6inv), __fmaf_rn(c_P.lj1[tp], r6inv, -c_P.lj2[tp])), r2inv);
    A.px = __fmaf_rn(dxf, fpair, A.px); A.py = __fmaf_rn(dyf, fpair, A.py); A.pz = __fmaf_rn(dzf, fpair, A.pz);
    if (EV) {
      A.evdwl += (double)(factor * (r6inv * (c_P.lj3[tp] * r6inv - c_P.lj4[tp]) - c_P.offset[tp]));
      A.pv[0] += (double)(dxf * dxf * fpair); A.pv[1] += (double)(dyf * dyf * fpair); A.pv[2] += (double)(dzf * dzf * fpair);
      A.pv[3] += (double)(dxf * dyf * fpair); A.pv[4] += (double)(dxf * dzf * fpair); A.pv[5] += (double)(dyf * dzf * fpair);
    }
  }
}

template <int EV>
__device__ __forceinline__ void bond_term(ForceAcc &A, const Dev &d, const int4 pi, const int4 pj, unsigned e, int tagi) {
  const int bt = e >> 28;
  const double dx = (double)(int)((unsigned)pi.x - (unsigned)pj.x) * c_P.scale[0];
  const double dy = (double)(int)((unsigned)pi.y - (unsigned)pj.y) * c_P.scale[1];
  const double dz = (double)(int)((unsigned)pi.z - (unsigned)pj.z) * c_P.scale[2];
  const double rsq = dx * dx + dy * dy + dz * dz;
  double fbond;
  const int style = c_P.bstyle[bt];
  if (style == 1) {  // FENE (bond_fene.cpp:79-117)
    double rlogarg = 1.0 - rsq * c_P.binvr0sq_d[bt];
    if (rlogarg < 0.1) {
      if (EV) A.warn += 0.5;  // each long bond is seen from both ends
      if (rlogarg <= -3.0) le_raise(d.ctrl, LE_DERR_BAD_FENE, tagi, (int)(e & BOND_IDX_MASK));
      rlogarg = 0.1;
    }
    // one reciprocal serves both 1/rlogarg and 1/rsq
    const double q = rlogarg * rsq;
    const double t = le_rcp(q);
    const double inv_rl = t * rsq, inv_rsq = t * rlogarg;
    fbond = -c_P.bk_d[bt] * inv_rl;
    double sr6 = 0.0;
    const bool core = rsq < c_P.bcore_d[bt];
    if (core) {
      const double sr2 = c_P.bsig2_d[bt] * inv_rsq;
      sr6 = sr2 * sr2 * sr2;
      fbond += 48.0 * c_P.beps_d[bt] * sr6 * (sr6 - 0.5) * inv_rsq;
    }
    if (EV) {
      double eb = -0.5 * c_P.bk_d[bt] * c_P.br0sq_d[bt] * log(rlogarg);
      if (core) eb += 4.0 * c_P.beps_d[bt] * sr6 * (sr6 - 1.0) + c_P.beps_d[bt];
      A.ebond += eb;
    }
  } else if (style == 2) {  // harmonic (bond_harmonic.cpp:71-80)
    double eb;
    fbond = harmonic_fbond(rsq, c_P.bk_d[bt], c_P.br0_d[bt], &eb);
    if (EV) A.ebond += eb;
  } else {
    fbond = 0.0;
  }
  A.fx += dx * fbond; A.fy += dy * fbond; A.fz += dz * fbond;
  if (EV) {
    A.bv[0] += dx * dx * fbond; A.bv[1] += dy * dy * fbond; A.bv[2] += dz * dz * fbond;
    A.bv[3] += dx * dy * fbond; A.bv[4] += dx * dz * fbond; A.bv[5] += dy * dz * fbond;
  }
}

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

#define STEP_THREADS 256
#define STEP_NB 4     // neighbor slots fetched in the first batch
#define STEP_BB 3     // bond slots fetched in the first batch

// P32: pair terms in fp32 (Params::pair32; a separate instantiation so that the fp64 kernels stay as they were measured)
template <int EV, int DD, int MINB = 4, int UNI = 0, int P32 = 0>
__global__ void __launch_bounds__(STEP_THREADS, EV ? 2 : MINB) k_step(Dev d, StepArgs a) {
  const int cap = d.cap;
  Ctrl *__restrict__ ctrl = d.ctrl;
  const int rd = ctrl->cur;
  const long long step = ctrl->step;
  const int own_end = d.own0 + ctrl->nown;
  const int4 *__restrict__ posr = d.pos[rd];
  int4 *__restrict__ posw = d.pos[rd ^ 1];
  const unsigned *__restrict__ neigh = d.neigh;
  const unsigned *__restrict__ bondrow = d.bondrow;
  const float sx = c_P.fscale[0], sy = c_P.fscale[1], sz = c_P.fscale[2];
  const int nt = c_P.ntypes;
  int i = d.own0 + blockIdx.x * STEP_THREADS + threadIdx.x;
  if (DD && !(a.skip & 32)) {
    // the two boundary slices first, the interior last: their halo stores are in flight while the interior computes.
    // Segment starts are kept multiples of 64 slots (own0 is one), so that every warp still reads whole 128-byte
    // lines: the left slice is widened to the next multiple, the right one starts at the previous multiple.
    const int g = blockIdx.x * STEP_THREADS + threadIdx.x;
    const int a_end = min(own_end, d.own0 + ((ctrl->send_l_end - d.own0 + 63) & ~63));
    const int b_beg = max(a_end, d.own0 + ((ctrl->send_r_beg - d.own0) & ~63));
    const int nl = a_end - d.own0, nr = own_end - b_beg, nrp = (nr + 63) & ~63;
    if (g < nl) i = d.own0 + g;
    else if (g < nl + nrp) i = (g - nl < nr) ? b_beg + (g - nl) : own_end;
    else { i = a_end + (g - nl - nrp); if (i >= b_beg) i = own_end; }
  }

  double acc[10];
  if (EV) {
#pragma unroll
    for (int q = 0; q < 10; q++) acc[q] = 0.0;
  }

  if (i < own_end) {
    // ---- batch 1: everything addressed by i ----
    const int4 pi = posr[i];
    float4 vi = d.vel[i];
    const unsigned cnt = d.counts[i];
    unsigned en[STEP_NB], eb[STEP_BB];
#pragma unroll
    for (int k = 0; k < STEP_NB; k++) en[k] = __ldg(&neigh[(size_t)k * cap + i]);       // rows exist up to maxneigh >= 4
#pragma unroll
    for (int m = 0; m < STEP_BB; m++) eb[m] = (m < d.bpa) ? __ldg(&bondrow[(size_t)m * cap + i]) : 0u;
    const int4 ph = d.pos_hold[i];
    int nn = cnt & 0xff, nb = (cnt >> 16) & 0xff;
    if (a.skip & 8) nn = 0;
    if (a.skip & 4) nb = 0;
    const int ti = pi.w & 7;
    const int tag = pi.w >> 3;
    // ---- batch 2: the gathers ----
    int4 pn[STEP_NB], pb[STEP_BB];
#pragma unroll
    for (int k = 0; k < STEP_NB; k++) pn[k] = __ldg(&posr[(k < nn && !(a.skip & 1)) ? (int)(en[k] & NEIGH_IDX_MASK) : i]);
#pragma unroll
    for (int m = 0; m < STEP_BB; m++) pb[m] = __ldg(&posr[(m < nb && !(a.skip & 1)) ? (int)(eb[m] & BOND_IDX_MASK) : i]);

    ForceAcc A;
    A.fx = A.fy = A.fz = 0.0;
    if (EV) {
      A.evdwl = A.ebond = A.warn = 0.0;
#pragma unroll
      for (int q = 0; q < 6; q++) { A.pv[q] = 0.0; A.bv[q] = 0.0; }
    }
    if (P32) {
      A.px = A.py = A.pz = 0.f;
      // pair terms in fp32, rows in ascending order (pair_term32)
#pragma unroll
      for (int k = 0; k < STEP_NB; k++)
        if (k < nn && pair_screen<UNI>(pi, pn[k], ti, nt, sx, sy, sz)) pair_term32<EV, UNI>(A, pi, pn[k], en[k], ti, nt, sx, sy, sz);
      for (int k = STEP_NB; k < nn; k++) {
        const unsigned e0 = __ldg(&neigh[(size_t)k * cap + i]);
        const int4 p0 = __ldg(&posr[e0 & NEIGH_IDX_MASK]);
        if (pair_screen<UNI>(pi, p0, ti, nt, sx, sy, sz)) pair_term32<EV, UNI>(A, pi, p0, e0, ti, nt, sx, sy, sz);
      }
      A.fx = (double)A.px; A.fy = (double)A.py; A.fz = (double)A.pz;
      nn = 0;                                   // nothing left for the fp64 pair code below
    }
    // screen every listed pair in fp32; the survivors (about one pair in five) are queued as a bit mask and
    // evaluated by one fp64 loop, so a warp runs the fp64 code max-over-lanes(#survivors) times, not once per slot
    unsigned hit = 0;
#pragma unroll
    for (int k = 0; k < STEP_NB; k++)
      if (k < nn && pair_screen<UNI>(pi, pn[k], ti, nt, sx, sy, sz)) hit |= 1u << k;
    for (int k = STEP_NB; k < nn; k += 2) {          // rows beyond the first batch, two at a time
      const int k1 = min(k + 1, nn - 1);
      const unsigned e0 = __ldg(&neigh[(size_t)k * cap + i]), e1 = __ldg(&neigh[(size_t)k1 * cap + i]);
      const int4 p0 = __ldg(&posr[e0 & NEIGH_IDX_MASK]), p1 = __ldg(&posr[e1 & NEIGH_IDX_MASK]);
      if (pair_screen<UNI>(pi, p0, ti, nt, sx, sy, sz)) pair_term<EV, UNI>(A, pi, p0, e0, ti, nt);
      if (k1 > k && pair_screen<UNI>(pi, p1, ti, nt, sx, sy, sz)) pair_term<EV, UNI>(A, pi, p1, e1, ti, nt);
    }
    if (a.skip & 2) hit = 0;
    while (hit) {
      const int k = __ffs(hit) - 1;
      hit &= hit - 1;
      const unsigned e = k == 0 ? en[0] : k == 1 ? en[1] : k == 2 ? en[2] : en[3];
      const int4 pj = __ldg(&posr[e & NEIGH_IDX_MASK]);     // second touch: an L1 hit
      pair_term<EV, UNI>(A, pi, pj, e, ti, nt);
    }
#pragma unroll
    for (int m = 0; m < STEP_BB; m++)
      if (m < nb) bond_term<EV>(A, d, pi, pb[m], eb[m], tag);
    for (int m = STEP_BB; m < nb; m++) {
      const unsigned e = __ldg(&bondrow[(size_t)m * cap + i]);
      const int4 pj = __ldg(&posr[e & BOND_IDX_MASK]);
      bond_term<EV>(A, d, pi, pj, e, tag);
    }
    double fx = A.fx, fy = A.fy, fz = A.fz;

    if (a.write_force) {
      double *fo = d.fout + (size_t)(tag - 1) * 3;
      fo[0] = fx; fo[1] = fy; fo[2] = fz;
    }

    // ---- Langevin drag + uniform noise (post_force); fp32, added to the rounded conservative force ----
    float lx = 0.f, ly = 0.f, lz = 0.f;
    if (a.langevin) {
      unsigned r[4];
      philox4x32_7((unsigned)tag, (unsigned)(step & 0xffffffffll), (unsigned)((unsigned long long)step >> 32), 0x4c45u,
                    c_P.seed_lo, c_P.seed_hi, r);
      // FixLangevin::compute_target (src/fix_langevin.cpp:784-820): linear ramp over the run
      float tsq = c_P.tsqrt_const;
      if (c_P.t_start != c_P.t_stop) {
        float delta = (float)(step - ctrl->run_begin);
        if (delta != 0.0f) delta /= (float)(ctrl->run_end - ctrl->run_begin);
        tsq = sqrtf(c_P.t_start + delta * (c_P.t_stop - c_P.t_start));
      }
      const float g1 = c_P.gfac1[ti], g2 = c_P.gfac2[ti] * tsq;
      const float u0 = (float)(r[0] >> 8) * 5.9604644775390625e-8f - 0.5f;
      const float u1 = (float)(r[1] >> 8) * 5.9604644775390625e-8f - 0.5f;
      const float u2 = (float)(r[2] >> 8) * 5.9604644775390625e-8f - 0.5f;
      lx = g1 * vi.x + g2 * u0;
      ly = g1 * vi.y + g2 * u1;
      lz = g1 * vi.z + g2 * u2;
    }

    // ---- velocity Verlet ----
    const float m = c_P.mass[ti];
    const float dtfm = c_P.dtfm[ti];
    const float ffx = (float)fx + lx, ffy = (float)fy + ly, ffz = (float)fz + lz;
    if (a.do_final) {
      vi.x += dtfm * ffx; vi.y += dtfm * ffy; vi.z += dtfm * ffz;
      if (c_P.vlimitsq > 0.0f) {   // FixNVELimit::final_integrate
        const float vsq = vi.x * vi.x + vi.y * vi.y + vi.z * vi.z;
        if (vsq > c_P.vlimitsq) { const float sc = sqrtf(c_P.vlimitsq / vsq); vi.x *= sc; vi.y *= sc; vi.z *= sc; }
      }
    }
    if (EV) {
      acc[0] = (double)m * ((double)vi.x * vi.x + (double)vi.y * vi.y + (double)vi.z * vi.z);
      acc[1] = 0.5 * A.evdwl;
      acc[2] = 0.5 * A.ebond;
#pragma unroll
      for (int q = 0; q < 6; q++) acc[3 + q] = 0.5 * (A.pv[q] + A.bv[q]);   // every pair / bond is seen from both ends
      acc[9] = A.warn;
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
      const int4 pnew = make_int4((int)nx, (int)ny, (int)nz, pi.w);
      posw[i] = pnew;
      // halo update fused into the integrator: atoms of the slab's boundary layers are also stored straight into
      // the neighbor GPU's ghost slots over NVLink (the receiving slot was fixed at the last rebuild)
      if (DD) {
        if (i < ctrl->send_l_end) d.peer[left_rank(d)].pos[rd ^ 1][d.gr0 + (i - d.own0)] = pnew;
        const int srb = ctrl->send_r_beg;
        if (i >= srb) d.peer[right_rank(d)].pos[rd ^ 1][i - srb] = pnew;
      }
      // displacement since the last rebuild
      const float hx = (float)(int)(nx - (unsigned)ph.x) * sx;
      const float hy = (float)(int)(ny - (unsigned)ph.y) * sy;
      const float hz = (float)(int)(nz - (unsigned)ph.z) * sz;
      if (hx * hx + hy * hy + hz * hz > c_P.triggersq) ctrl->moved = 1;
    }
    d.vel[i] = vi;
  }

  if (EV) {
    __shared__ double red[STEP_THREADS / 32][10];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 10; q++) {
      double s = warp_sum(acc[q]);
      if (lane == 0) red[warp][q] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int q = 0; q < 10; q++) {
        double s = (lane < STEP_THREADS / 32) ? red[lane][q] : 0.0;
        s = warp_sum(s);
        if (lane == 0 && s != 0.0) atomicAdd(&d.thermo[(size_t)a.slot * LE_THERMO_W + q], s);
      }
    }
  }

}

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
  if (r) { moved = 0; ago = 0; c->nbuilds++; }
  *reinterpret_cast<int4 *>(c) = make_int4(moved, 0, r, ago);
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

#define SCAN_BLOCK 1024
#define SCAN_ITEMS 4                          // consecutive cells per thread
#define SCAN_TILE (SCAN_BLOCK * SCAN_ITEMS)
// exclusive scan of cell_count over the owned region's cell slots -> cell_start (three launches); re-zeroes cell_count
__device__ __forceinline__ int own_cell_first(const Dev &d) { return cell_slot(d, d.halo, 0, 0); }
__device__ __forceinline__ int own_cell_count(const Dev &d) { return (d.nlx - 2 * d.halo) * d.ncell[1] * d.ncell[2]; }

__global__ void k_scan_partial(Dev d) {
  __shared__ int sh[32];
  const int idx = (blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
  const int n = own_cell_count(d), first = own_cell_first(d);
  int s = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) s += (idx + q < n) ? d.cell_count[first + idx + q] : 0;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = sh[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) d.blocksum[blockIdx.x] = t;
  }
}

// block-wide exclusive scan of one value per thread (SCAN_BLOCK threads): warp shuffles + one shared pass
__device__ __forceinline__ int block_excl_scan(int v, int *total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const int base = warp ? wsum[warp - 1] : 0;
  if (total) *total = wsum[31];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_blocks(Dev d) {  // one block
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int nblocks = (own_cell_count(d) + SCAN_TILE - 1) / SCAN_TILE;
  for (int base = 0; base < nblocks; base += SCAN_BLOCK) {
    const int idx = base + threadIdx.x;
    const int v = (idx < nblocks) ? d.blocksum[idx] : 0;
    int tot;
    const int ex = block_excl_scan(v, &tot);
    if (idx < nblocks) d.blocksum[idx] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  // the owned population after migration; the sentinel slot behind the owned region closes its last cell
  if (threadIdx.x == 0) {
    d.ctrl->nown = carry;
    d.cell_start[own_cell_first(d) + own_cell_count(d)] = d.own0 + carry;
  }
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(Dev d) {
  const int idx = (blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
  const int n = own_cell_count(d), first = own_cell_first(d);
  int v[SCAN_ITEMS], tsum = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) { v[q] = (idx + q < n) ? d.cell_count[first + idx + q] : 0; tsum += v[q]; }
  int run = d.own0 + d.blocksum[blockIdx.x] + block_excl_scan(tsum, nullptr);
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++)
    if (idx + q < n) {
      d.cell_start[first + idx + q] = run;
      d.cell_count[first + idx + q] = 0;
      run += v[q];
    }
}

__global__ void k_cell_scatter(Dev d) {
  const int lo = d.own0, hi = d.own0 + (d.nranks > 1 ? d.ctrl->nown_unsorted : d.N);
  for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
    const int c = d.cellid[i];
    if (c >= 0) d.order[d.cell_start[c] + d.slot[i]] = i;
  }
}

// gather into the new local order: pos_hold (= new xhold), vel_tmp, img_hold; refresh the tag map.
// Inside a cell atoms are ordered by tag: slot k of a cell takes the member whose tag has rank k - cell_start among
// the cell's members (cells hold a handful of atoms, so every slot simply ranks them all -- no separate sort pass).
__global__ void k_gather(Dev d) {
  const int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  const int lo = d.own0, hi = d.own0 + d.ctrl->nown;
  for (int k = lo + blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += gridDim.x * blockDim.x) {
    int i = d.order[k];
    const int c = d.cellid[i];
    const int s = d.cell_start[c], e = d.cell_start[c + 1];
    if (e - s > 1) {
      const int want = k - s;
      for (int a = s; a < e; a++) {
        const int ia = d.order[a];
        const int ta = pos[ia].w >> 3;
        int rank = 0;
        for (int b = s; b < e; b++) rank += (pos[d.order[b]].w >> 3) < ta;
        if (rank == want) { i = ia; break; }
      }
    }
    const int4 p = pos[i];
    d.pos_hold[k] = p;
    d.vel_tmp[k] = d.vel[i];
    d.img_hold[k] = d.img[i];
    d.map[(p.w >> 3) - 1] = k;
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

// NPair::find_special (src/npair.h:112-136) on a row whose first four entries are already in registers
__device__ __forceinline__ int find_special(const int *__restrict__ row, int s0, int s1, int s2, int s3,
                                            int n1, int n2, int nscan, int tag) {
  int k = -1;
  if (nscan > 0 && s0 == tag) k = 0;
  else if (nscan > 1 && s1 == tag) k = 1;
  else if (nscan > 2 && s2 == tag) k = 2;
  else if (nscan > 3 && s3 == tag) k = 3;
  else
    for (int q = 4; q < nscan; q++)
      if (row[q] == tag) { k = q; break; }
  if (k < 0) return 0;
  const int tier = (k < n1) ? 1 : (k < n2) ? 2 : 3;
  const int f = c_P.special_flag[tier];
  if (f == 0) return -1;
  if (f == 1) return 0;
  return tier;
}

// rare path of the list build: a pair whose fp32 distance falls in the sliver around cutneighsq is decided by
// the reference's fp64 arithmetic
__device__ __noinline__ int build_border(int4 pi, int4 pj, int tp) {
  const unsigned ui[3] = {(unsigned)pi.x, (unsigned)pi.y, (unsigned)pi.z};
  const unsigned uj[3] = {(unsigned)pj.x, (unsigned)pj.y, (unsigned)pj.z};
  return le_pair_rsq_ref(c_P, ui, uj) <= c_P.cutneighsq[tp];
}

// ------------------------------------------------------------------------------------------------
// rebuild, part 3: neighbor + bond list build, one thread per owned atom.
//   The atom's 3x3 columns of cells are walked one after the other (the three z-cells of a column are
//   contiguous in the local order, so a column is one slot range); positions carry tag and type, so a
//   candidate costs one 16-byte load.  The distance is taken in
//   fp32 on the exact fixed-point differences; only the 1e-5 sliver around cutneighsq is re-evaluated
//   in fp64 on the dequantised coordinates with the reference's operation order (delx = xi - xj',
//   rsq = dx*dx+dy*dy+dz*dz, rsq <= cutneighsq; npair_half_bin_newton.cpp:98-103), so the pair set is
//   bit-identical to NPairHalfBinNewton::build on the same coordinates.  Every accepted pair goes into
//   the full row of BOTH atoms with the special-bond bits of find_special; which of the two the
//   reference's half list stores it on (same-bin rule :84-91, upper-half stencil
//   nstencil_half_bin_3d_newton.cpp:26-38) is only needed for the (t,t+2) pairs fix ex_load scans
//   (derived there from pos_hold) and for le_download_neighlist.
//   Domain::minimum_image_check (npair_half_bin_newton.cpp:111) cannot fire here: the difference is
//   the minimum image by construction and the box is at least two neighbor cutoffs wide.
//   Also: bond partner rows (NTopoBondAll::build, src/ntopo_bond_all.cpp:39-86) and the copy of the
//   sorted state back into the live arrays.
// ------------------------------------------------------------------------------------------------
#define BUILD_THREADS 128
#define BUILD_QUEUE 12

struct BuildCtx {
  unsigned *row;
  const int *srow;
  int s0, s1, s2, s3, n1, n2, nscan;
  int tagi, ti, nt, maxn, cap, n;
};

// decide one screened candidate and append it to the row
__device__ __forceinline__ void build_accept(const Dev &d, BuildCtx &B, const int4 pi, const int4 pj, int j, float rsqf) {
  const int tp = c_P.pair_uniform ? 0 : B.ti * B.nt + (pj.w & 7);
  if (rsqf > c_P.cutneigh_hi[tp]) return;
  const int tagj = pj.w >> 3;
  const int which = find_special(B.srow, B.s0, B.s1, B.s2, B.s3, B.n1, B.n2, B.nscan, tagj);
  if (which < 0) return;
  if (rsqf >= c_P.cutneigh_lo[tp] && !build_border(pi, pj, tp)) return;
  if (B.n >= B.maxn) { le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, B.tagi, B.maxn); return; }
  B.row[(size_t)B.n * B.cap] = (unsigned)j | ((unsigned)which << 30);
  B.n++;
}

// (4 candidates per trip of the inner loop: 2 or 8 measured 30 % slower, profiles/r01_step_variants.txt)
template <int MINB>
__global__ void __launch_bounds__(BUILD_THREADS, MINB) k_build(Dev d) {
  __shared__ int s_q[BUILD_QUEUE][BUILD_THREADS];
  const int cap = d.cap;
  const int cur = d.ctrl->cur;
  const int4 *__restrict__ ph = d.pos_hold;
  const int t = threadIdx.x;
  const int i = d.own0 + blockIdx.x * BUILD_THREADS + t;
  if (i >= d.own0 + d.ctrl->nown) return;
  const int4 pi = ph[i];
  BuildCtx B;
  B.tagi = pi.w >> 3; B.ti = pi.w & 7; B.nt = c_P.ntypes; B.maxn = d.maxneigh; B.cap = cap; B.n = 0;
  B.row = d.neigh + i;
  B.srow = d.special + (size_t)(B.tagi - 1) * d.maxspecial;
  // bond table of this atom: issue the tag-order loads now, they are consumed after the candidate scan
  const int tagi = B.tagi;
  const int nb = d.num_bond[tagi - 1];
  int bpart[4] = {0, 0, 0, 0}, btyp[4] = {0, 0, 0, 0};
  if (d.bpa == 4) {
    const int4 a4 = *(const int4 *)(d.bond_atom + (size_t)(tagi - 1) * 4);
    const int4 t4 = *(const int4 *)(d.bond_type + (size_t)(tagi - 1) * 4);
    bpart[0] = a4.x; bpart[1] = a4.y; bpart[2] = a4.z; bpart[3] = a4.w;
    btyp[0] = t4.x; btyp[1] = t4.y; btyp[2] = t4.z; btyp[3] = t4.w;
  }
  d.pos[cur][i] = pi;
  d.vel[i] = d.vel_tmp[i];
  d.img[i] = d.img_hold[i];
  {
    const int *ns = d.nspecial + (size_t)(B.tagi - 1) * 3;
    const int n1 = ns[0], n2 = ns[1], n3 = ns[2];
    B.n1 = n1; B.n2 = n2;
    B.nscan = c_P.nscan_tier == 0 ? 0 : c_P.nscan_tier == 1 ? n1 : c_P.nscan_tier == 2 ? n2 : n3;
  }
  B.s0 = B.nscan > 0 ? B.srow[0] : 0; B.s1 = B.nscan > 1 ? B.srow[1] : 0;
  B.s2 = B.nscan > 2 ? B.srow[2] : 0; B.s3 = B.nscan > 3 ? B.srow[3] : 0;

  const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
  const int cx = __umulhi((unsigned)pi.x, (unsigned)ncx);
  const int cy = __umulhi((unsigned)pi.y, (unsigned)ncy);
  const int cz = __umulhi((unsigned)pi.z, (unsigned)ncz);
  const int lx = local_layer(d, cx);
  // per column (lx', cy') the three z-cells are one contiguous range of the local order; a column that wraps in z
  // gets its far cell as a second, single-cell range (pass 1)
  const int zlo = d.cell_abs[2] ? 0 : max(cz - 1, 0), zhi = d.cell_abs[2] ? ncz - 1 : min(cz + 1, ncz - 1);
  const int zwrap = d.cell_abs[2] ? -1 : (cz == 0 ? ncz - 1 : (cz == ncz - 1 ? 0 : -1));
  const float screen = c_P.cutneighmaxsq_f;
  const float fsx = c_P.fscale[0], fsy = c_P.fscale[1], fsz = c_P.fscale[2];
  int nq = 0;
  // the partners' slots: the map entries were written by k_gather / k_ghost_map before this kernel
  int bslot[4] = {0, 0, 0, 0};
  if (d.bpa == 4) {
#pragma unroll
    for (int m = 0; m < 4; m++) bslot[m] = (m < nb) ? __ldg(&d.map[bpart[m] - 1]) : 0;
  }

  // ---- phase 1: fp32 screen of every candidate, four independent loads at a time ----
  for (int pass = 0; pass < (zwrap >= 0 ? 2 : 1); pass++) {
    const int za = pass ? zwrap : zlo, zb = pass ? zwrap : zhi;
    for (int ox = 0; ox < d.cell_span[0]; ox++) {
      int xc = d.cell_abs[0] ? ox : lx - 1 + ox;
      if (d.nranks == 1) { if (xc < 0) xc += ncx; else if (xc >= ncx) xc -= ncx; }   // one GPU: the slab is the whole box
      for (int oy = 0; oy < d.cell_span[1]; oy++) {
        int yc = d.cell_abs[1] ? oy : cy - 1 + oy;
        if (yc < 0) yc += ncy; else if (yc >= ncy) yc -= ncy;
        const int base = cell_slot(d, xc, yc, 0);
        const int lo = __ldg(&d.cell_start[base + za]), hi = __ldg(&d.cell_start[base + zb + 1]);
        for (int j = lo; j < hi; j += 4) {
          int4 p[4];
#pragma unroll
          for (int u = 0; u < 4; u++) p[u] = __ldg(&ph[min(j + u, hi - 1)]);
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int idx = (int)((unsigned)p[u].x - (unsigned)pi.x);
            const int idy = (int)((unsigned)p[u].y - (unsigned)pi.y);
            const int idz = (int)((unsigned)p[u].z - (unsigned)pi.z);
            const float fx = (float)idx * fsx, fy = (float)idy * fsy, fz = (float)idz * fsz;
            const float rsqf = fx * fx + fy * fy + fz * fz;
            if (rsqf <= screen && j + u < hi && j + u != i) {
              if (nq < BUILD_QUEUE) s_q[nq][t] = j + u;
              else build_accept(d, B, pi, p[u], j + u, rsqf);
              nq++;
            }
          }
        }
      }
    }
  }
  // ---- phase 2: decide the queued candidates ----
  const int nqq = min(nq, BUILD_QUEUE);
  for (int q = 0; q < nqq; q++) {
    const int jq = s_q[q][t];
    const int4 pj = __ldg(&ph[jq]);
    const int idx = (int)((unsigned)pj.x - (unsigned)pi.x);
    const int idy = (int)((unsigned)pj.y - (unsigned)pi.y);
    const int idz = (int)((unsigned)pj.z - (unsigned)pi.z);
    const float fx = (float)idx * fsx, fy = (float)idy * fsy, fz = (float)idz * fsz;
    build_accept(d, B, pi, pj, jq, fx * fx + fy * fy + fz * fz);
  }

  // bond partner rows
  bool missing = false;
  if (d.bpa == 4) {
#pragma unroll
    for (int m = 0; m < 4; m++)
      if (m < nb) {
        if (bslot[m] < 0) missing = true;
        else d.bondrow[(size_t)m * cap + i] = (unsigned)bslot[m] | ((unsigned)(btyp[m] - 1) << 28);
      }
  } else {
    for (int m = 0; m < nb; m++) {
      const int pt = d.bond_atom[(size_t)(tagi - 1) * d.bpa + m];
      const int bt = d.bond_type[(size_t)(tagi - 1) * d.bpa + m];
      const int jb = d.map[pt - 1];
      if (jb < 0) { missing = true; continue; }
      d.bondrow[(size_t)m * cap + i] = (unsigned)jb | ((unsigned)(bt - 1) << 28);
    }
  }
  if (missing) le_raise(d.ctrl, LE_DERR_MISSING_ATOM, tagi, nb);
  d.counts[i] = (unsigned)B.n | ((unsigned)nb << 16);
}

// list statistics on demand
__global__ void k_count_pairs(Dev d, unsigned long long *out) {
  unsigned long long h = 0, f = 0;
  const int lo = d.own0, hi = d.own0 + d.ctrl->nown;
  for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
    const unsigned c = d.counts[i];
    f += c & 0xff;
  }
  for (int o = 16; o > 0; o >>= 1) {
    h += __shfl_xor_sync(0xffffffffu, h, o);
    f += __shfl_xor_sync(0xffffffffu, f, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], h); atomicAdd(&out[1], f); }
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
