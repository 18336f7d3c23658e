// le_step2.cuh -- k_step2 / k_step2p: second generation of the fused step kernel (plain steps: no energy / virial tally).
//
// Same physics and -- operation for operation, rounding for rounding -- the same arithmetic as k_step<0,*,*,1>
// (le_md.cuh), so trajectories are bit-identical to it (tests/test_gpu_step2.py, scripts/step_ab.py check exactly
// that); what changes is how the work is issued.  The round-1 ncu capture of k_step
// (profiles/r01_kstep_sass_hotspots.txt) showed three things this kernel removes:
//   * every warp began by waiting for Ctrl::cur (a dependent L2 round trip, 9 % of the stall samples) before it
//     could address pos[cur]: the host knows the buffer parity of every launch -- also inside the captured graphs,
//     which are instantiated once per starting parity -- and passes it as a kernel argument (StepArgs::rdp1);
//     all twelve loads of the first batch are now independent of any earlier load;
//   * neighbor rows beyond the first four were walked two at a time through a row -> position dependency
//     (10 % of the stall samples for 10 % of the atoms): rows 4..7 are now requested together with the gathers of
//     the first batch and their positions come back as one batch of four;
//   * 842 warp instructions per 32 atoms, a third of them bookkeeping: branchy screens, development switches,
//     per-type coefficient look-ups of the uniform case, 64-bit image-flag comparisons, spills.  Here: branch-free
//     screens, one survivor queue per row group, carry-based image flags, wide multiplies in Philox, no switches:
//     758 instructions.
// The persistent form k_step2p (one wave of blocks, grid stride) is the default: 44.5 us against k_step's 54.5 us at
// 10^6 beads.  Written after the last GPU minute of round 1 and still off: the fused reneighbor decision (FUSE), fp32
// pair terms (P32), dynamic tile fetch (k_step2d).  Measured and removed again (profiles/r01_step_variants.txt; git history has the code): L2 / L1 prefetch
// of a thread's next atom, software pipelining with the next atom's head in registers, the thermostat force computed
// under the gathers + two FENE bonds evaluated side by side, int -> double conversion on the fp64 pipe, a 32-byte
// per-atom head record instead of eight arrays -- none of them faster than this form.
// Specialisation (the host falls back to k_step otherwise): one lj/cut coefficient set for all type pairs, special
// weights in {0, 1} only (every listed pair has factor 1).
//   reference: PairLJCut::compute src/pair_lj_cut.cpp:68-140, BondFENE::compute src/MOLECULE/bond_fene.cpp:52-128,
//   BondHarmonic::compute bond_harmonic.cpp:48-100, FixLangevin::post_force_templated src/fix_langevin.cpp:587-777,
//   FixNVE::initial/final_integrate src/fix_nve.cpp:64-140, Neighbor::check_distance src/neighbor.cpp:1962-2014.
#pragma once
#include "le_md.cuh"

__device__ __forceinline__ double le_rcp2(double x) {   // == le_rcp, with the contractions written out
  double t;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(x));
  t = __dmul_rn(t, __fma_rn(-x, t, 2.0));
  t = __dmul_rn(t, __fma_rn(-x, t, 2.0));
  return t;
}

// fp32 screen on the exact fixed-point differences (pair_screen of le_md.cuh without the branch)
__device__ __forceinline__ bool screen2(const int4 pi, const int4 pj, float sx, float sy, float sz, float cs) {
  const float dxf = (float)(int)((unsigned)pi.x - (unsigned)pj.x) * sx;
  const float dyf = (float)(int)((unsigned)pi.y - (unsigned)pj.y) * sy;
  const float dzf = (float)(int)((unsigned)pi.z - (unsigned)pj.z) * sz;
  return __fmaf_rn(dzf, dzf, __fmaf_rn(dxf, dxf, __fmul_rn(dyf, dyf))) < cs;
}

// WCA term of one listed pair in fp64 (pair_term<0,1> with factor_lj == 1)
__device__ __forceinline__ void pair_eval2(double &fx, double &fy, double &fz, const int4 pi, const int4 pj) {
  const double dy = __dmul_rn((double)((int)((unsigned)pi.y - (unsigned)pj.y)), c_P.scale[1]);
  const double dx = __dmul_rn((double)((int)((unsigned)pi.x - (unsigned)pj.x)), c_P.scale[0]);
  const double dz = __dmul_rn((double)((int)((unsigned)pi.z - (unsigned)pj.z)), c_P.scale[2]);
  const double rsq = __fma_rn(dz, dz, __fma_rn(dx, dx, __dmul_rn(dy, dy)));
  if (rsq < c_P.cutsq_d[0]) {
    const double r2inv = le_rcp2(rsq);
    const double r6inv = __dmul_rn(r2inv, __dmul_rn(r2inv, r2inv));
    const double fpair = __dmul_rn(r2inv, __dmul_rn(r6inv, __fma_rn(r6inv, c_P.lj1_d[0], -c_P.lj2_d[0])));
    fx = __fma_rn(dx, fpair, fx); fy = __fma_rn(dy, fpair, fy); fz = __fma_rn(dz, fpair, fz);
  }
}

// one listed pair in fp32, added to the atom's fp32 pair sum (pair_term32<0,1> of le_md.cuh without the branch: a row
// that is empty or outside the cutoff adds an exact zero, so the sums carry the same bits)
__device__ __forceinline__ void pair_acc32(float &px, float &py, float &pz, const int4 pi, const int4 pj, bool live, float sx, float sy, float sz) {
  const float dxf = __fmul_rn((float)(int)((unsigned)pi.x - (unsigned)pj.x), sx);
  const float dyf = __fmul_rn((float)(int)((unsigned)pi.y - (unsigned)pj.y), sy);
  const float dzf = __fmul_rn((float)(int)((unsigned)pi.z - (unsigned)pj.z), sz);
  const float rsqf = __fmaf_rn(dzf, dzf, __fmaf_rn(dxf, dxf, __fmul_rn(dyf, dyf)));
  const bool in = live && rsqf < c_P.cutsq[0];
  const float r2inv = __frcp_rn(in ? rsqf : 1.0f);
  const float r6inv = __fmul_rn(__fmul_rn(r2inv, r2inv), r2inv);
  const float fpair = in ? __fmul_rn(__fmul_rn(r6inv, __fmaf_rn(c_P.lj1[0], r6inv, -c_P.lj2[0])), r2inv) : 0.f;
  px = __fmaf_rn(dxf, fpair, px); py = __fmaf_rn(dyf, fpair, py); pz = __fmaf_rn(dzf, fpair, pz);
}

// FENE / harmonic term of one bond partner (bond_term<0>)
__device__ __forceinline__ void bond_eval2(double &fx, double &fy, double &fz, Ctrl *ctrl, const int4 pi, const int4 pj,
                                           unsigned e, int tagi) {
  const int bt = e >> 28;
  const double dy = __dmul_rn((double)((int)((unsigned)pi.y - (unsigned)pj.y)), c_P.scale[1]);
  const double dx = __dmul_rn((double)((int)((unsigned)pi.x - (unsigned)pj.x)), c_P.scale[0]);
  const double dz = __dmul_rn((double)((int)((unsigned)pi.z - (unsigned)pj.z)), c_P.scale[2]);
  const double rsq = __fma_rn(dz, dz, __fma_rn(dx, dx, __dmul_rn(dy, dy)));
  const int style = c_P.bstyle[bt];
  double fbond;
  if (style == 1) {  // FENE (bond_fene.cpp:79-117)
    double rlogarg = __fma_rn(-rsq, c_P.binvr0sq_d[bt], 1.0);
    if (rlogarg < 0.1) {
      if (rlogarg <= -3.0) le_raise(ctrl, LE_DERR_BAD_FENE, tagi, (int)(e & BOND_IDX_MASK));
      rlogarg = 0.1;
    }
    const double t = le_rcp2(__dmul_rn(rlogarg, rsq));          // one reciprocal serves 1/rlogarg and 1/rsq
    const double inv_rl = __dmul_rn(t, rsq), inv_rsq = __dmul_rn(t, rlogarg);
    fbond = __dmul_rn(-c_P.bk_d[bt], inv_rl);
    if (rsq < c_P.bcore_d[bt]) {
      const double sr2 = __dmul_rn(c_P.bsig2_d[bt], inv_rsq);
      const double sr6 = __dmul_rn(sr2, __dmul_rn(sr2, sr2));
      fbond = __fma_rn(__dmul_rn(__dmul_rn(c_P.beps48_d[bt], sr6), __dadd_rn(sr6, -0.5)), inv_rsq, fbond);
    }
  } else if (style == 2) {  // harmonic (bond_harmonic.cpp:71-80)
    double eb;
    fbond = harmonic_fbond(rsq, c_P.bk_d[bt], c_P.br0_d[bt], &eb);
  } else {
    fbond = 0.0;
  }
  fx = __fma_rn(dx, fbond, fx); fy = __fma_rn(dy, fbond, fy); fz = __fma_rn(dz, fbond, fz);
}


// same generator as philox4x32_7 (le_common.cuh) with each 32x32 -> 64-bit product taken as one wide multiply
__device__ __forceinline__ void philox4x32_7w(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]) {
#pragma unroll
  for (int r = 0; r < 7; r++) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (unsigned)p1; c2 = n2; c3 = (unsigned)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Langevin drag + uniform noise of one atom (FixLangevin::post_force_templated, src/fix_langevin.cpp:587-777)
__device__ __forceinline__ void step2_langevin(float &lx, float &ly, float &lz, const Ctrl *ctrl, const float4 vi, int tag, int ti, long long step) {
  unsigned r[4];
  philox4x32_7w((unsigned)tag, (unsigned)(step & 0xffffffffll), (unsigned)((unsigned long long)step >> 32), 0x4c45u,
                 c_P.seed_lo, c_P.seed_hi, r);
  float tsq = c_P.tsqrt_const;
  if (c_P.t_start != c_P.t_stop) {   // FixLangevin::compute_target (src/fix_langevin.cpp:784-820)
    float delta = (float)(step - ctrl->run_begin);
    if (delta != 0.0f) delta /= (float)(ctrl->run_end - ctrl->run_begin);
    tsq = sqrtf(__fmaf_rn(delta, __fadd_rn(c_P.t_stop, -c_P.t_start), c_P.t_start));
  }
  const float g1 = c_P.gfac1[ti], g2 = __fmul_rn(c_P.gfac2[ti], tsq);
  const float u0 = __fmaf_rn((float)(r[0] >> 8), 5.9604644775390625e-8f, -0.5f);
  const float u1 = __fmaf_rn((float)(r[1] >> 8), 5.9604644775390625e-8f, -0.5f);
  const float u2 = __fmaf_rn((float)(r[2] >> 8), 5.9604644775390625e-8f, -0.5f);
  lx = __fmaf_rn(g1, vi.x, __fmul_rn(g2, u0));
  ly = __fmaf_rn(g1, vi.y, __fmul_rn(g2, u1));
  lz = __fmaf_rn(g1, vi.z, __fmul_rn(g2, u2));
}

// the loads of an atom that the gathers depend on: its own position, the list counts, the first four neighbor rows and
// the first three bond rows.  No load depends on another.
struct Step2Head { int4 pi; unsigned cnt, en0, en1, en2, en3, eb0, eb1, eb2; };

__device__ __forceinline__ Step2Head step2_head(const Dev &d, const int4 *__restrict__ posr, const int i) {
  const int cap = d.cap;
  const unsigned *__restrict__ neigh = d.neigh;
  const unsigned *__restrict__ bondrow = d.bondrow;
  Step2Head h;
  h.pi = posr[i];
  h.cnt = d.counts[i];
  h.en0 = __ldg(&neigh[i]); h.en1 = __ldg(&neigh[(size_t)cap + i]);
  h.en2 = __ldg(&neigh[(size_t)2 * cap + i]); h.en3 = __ldg(&neigh[(size_t)3 * cap + i]);
  h.eb0 = __ldg(&bondrow[i]);                 // bondrow holds d.bpa rows
  h.eb1 = d.bpa > 1 ? __ldg(&bondrow[(size_t)cap + i]) : 0u;
  h.eb2 = d.bpa > 2 ? __ldg(&bondrow[(size_t)2 * cap + i]) : 0u;
  return h;
}

// one atom (slot i) of one timestep, its head already requested.  P32: pair terms in fp32 (Params::pair32)
template <int DD, int P32 = 0>
__device__ __forceinline__ void step2_atom(const Dev &d, const StepArgs &a, const int i, const int rd, const Step2Head &h) {
  const int cap = d.cap;
  Ctrl *__restrict__ ctrl = d.ctrl;
  const unsigned *__restrict__ neigh = d.neigh;
  const unsigned *__restrict__ bondrow = d.bondrow;
  const int4 *__restrict__ posr = d.pos[rd];
  int4 *__restrict__ posw = d.pos[rd ^ 1];

  // ---- batch 1 (with the head): everything addressed by i ----
  const int4 pi = h.pi;
  const unsigned cnt = h.cnt, en0 = h.en0, en1 = h.en1, en2 = h.en2, en3 = h.en3, eb0 = h.eb0, eb1 = h.eb1, eb2 = h.eb2;
  float4 vi = d.vel[i];
  const int4 ph = d.pos_hold[i];
  const long long step = ctrl->step;
  const int nn = cnt & 0xff, nb = (cnt >> 16) & 0xff;
  const int ti = pi.w & 7;
  const int tag = pi.w >> 3;
  // ---- batch 2: the neighbor gathers (a slot beyond the count gathers the atom itself), and neighbor rows 4..7 ----
  const int4 pn0 = __ldg(&posr[0 < nn ? (int)(en0 & NEIGH_IDX_MASK) : i]);
  const int4 pn1 = __ldg(&posr[1 < nn ? (int)(en1 & NEIGH_IDX_MASK) : i]);
  const int4 pn2 = __ldg(&posr[2 < nn ? (int)(en2 & NEIGH_IDX_MASK) : i]);
  const int4 pn3 = __ldg(&posr[3 < nn ? (int)(en3 & NEIGH_IDX_MASK) : i]);
  unsigned et0 = 0, et1 = 0, et2 = 0, et3 = 0;
  if (nn > 4) {
    const unsigned *__restrict__ r4 = neigh + (size_t)4 * cap + i;
    et0 = __ldg(r4);
    et1 = nn > 5 ? __ldg(r4 + cap) : 0u;
    et2 = nn > 6 ? __ldg(r4 + 2 * (size_t)cap) : 0u;
    et3 = nn > 7 ? __ldg(r4 + 3 * (size_t)cap) : 0u;
  }

  const float sx = c_P.fscale[0], sy = c_P.fscale[1], sz = c_P.fscale[2];
  int4 pb0, pb1, pb2;      // the bond partners' positions
  double fx = 0.0, fy = 0.0, fz = 0.0;
  if (P32) {
    // pair terms in fp32, rows in ascending order, no branch for the first four; the bond partners' positions are
    // requested in between and arrive while the tail rows are worked on
    float px = 0.f, py = 0.f, pz = 0.f;
    pair_acc32(px, py, pz, pi, pn0, 0 < nn, sx, sy, sz);
    pair_acc32(px, py, pz, pi, pn1, 1 < nn, sx, sy, sz);
    pair_acc32(px, py, pz, pi, pn2, 2 < nn, sx, sy, sz);
    pair_acc32(px, py, pz, pi, pn3, 3 < nn, sx, sy, sz);
    pb0 = __ldg(&posr[0 < nb ? (int)(eb0 & BOND_IDX_MASK) : i]);
    pb1 = __ldg(&posr[1 < nb ? (int)(eb1 & BOND_IDX_MASK) : i]);
    pb2 = __ldg(&posr[2 < nb ? (int)(eb2 & BOND_IDX_MASK) : i]);
#pragma unroll 1
    for (int kb = 4; kb < nn; kb += 4) {
      if (kb > 4) {
        et0 = __ldg(&neigh[(size_t)kb * cap + i]);
        et1 = kb + 1 < nn ? __ldg(&neigh[(size_t)(kb + 1) * cap + i]) : 0u;
        et2 = kb + 2 < nn ? __ldg(&neigh[(size_t)(kb + 2) * cap + i]) : 0u;
        et3 = kb + 3 < nn ? __ldg(&neigh[(size_t)(kb + 3) * cap + i]) : 0u;
      }
      const int4 q0 = __ldg(&posr[(int)(et0 & NEIGH_IDX_MASK)]);
      const int4 q1 = __ldg(&posr[kb + 1 < nn ? (int)(et1 & NEIGH_IDX_MASK) : i]);
      const int4 q2 = __ldg(&posr[kb + 2 < nn ? (int)(et2 & NEIGH_IDX_MASK) : i]);
      const int4 q3 = __ldg(&posr[kb + 3 < nn ? (int)(et3 & NEIGH_IDX_MASK) : i]);
      pair_acc32(px, py, pz, pi, q0, true, sx, sy, sz);
      pair_acc32(px, py, pz, pi, q1, kb + 1 < nn, sx, sy, sz);
      pair_acc32(px, py, pz, pi, q2, kb + 2 < nn, sx, sy, sz);
      pair_acc32(px, py, pz, pi, q3, kb + 3 < nn, sx, sy, sz);
    }
    fx = (double)px; fy = (double)py; fz = (double)pz;
  } else {
    const float cs = c_P.cutsq_screen[0];
    // screen every listed pair in fp32; bit k of `hit` = row k is (a hair more than) inside the force cutoff
    // (the first four without a branch: an empty slot holds the atom itself, passes, and is masked off by the count)
    unsigned hit = (screen2(pi, pn0, sx, sy, sz, cs) ? 1u : 0u) | (screen2(pi, pn1, sx, sy, sz, cs) ? 2u : 0u) |
                   (screen2(pi, pn2, sx, sy, sz, cs) ? 4u : 0u) | (screen2(pi, pn3, sx, sy, sz, cs) ? 8u : 0u);
    hit &= (1u << min(nn, 4)) - 1u;
    // the bond partners' positions: requested now (the registers of the four screened positions are free again),
    // they arrive while the survivors are evaluated
    pb0 = __ldg(&posr[0 < nb ? (int)(eb0 & BOND_IDX_MASK) : i]);
    pb1 = __ldg(&posr[1 < nb ? (int)(eb1 & BOND_IDX_MASK) : i]);
    pb2 = __ldg(&posr[2 < nb ? (int)(eb2 & BOND_IDX_MASK) : i]);
    // rows 4.. four at a time (the first group was requested with batch 2)
#pragma unroll 1
    for (int kb = 4; kb < nn && kb < 32; kb += 4) {
      if (kb > 4) {
        et0 = __ldg(&neigh[(size_t)kb * cap + i]);
        et1 = kb + 1 < nn ? __ldg(&neigh[(size_t)(kb + 1) * cap + i]) : 0u;
        et2 = kb + 2 < nn ? __ldg(&neigh[(size_t)(kb + 2) * cap + i]) : 0u;
        et3 = kb + 3 < nn ? __ldg(&neigh[(size_t)(kb + 3) * cap + i]) : 0u;
      }
      const int4 q0 = __ldg(&posr[(int)(et0 & NEIGH_IDX_MASK)]);
      const int4 q1 = __ldg(&posr[kb + 1 < nn ? (int)(et1 & NEIGH_IDX_MASK) : i]);
      const int4 q2 = __ldg(&posr[kb + 2 < nn ? (int)(et2 & NEIGH_IDX_MASK) : i]);
      const int4 q3 = __ldg(&posr[kb + 3 < nn ? (int)(et3 & NEIGH_IDX_MASK) : i]);
      unsigned h = 0;
      if (screen2(pi, q0, sx, sy, sz, cs)) h |= 1u;
      if (kb + 1 < nn && screen2(pi, q1, sx, sy, sz, cs)) h |= 2u;
      if (kb + 2 < nn && screen2(pi, q2, sx, sy, sz, cs)) h |= 4u;
      if (kb + 3 < nn && screen2(pi, q3, sx, sy, sz, cs)) h |= 8u;
      hit |= h << kb;
    }

    // rows 32.. (dense systems only): no queue bit left, evaluate directly
#pragma unroll 1
    for (int k = 32; k < nn; k++) {
      const unsigned e = __ldg(&neigh[(size_t)k * cap + i]);
      pair_eval2(fx, fy, fz, pi, __ldg(&posr[e & NEIGH_IDX_MASK]));
    }
    // the survivors, evaluated in fp64 in the order k_step adds them: rows 4, 5, ... first, then 0..3; a warp runs each
    // loop max-over-lanes(#survivors) times.  Rows 4.. are rare (their row entry is fetched again: an L1 hit)
#pragma unroll 1
    for (unsigned m = hit >> 4; m; m &= m - 1) {
      const unsigned e = __ldg(&neigh[(size_t)(__ffs(m) + 3) * cap + i]);
      pair_eval2(fx, fy, fz, pi, __ldg(&posr[e & NEIGH_IDX_MASK]));
    }
#pragma unroll 1
    for (unsigned m = hit & 15u; m; m &= m - 1) {
      const unsigned b = m & (0u - m);                         // lowest survivor: 1, 2, 4 or 8
      const unsigned e = (b & 3u) ? ((b & 1u) ? en0 : en1) : ((b & 4u) ? en2 : en3);
      pair_eval2(fx, fy, fz, pi, __ldg(&posr[e & NEIGH_IDX_MASK]));   // second touch of the position: an L1 hit
    }
  }
  if (0 < nb) bond_eval2(fx, fy, fz, ctrl, pi, pb0, eb0, tag);
  if (1 < nb) bond_eval2(fx, fy, fz, ctrl, pi, pb1, eb1, tag);
  if (2 < nb) bond_eval2(fx, fy, fz, ctrl, pi, pb2, eb2, tag);
#pragma unroll 1
  for (int mth = 3; mth < nb; mth++) {
    const unsigned e = __ldg(&bondrow[(size_t)mth * cap + i]);
    bond_eval2(fx, fy, fz, ctrl, pi, __ldg(&posr[e & BOND_IDX_MASK]), e, tag);
  }

  // ---- Langevin drag + uniform noise (post_force); fp32, added to the rounded conservative force ----
  float lx = 0.f, ly = 0.f, lz = 0.f;
  if (a.langevin) step2_langevin(lx, ly, lz, ctrl, vi, tag, ti, step);

  // ---- velocity Verlet ----
  const float dtfm = c_P.dtfm[ti];
  const float ffx = __fadd_rn(__double2float_rn(fx), lx), ffy = __fadd_rn(__double2float_rn(fy), ly), ffz = __fadd_rn(__double2float_rn(fz), lz);
  if (a.do_final) {
    vi.x = __fmaf_rn(dtfm, ffx, vi.x); vi.y = __fmaf_rn(dtfm, ffy, vi.y); vi.z = __fmaf_rn(dtfm, ffz, vi.z);
    if (c_P.vlimitsq > 0.0f) {   // FixNVELimit::final_integrate
      const float vsq = __fmaf_rn(vi.z, vi.z, __fmaf_rn(vi.x, vi.x, __fmul_rn(vi.y, vi.y)));
      if (vsq > c_P.vlimitsq) { const float sc = sqrtf(c_P.vlimitsq / vsq); vi.x *= sc; vi.y *= sc; vi.z *= sc; }
    }
  }
  if (a.do_initial) {
    vi.x = __fmaf_rn(dtfm, ffx, vi.x); vi.y = __fmaf_rn(dtfm, ffy, vi.y); vi.z = __fmaf_rn(dtfm, ffz, vi.z);
    if (c_P.vlimitsq > 0.0f) {   // FixNVELimit::initial_integrate
      const float vsq = __fmaf_rn(vi.z, vi.z, __fmaf_rn(vi.x, vi.x, __fmul_rn(vi.y, vi.y)));
      if (vsq > c_P.vlimitsq) { const float sc = sqrtf(c_P.vlimitsq / vsq); vi.x *= sc; vi.y *= sc; vi.z *= sc; }
    }
    const int dux = __float2int_rn(__fmul_rn(__fmul_rn(c_P.dt, vi.x), c_P.inv_fscale[0]));
    const int duy = __float2int_rn(__fmul_rn(__fmul_rn(c_P.dt, vi.y), c_P.inv_fscale[1]));
    const int duz = __float2int_rn(__fmul_rn(__fmul_rn(c_P.dt, vi.z), c_P.inv_fscale[2]));
    // image flags: a wrap of the 32-bit coordinate is a periodic crossing (Domain::remap); the carry of the
    // unsigned add plus the sign of the step is +1 / -1 / 0
    const unsigned long long ax = (unsigned long long)(unsigned)pi.x + (unsigned)dux;
    const unsigned long long ay = (unsigned long long)(unsigned)pi.y + (unsigned)duy;
    const unsigned long long az = (unsigned long long)(unsigned)pi.z + (unsigned)duz;
    const unsigned nx = (unsigned)ax, ny = (unsigned)ay, nz = (unsigned)az;
    const int wx = (int)(ax >> 32) + (dux >> 31), wy = (int)(ay >> 32) + (duy >> 31), wz = (int)(az >> 32) + (duz >> 31);
    if (wx | wy | wz) {
      const int im = d.img[i];
      const int ix = (im & 1023) - 512 + wx;
      const int iy = ((im >> 10) & 1023) - 512 + wy;
      const int iz = ((im >> 20) & 1023) - 512 + wz;  // 10+10+10 packing of LAMMPS_SMALLBIG (src/lmptype.h)
      d.img[i] = ((ix + 512) & 1023) | (((iy + 512) & 1023) << 10) | (((iz + 512) & 1023) << 20);
    }
    const int4 pnew = make_int4((int)nx, (int)ny, (int)nz, pi.w);
    posw[i] = pnew;
    if (DD) {   // halo update fused into the integrator (see k_step)
      if (i < ctrl->send_l_end) d.peer[left_rank(d)].pos[rd ^ 1][d.gr0 + (i - d.own0)] = pnew;
      const int srb = ctrl->send_r_beg;
      if (i >= srb) d.peer[right_rank(d)].pos[rd ^ 1][i - srb] = pnew;
    }
    // displacement since the last rebuild
    const float hx = __fmul_rn((float)(int)(nx - (unsigned)ph.x), sx);
    const float hy = __fmul_rn((float)(int)(ny - (unsigned)ph.y), sy);
    const float hz = __fmul_rn((float)(int)(nz - (unsigned)ph.z), sz);
    if (__fmaf_rn(hz, hz, __fmaf_rn(hx, hx, __fmul_rn(hy, hy))) > c_P.triggersq) ctrl->moved = 1;
  }
  d.vel[i] = vi;
}

// Work order on a slab of a multi-GPU run: the two boundary slices first, the interior last, so that the halo stores
// are in flight while the interior computes (see k_step).  Segment starts are multiples of 64 slots.  Maps position g
// of that order to a slot; returns own_end for a padding position.
struct Step2Order { int own0, own_end, a_end, b_beg, nl, nr, nrp; };
__device__ __forceinline__ Step2Order step2_order(const Dev &d) {
  const Ctrl *__restrict__ ctrl = d.ctrl;
  Step2Order o;
  o.own0 = d.own0;
  o.own_end = d.own0 + ctrl->nown;
  o.a_end = min(o.own_end, d.own0 + ((ctrl->send_l_end - d.own0 + 63) & ~63));
  o.b_beg = max(o.a_end, d.own0 + ((ctrl->send_r_beg - d.own0) & ~63));
  o.nl = o.a_end - d.own0; o.nr = o.own_end - o.b_beg; o.nrp = (o.nr + 63) & ~63;
  return o;
}
__device__ __forceinline__ int step2_slot(const Step2Order &o, int g) {
  if (g < o.nl) return o.own0 + g;
  if (g < o.nl + o.nrp) return (g - o.nl < o.nr) ? o.b_beg + (g - o.nl) : o.own_end;
  const int i = o.a_end + (g - o.nl - o.nrp);
  return i >= o.b_beg ? o.own_end : i;
}

// one block per NT atoms.  DD: multi-GPU slab (halo stores fused in, boundary slices first); NT: threads per block
// (1024 / NT blocks per SM)
template <int DD, int NT, int P32 = 0>
__global__ void __launch_bounds__(NT, 1024 / NT) k_step2(Dev d, StepArgs a) {
  int i = d.own0 + blockIdx.x * NT + threadIdx.x;
  if (DD) {
    const Step2Order o = step2_order(d);
    i = step2_slot(o, blockIdx.x * NT + threadIdx.x);
    if (i >= o.own_end) return;
  } else {
    if (i >= d.own0 + d.N) return;   // one GPU owns every atom: no look at the control block before the loads
  }
  const int rd = a.rdp1 ? a.rdp1 - 1 : d.ctrl->cur;
  step2_atom<DD, P32>(d, a, i, rd, step2_head(d, d.pos[rd], i));
}

// persistent form (the default): one wave of blocks walks the atoms with a grid stride -- no block launches inside
// the step, no partial last wave; on a slab the stride runs over the boundary-first order.  FUSE: with the epilogue
// that takes the reneighbor decision of the next timestep (StepArgs::fuse); P32: pair terms in fp32 (Params::pair32)
template <int DD, int NT, int FUSE = 0, int P32 = 0>
__global__ void __launch_bounds__(NT, 1024 / NT) k_step2p(Dev d, StepArgs a) {
  const int rd = a.rdp1 ? a.rdp1 - 1 : d.ctrl->cur;
  const int stride = gridDim.x * NT;
  if (DD) {
    const Step2Order o = step2_order(d);
    const int gend = o.nl + o.nrp + (o.b_beg - o.a_end);      // positions of the order: left slice, padded right slice, interior
#pragma unroll 1
    for (int g = blockIdx.x * NT + threadIdx.x; g < gend; g += stride) {
      const int i = step2_slot(o, g);
      if (i < o.own_end) step2_atom<DD, P32>(d, a, i, rd, step2_head(d, d.pos[rd], i));
    }
  } else {
    const int end = d.own0 + d.N;
#pragma unroll 1
    for (int i = d.own0 + blockIdx.x * NT + threadIdx.x; i < end; i += stride)
      step2_atom<DD, P32>(d, a, i, rd, step2_head(d, d.pos[rd], i));
    if (FUSE && a.fuse) {
      // the block that finishes last has seen the `moved` stores of all the others (fence + counter): it does k_decide's
      // work for the next timestep -- one kernel and one dependency bubble less per step inside the steady-state graph
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&d.ctrl->blocks_done, 1u) == gridDim.x - 1) {
          d.ctrl->blocks_done = 0;
          __threadfence();
          decide_body(d, (cudaGraphConditionalHandle)a.handle, 1, 1);
        }
      }
    }
  }
}

// persistent with dynamic tile fetch (one GPU): a block takes its first tile by its index and every further one from
// a counter, fetched one tile ahead so that the atomic's latency hides behind the work.  With the static grid stride
// 3907 tiles over 592 blocks leave a third of the blocks idle for the last seventh of the kernel (warps active 47 % of
// a possible 50 %).  Which block works on which tile does not change any result.
template <int NT, int P32 = 0>
__global__ void __launch_bounds__(NT, 1024 / NT) k_step2d(Dev d, StepArgs a) {
  __shared__ int s_next;
  const int rd = a.rdp1 ? a.rdp1 - 1 : d.ctrl->cur;
  const int end = d.own0 + d.N, ntiles = (d.N + NT - 1) / NT;
  int tile = blockIdx.x;
  while (tile < ntiles) {                       // uniform over the block
    if (threadIdx.x == 0) s_next = (int)atomicAdd(&d.ctrl->tile_next, 1u) + (int)gridDim.x;
    const int i = d.own0 + tile * NT + threadIdx.x;
    if (i < end) step2_atom<0, P32>(d, a, i, rd, step2_head(d, d.pos[rd], i));
    __syncthreads();
    tile = s_next;
    __syncthreads();                            // all have read s_next before thread 0 writes the next one
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&d.ctrl->blocks_done, 1u) == gridDim.x - 1) { d.ctrl->blocks_done = 0; d.ctrl->tile_next = 0; }
  }
}
