// le_step3.cuh -- k_step3: the fused timestep kernel (third generation), one warp per tile of 32 owned atoms.
//
//   reference: PairLJCut::compute src/pair_lj_cut.cpp:68-140, BondFENE::compute src/MOLECULE/bond_fene.cpp:52-128,
//   BondHarmonic::compute bond_harmonic.cpp:48-100, FixLangevin::post_force_templated src/fix_langevin.cpp:587-777,
//   FixNVE::initial/final_integrate src/fix_nve.cpp:64-140, Neighbor::check_distance src/neighbor.cpp:1962-2014.
//
// What the round-1/2 kernels (one thread per atom over ELL neighbor rows, k_step / k_step2p) cost at 10^6 beads: 758 warp
// instructions per 32 atoms, 30 % of them in loops that a warp ran for one to five of its lanes (rows beyond the first
// four, the fp64 survivors of the screen), 98 bytes per atom and step through DRAM for a working set (96 MB) that just
// misses the 126 MB L2.  This kernel changes the division of labour:
//   * PAIRS are evaluated pair-parallel: the tile's list is one flat run of (owner lane, neighbor slot) entries
//     (k_build3), lane l takes entries l, l+32, ...: every lane busy whatever the spread of neighbor counts, the run is
//     read as whole 128-byte lines (4 n_full bytes per atom instead of four ELL rows), the neighbor gathers of a round
//     are 32 independent loads.  The WCA term is fp64 on exact fixed-point differences (see pair3); the lane parks its
//     term in shared memory and the OWNER adds its terms in list order -- a fixed order, so results are reproducible
//     run to run;
//   * BONDS stay lane-per-owner (two to three per bead, no spread) and fp64: FENE and its WCA core cancel to a tenth
//     of their size;
//   * the displacement test of Neighbor::check_distance reads pos_hold only for atoms whose accumulated path length
//     (an upper bound of the displacement, carried in the aux word of the velocity) has reached skin/2; list counts
//     live in the same word: per step an atom costs pos 16 r + 16 w, vel 16 r + 16 w, 3 bond rows, its list entries.
// Energy / virial tally (EV) and force output use the SAME arithmetic (one code path), so the per-atom force parity
// tests exercise exactly what the production steps run.
#pragma once
#include "le_md.cuh"

#define STEP3_THREADS 256
#define STEP3_WARPS (STEP3_THREADS / 32)
#define STEP3_CH 4                       // rounds of 32 entries per pass of the pair phase

__device__ __forceinline__ double le_rcp2(double x) {   // 1/x to full double accuracy: MUFU.RCP64H seed + two Newton steps
  double t;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(x));
  t = __dmul_rn(t, __fma_rn(-x, t, 2.0));
  t = __dmul_rn(t, __fma_rn(-x, t, 2.0));
  return t;
}

struct EvAcc { double evdwl, ebond, v[6], warn; };

// WCA term of one list entry, fp64 (PairLJCut::compute, pair_lj_cut.cpp:102-118): force ON THE OWNER (position po) from
// the neighbor pj, written to (tx, ty, tz).  A dead entry (beyond the run, or outside the cutoff) yields an exact zero.
// fp32 pair terms were measured (round 2): 2.6 % faster, but the per-atom force error reaches 1.3e-5 of the rms force
// in a dense melt (five r^-14 terms of 30..50 per atom against a net force of 15): the 1e-5 bar needs fp64 here, and
// on this kernel -- bound by instruction issue, fp64 pipe at 10 % -- an fp64 instruction costs the same issue slot.
template <int EV, int UNI>
__device__ __forceinline__ void pair3(double &tx, double &ty, double &tz, const int4 po, const int4 pj, unsigned e, bool live, int nt, EvAcc &A) {
  const double dy = __dmul_rn((double)((int)((unsigned)po.y - (unsigned)pj.y)), c_P.scale[1]);
  const double dx = __dmul_rn((double)((int)((unsigned)po.x - (unsigned)pj.x)), c_P.scale[0]);
  const double dz = __dmul_rn((double)((int)((unsigned)po.z - (unsigned)pj.z)), c_P.scale[2]);
  const double rsq = __fma_rn(dz, dz, __fma_rn(dx, dx, __dmul_rn(dy, dy)));
  const int tp = UNI ? 0 : (po.w & 7) * nt + (pj.w & 7);
  const bool in = live && rsq < c_P.cutsq_d[tp];
  const double r2inv = le_rcp2(in ? rsq : 1.0);
  const double r6inv = __dmul_rn(r2inv, __dmul_rn(r2inv, r2inv));
  double fpair = __dmul_rn(r2inv, __dmul_rn(r6inv, __fma_rn(r6inv, c_P.lj1_d[tp], -c_P.lj2_d[tp])));
  const double factor = UNI ? 1.0 : (double)c_P.special_lj[e >> 30];
  if (!UNI) fpair *= factor;
  if (!in) fpair = 0.0;
  if (EV && in) {
    A.evdwl += factor * (r6inv * (c_P.lj3_d[tp] * r6inv - c_P.lj4_d[tp]) - c_P.offset_d[tp]);
    A.v[0] += dx * dx * fpair; A.v[1] += dy * dy * fpair; A.v[2] += dz * dz * fpair;
    A.v[3] += dx * dy * fpair; A.v[4] += dx * dz * fpair; A.v[5] += dy * dz * fpair;
  }
  tx = __dmul_rn(dx, fpair); ty = __dmul_rn(dy, fpair); tz = __dmul_rn(dz, fpair);
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

// FENE / harmonic term of one bond partner, fp64 (bond_fene.cpp:79-117, bond_harmonic.cpp:71-80)
template <int EV>
__device__ __forceinline__ void bond3(double &fx, double &fy, double &fz, Ctrl *ctrl, const int4 pi, const int4 pj, unsigned e, int tagi, EvAcc &A) {
  const int bt = e >> 28;
  const double dy = __dmul_rn((double)((int)((unsigned)pi.y - (unsigned)pj.y)), c_P.scale[1]);
  const double dx = __dmul_rn((double)((int)((unsigned)pi.x - (unsigned)pj.x)), c_P.scale[0]);
  const double dz = __dmul_rn((double)((int)((unsigned)pi.z - (unsigned)pj.z)), c_P.scale[2]);
  const double rsq = __fma_rn(dz, dz, __fma_rn(dx, dx, __dmul_rn(dy, dy)));
  const int style = c_P.bstyle[bt];
  double fbond;
  if (style == 1) {
    double rlogarg = __fma_rn(-rsq, c_P.binvr0sq_d[bt], 1.0);
    if (rlogarg < 0.1) {
      if (EV) A.warn += 0.5;                                  // each long bond is seen from both ends
      if (rlogarg <= -3.0) le_raise(ctrl, LE_DERR_BAD_FENE, tagi, (int)(e & BOND_IDX_MASK));
      rlogarg = 0.1;
    }
    const double t = le_rcp2(__dmul_rn(rlogarg, rsq));          // one reciprocal serves 1/rlogarg and 1/rsq
    const double inv_rl = __dmul_rn(t, rsq), inv_rsq = __dmul_rn(t, rlogarg);
    fbond = __dmul_rn(-c_P.bk_d[bt], inv_rl);
    double sr6 = 0.0;
    const bool core = rsq < c_P.bcore_d[bt];
    if (core) {
      const double sr2 = __dmul_rn(c_P.bsig2_d[bt], inv_rsq);
      sr6 = __dmul_rn(sr2, __dmul_rn(sr2, sr2));
      fbond = __fma_rn(__dmul_rn(__dmul_rn(c_P.beps48_d[bt], sr6), __dadd_rn(sr6, -0.5)), inv_rsq, fbond);
    }
    if (EV) {
      double eb = -0.5 * c_P.bk_d[bt] * c_P.br0sq_d[bt] * log(rlogarg);
      if (core) eb += 4.0 * c_P.beps_d[bt] * sr6 * (sr6 - 1.0) + c_P.beps_d[bt];
      A.ebond += eb;
    }
  } else if (style == 2) {
    double eb;
    fbond = harmonic_fbond(rsq, c_P.bk_d[bt], c_P.br0_d[bt], &eb);
    if (EV) A.ebond += eb;
  } else {
    fbond = 0.0;
  }
  fx = __fma_rn(dx, fbond, fx); fy = __fma_rn(dy, fbond, fy); fz = __fma_rn(dz, fbond, fz);
  if (EV) {
    A.v[0] += dx * dx * fbond; A.v[1] += dy * dy * fbond; A.v[2] += dz * dz * fbond;
    A.v[3] += dx * dy * fbond; A.v[4] += dx * dz * fbond; A.v[5] += dy * dz * fbond;
  }
}

// Philox4x32-7 with each 32x32 -> 64-bit product taken as one wide multiply
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
__device__ __forceinline__ void step3_langevin(float &lx, float &ly, float &lz, const Ctrl *ctrl, const float4 vi, int tag, int ti, long long step) {
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

// Work order on a slab of a multi-GPU run: the tiles of the two boundary slices first, the interior last, so that the
// halo stores are in flight while the interior computes.  Slice ends are multiples of 64 slots (own0 is one).
struct Step3Order { int a_end_t, b_beg_t, nl_t, nr_t, total; };
__device__ __forceinline__ Step3Order step3_order(const Dev &d, int own_end) {
  const Ctrl *__restrict__ ctrl = d.ctrl;
  const int a_end = min(own_end, d.own0 + ((ctrl->send_l_end - d.own0 + 63) & ~63));
  const int b_beg = max(a_end, d.own0 + ((ctrl->send_r_beg - d.own0) & ~63));
  Step3Order o;
  o.a_end_t = (a_end - d.own0 + TILE - 1) >> 5;
  o.b_beg_t = (b_beg - d.own0) >> 5;
  o.nl_t = o.a_end_t;
  o.nr_t = (own_end - b_beg + TILE - 1) >> 5;
  if (o.b_beg_t < o.a_end_t) o.b_beg_t = o.a_end_t;            // (only when the slab is a single slice)
  o.total = o.nl_t + o.nr_t + (o.b_beg_t - o.a_end_t);
  return o;
}
__device__ __forceinline__ int step3_tile_of(const Step3Order &o, int g) {
  if (g < o.nl_t) return g;
  if (g < o.nl_t + o.nr_t) return o.b_beg_t + (g - o.nl_t);
  return o.a_end_t + (g - o.nl_t - o.nr_t);
}

template <int EV, int DD, int UNI>
__global__ void __launch_bounds__(STEP3_THREADS, EV ? 2 : 4) k_step3(Dev d, StepArgs a) {
  __shared__ int4 s_pos[STEP3_WARPS][TILE];
  __shared__ double s_fx[STEP3_WARPS][STEP3_CH * TILE], s_fy[STEP3_WARPS][STEP3_CH * TILE], s_fz[STEP3_WARPS][STEP3_CH * TILE];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cap = d.cap;
  Ctrl *__restrict__ ctrl = d.ctrl;
  const int rd = a.rdp1 ? a.rdp1 - 1 : ctrl->cur;
  const int own_end = d.own0 + (DD ? ctrl->nown : d.N);        // one GPU owns every atom: no look at the control block
  const int4 *__restrict__ posr = d.pos[rd];
  int4 *__restrict__ posw = d.pos[rd ^ 1];
  const unsigned *__restrict__ bondrow = d.bondrow;
  const float sx = c_P.fscale[0], sy = c_P.fscale[1], sz = c_P.fscale[2];
  const int nt = c_P.ntypes;
  const int tcap = d.tcap;
  const long long step = ctrl->step;
  Step3Order ord;
  int ntiles = (own_end - d.own0 + TILE - 1) >> 5;
  if (DD) { ord = step3_order(d, own_end); ntiles = ord.total; }

  EvAcc A;
  double ke = 0.0;
  if (EV) { A.evdwl = A.ebond = A.warn = 0.0; for (int q = 0; q < 6; q++) A.v[q] = 0.0; }

#pragma unroll 1
  for (int g = blockIdx.x * STEP3_WARPS + wib; g < ntiles; g += gridDim.x * STEP3_WARPS) {
    const int tile = DD ? step3_tile_of(ord, g) : g;
    const int i0 = d.own0 + tile * TILE;
    const bool valid = i0 + lane < own_end;
    const int i = valid ? i0 + lane : own_end - 1;            // lanes beyond the end shadow the last atom (loads only)
    // ---- level 1: everything addressed by the tile / the atom ----
    const int4 pi = posr[i];
    float4 vi = d.vel[i];
    const unsigned cnt = __ldg(&d.tile_cnt[tile]);
    const unsigned *__restrict__ run = d.nbr + (size_t)tile * tcap;
    const unsigned e0 = __ldg(&run[lane]), e1 = __ldg(&run[TILE + lane]), e2 = __ldg(&run[2 * TILE + lane]);   // tcap >= 128
    const unsigned eb0 = __ldg(&bondrow[i]);
    const unsigned eb1 = d.bpa > 1 ? __ldg(&bondrow[(size_t)cap + i]) : 0u;
    const unsigned eb2 = d.bpa > 2 ? __ldg(&bondrow[(size_t)2 * cap + i]) : 0u;
    const unsigned aux = __float_as_uint(vi.w);
    const int nn = valid ? (int)AUX_NN(aux) : 0, nb = valid ? (int)AUX_NB(aux) : 0;
    const int ti = pi.w & 7, tag = pi.w >> 3;
    s_pos[wib][lane] = pi;
    // start of this lane's entries inside the run
    int inc = nn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    const int off = inc - nn;
    __syncwarp();
    // ---- level 2: the gathers of the first two rounds and of the bond partners ----
    const bool l0 = (unsigned)lane < cnt, l1 = (unsigned)(TILE + lane) < cnt, l2 = (unsigned)(2 * TILE + lane) < cnt;
    const int4 pj0 = __ldg(&posr[l0 ? (int)(e0 & NEIGH_IDX_MASK) : i]);
    const int4 pj1 = __ldg(&posr[l1 ? (int)(e1 & NEIGH_IDX_MASK) : i]);
    const int4 pj2 = __ldg(&posr[l2 ? (int)(e2 & NEIGH_IDX_MASK) : i]);
    const int4 pb0 = __ldg(&posr[0 < nb ? (int)(eb0 & BOND_IDX_MASK) : i]);
    const int4 pb1 = __ldg(&posr[1 < nb ? (int)(eb1 & BOND_IDX_MASK) : i]);
    const int4 pb2 = __ldg(&posr[2 < nb ? (int)(eb2 & BOND_IDX_MASK) : i]);

    // ---- pairs: STEP3_CH rounds per pass; the owner adds its terms in list order after each pass ----
    double fx = 0.0, fy = 0.0, fz = 0.0;
    double *__restrict__ sfx = s_fx[wib], *__restrict__ sfy = s_fy[wib], *__restrict__ sfz = s_fz[wib];
    {
      pair3<EV, UNI>(sfx[lane], sfy[lane], sfz[lane], s_pos[wib][(e0 >> NEIGH_IDX_BITS) & 31], pj0, e0, l0, nt, A);
      pair3<EV, UNI>(sfx[TILE + lane], sfy[TILE + lane], sfz[TILE + lane], s_pos[wib][(e1 >> NEIGH_IDX_BITS) & 31], pj1, e1, l1, nt, A);
      if (cnt > 2 * TILE)                                     // warp-uniform (the third round's loads were issued with the first two)
        pair3<EV, UNI>(sfx[2 * TILE + lane], sfy[2 * TILE + lane], sfz[2 * TILE + lane], s_pos[wib][(e2 >> NEIGH_IDX_BITS) & 31], pj2, e2, l2, nt, A);
      if (cnt > 3 * TILE) {
        const bool l3 = (unsigned)(3 * TILE + lane) < cnt;
        const unsigned e3 = l3 ? __ldg(&run[3 * TILE + lane]) : 0u;
        const int4 pj3 = __ldg(&posr[l3 ? (int)(e3 & NEIGH_IDX_MASK) : i]);
        pair3<EV, UNI>(sfx[3 * TILE + lane], sfy[3 * TILE + lane], sfz[3 * TILE + lane], s_pos[wib][(e3 >> NEIGH_IDX_BITS) & 31], pj3, e3, l3, nt, A);
      }
      __syncwarp();
      const int hi = min(off + nn, STEP3_CH * TILE);
#pragma unroll 1
      for (int k = off; k < hi; k++) { fx += sfx[k]; fy += sfy[k]; fz += sfz[k]; }
      __syncwarp();
    }
#pragma unroll 1
    for (unsigned base = STEP3_CH * TILE; base < cnt; base += STEP3_CH * TILE) {   // dense systems only
      unsigned e[STEP3_CH]; int4 pj[STEP3_CH]; bool lv[STEP3_CH];
#pragma unroll
      for (int r = 0; r < STEP3_CH; r++) { lv[r] = base + r * TILE + lane < cnt; e[r] = lv[r] ? __ldg(&run[base + r * TILE + lane]) : 0u; }
#pragma unroll
      for (int r = 0; r < STEP3_CH; r++) pj[r] = __ldg(&posr[lv[r] ? (int)(e[r] & NEIGH_IDX_MASK) : i]);
#pragma unroll
      for (int r = 0; r < STEP3_CH; r++)
        pair3<EV, UNI>(sfx[r * TILE + lane], sfy[r * TILE + lane], sfz[r * TILE + lane], s_pos[wib][(e[r] >> NEIGH_IDX_BITS) & 31], pj[r], e[r], lv[r], nt, A);
      __syncwarp();
      const int lo = max(off, (int)base), hi = min(off + nn, (int)base + STEP3_CH * TILE);
      for (int k = lo; k < hi; k++) { fx += sfx[k - (int)base]; fy += sfy[k - (int)base]; fz += sfz[k - (int)base]; }
      __syncwarp();
    }

    // ---- bonds, fp64, lane per owner ----
    if (0 < nb) bond3<EV>(fx, fy, fz, ctrl, pi, pb0, eb0, tag, A);
    if (1 < nb) bond3<EV>(fx, fy, fz, ctrl, pi, pb1, eb1, tag, A);
    if (2 < nb) bond3<EV>(fx, fy, fz, ctrl, pi, pb2, eb2, tag, A);
#pragma unroll 1
    for (int m = 3; m < nb; m++) {
      const unsigned e = __ldg(&bondrow[(size_t)m * cap + i]);
      bond3<EV>(fx, fy, fz, ctrl, pi, __ldg(&posr[e & BOND_IDX_MASK]), e, tag, A);
    }
    if (!valid) continue;                                      // (no warp-level operation below this line)
    if (a.angles) { fx += d.fang[3 * (size_t)i]; fy += d.fang[3 * (size_t)i + 1]; fz += d.fang[3 * (size_t)i + 2]; }   // angle_style cosine (le_angle.cuh)

    if (a.write_force) {
      double *fo = d.fout + (size_t)(tag - 1) * 3;
      fo[0] = fx; fo[1] = fy; fo[2] = fz;
    }

    // ---- Langevin drag + uniform noise (post_force); fp32, added to the rounded conservative force ----
    float lx = 0.f, ly = 0.f, lz = 0.f;
    if (a.langevin) step3_langevin(lx, ly, lz, ctrl, vi, tag, ti, step);

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
    if (EV) ke += (double)c_P.mass[ti] * ((double)vi.x * vi.x + (double)vi.y * vi.y + (double)vi.z * vi.z);
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
      if (DD) {   // halo update fused into the integrator: boundary atoms are also stored into the neighbor GPU's ghost slots
        if (i < ctrl->send_l_end) d.peer[left_rank(d)].pos[rd ^ 1][d.gr0 + (i - d.own0)] = pnew;
        const int srb = ctrl->send_r_beg;
        if (i >= srb) d.peer[right_rank(d)].pos[rd ^ 1][i - srb] = pnew;
      }
      // displacement since the last rebuild (Neighbor::check_distance): the path length of this step, rounded up, joins
      // the bound; pos_hold is consulted only once the bound has reached skin/2
      const float hx = __fmul_rn((float)dux, sx), hy = __fmul_rn((float)duy, sy), hz = __fmul_rn((float)duz, sz);
      const float len = sqrtf(__fmaf_rn(hz, hz, __fmaf_rn(hx, hx, __fmul_rn(hy, hy))));
      const unsigned binc = min(__float2uint_ru(__fmul_ru(__fmul_ru(len, 1.000001f), c_P.inv_bound_unit)), AUX_BOUND_MAX);
      const unsigned bound = min(AUX_BOUND(aux) + binc + 1u, AUX_BOUND_MAX);
      if (bound >= AUX_BOUND_ONE) {
        const int4 ph = d.pos_hold[i];
        const float qx = __fmul_rn((float)(int)(nx - (unsigned)ph.x), sx);
        const float qy = __fmul_rn((float)(int)(ny - (unsigned)ph.y), sy);
        const float qz = __fmul_rn((float)(int)(nz - (unsigned)ph.z), sz);
        if (__fmaf_rn(qz, qz, __fmaf_rn(qx, qx, __fmul_rn(qy, qy))) > c_P.triggersq) ctrl->moved = 1;
      }
      vi.w = __uint_as_float((aux & 0xfffu) | (bound << 12));
    }
    d.vel[i] = vi;
  }

  if (EV) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {      // what `thermo_style custom ... bonds f_ID[k]` prints on this step
      double *sl = d.thermo + (size_t)a.slot * LE_THERMO_W;
      sl[16] = (double)ctrl->nbonds;
#pragma unroll
      for (int q = 0; q < 3; q++) { sl[17 + q] = (double)ctrl->le_count[q]; sl[20 + q] = (double)ctrl->le_count[4 + q]; }
    }
    double acc[10];
    acc[0] = ke; acc[1] = 0.5 * A.evdwl; acc[2] = 0.5 * A.ebond;
#pragma unroll
    for (int q = 0; q < 6; q++) acc[3 + q] = 0.5 * A.v[q];     // every pair / bond is seen from both ends
    acc[9] = A.warn;
    __shared__ double red[STEP3_WARPS][10];
#pragma unroll
    for (int q = 0; q < 10; q++) {
      const double s = warp_sum(acc[q]);
      if (lane == 0) red[wib][q] = s;
    }
    __syncthreads();
    if (wib == 0) {
#pragma unroll
      for (int q = 0; q < 10; q++) {
        double s = (lane < STEP3_WARPS) ? red[lane][q] : 0.0;
        s = warp_sum(s);
        if (lane == 0 && s != 0.0) atomicAdd(&d.thermo[(size_t)a.slot * LE_THERMO_W + q], s);
      }
    }
  }
}
