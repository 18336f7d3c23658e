// le_step4.cuh -- k_step4: the fused timestep kernel, fourth generation.  Same division of labour and the SAME arithmetic,
// operation for operation and in the same order, as k_step3 (le_step3.cuh: pair-parallel WCA terms parked in shared
// memory and added by the owner in list order, lane-per-owner fp64 bonds, Philox Langevin noise, velocity Verlet,
// displacement bound) -- trajectories are bit-identical (tests/test_gpu_step3.py) -- with fewer instructions per tile:
// k_step3 executed 866 warp instructions per tile of 32 atoms at 10^6 beads (profiles/r02_ncu_full_kstep3_kbuild3.txt),
// of which 105 in the owner's add loop, ~215 in per-tile set-up (spilled gathers, recomputed lane / warp indices and
// constants under a 64-register cap) and 165 in the bonds.
//   * parked terms as {x,y} + {z}: two shared loads per term instead of three; the owner adds its first four terms
//     with predicated straight-line code (a bead of a dilute chain has 2.2 listed neighbors), a loop only beyond that;
//   * 128-thread blocks: the register budget can sit between the 64 / 80 steps of 256-thread blocks; the velocity is fetched
//     after the forces (only its aux word is needed before), which brings the UNI kernel to 64 registers: 8 resident blocks;
//   * the third bond slot and the third/fourth list round are fetched only by warps that have them;
//   * the step's path length for the displacement bound is the 1-norm of the step (>= the 2-norm: still an upper
//     bound of the displacement), no square root.
//   reference: PairLJCut::compute src/pair_lj_cut.cpp:68-140, BondFENE::compute src/MOLECULE/bond_fene.cpp:52-128,
//   BondHarmonic::compute bond_harmonic.cpp:48-100, FixLangevin::post_force_templated src/fix_langevin.cpp:587-777,
//   FixNVE::initial/final_integrate src/fix_nve.cpp:64-140, Neighbor::check_distance src/neighbor.cpp:1962-2014.
#pragma once
#include "le_step3.cuh"

#define STEP4_THREADS 128
#define STEP4_WARPS (STEP4_THREADS / 32)

struct __align__(16) Step4Smem {
  int4 pos[TILE];
  double2 fxy[STEP3_CH * TILE];
  double fz[STEP3_CH * TILE];
};

template <int EV, int DD, int UNI>
__global__ void __launch_bounds__(STEP4_THREADS, EV ? 4 : (DD ? (UNI ? 7 : 6) : (UNI ? 8 : 7))) k_step4(Dev d, StepArgs a) {   // (one coefficient set: 64 registers without spills -> 8 blocks/SM; its slab form 71 -> 7)
  __shared__ Step4Smem s_all[STEP4_WARPS];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Step4Smem &S = s_all[wib];
  const int cap = d.cap;
  Ctrl *__restrict__ ctrl = d.ctrl;
  const int rd = a.rdp1 ? a.rdp1 - 1 : ctrl->cur;
  const int own_end = d.own0 + (DD ? ctrl->nown : d.N);        // one GPU owns every atom: no look at the control block
  const int4 *__restrict__ posr = d.pos[rd];
  int4 *__restrict__ posw = d.pos[rd ^ 1];
  const unsigned *__restrict__ bondrow = d.bondrow;
  const int nt = c_P.ntypes;
  const int tcap = d.tcap;
  const long long step = ctrl->step;
  __shared__ Step3Order s_ord;                                   // (shared memory: five registers less across the tile loop of the slab form)
  int ntiles = (own_end - d.own0 + TILE - 1) >> 5;
  if (DD) {
    if (threadIdx.x == 0) s_ord = step3_order(d, own_end);
    __syncthreads();
    ntiles = s_ord.total;
  }
  const Step3Order &ord = s_ord;

  EvAcc A;
  double ke = 0.0;
  if (EV) { A.evdwl = A.ebond = A.warn = 0.0; for (int q = 0; q < 6; q++) A.v[q] = 0.0; }

#pragma unroll 1
  for (int g = blockIdx.x * STEP4_WARPS + wib; g < ntiles; g += gridDim.x * STEP4_WARPS) {
    const int tile = DD ? step3_tile_of(ord, g) : g;
    const int i0 = d.own0 + tile * TILE;
    const bool valid = i0 + lane < own_end;
    const int i = valid ? i0 + lane : own_end - 1;            // lanes beyond the end shadow the last atom (loads only)
    // ---- level 1: everything addressed by the tile / the atom ----
    const int4 pi = posr[i];
    const unsigned aux = __float_as_uint(d.vel[i].w);          // counts + displacement bound now; the velocity itself is fetched after the forces (3 registers less across the tile)
    const unsigned cnt = __ldg(&d.tile_cnt[tile]);
    const unsigned *__restrict__ run = d.nbr + (size_t)tile * tcap;
    const unsigned e0 = __ldg(&run[lane]), e1 = __ldg(&run[TILE + lane]);   // tcap >= 128
    const unsigned eb0 = __ldg(&bondrow[i]);
    const unsigned eb1 = d.bpa > 1 ? __ldg(&bondrow[(size_t)cap + i]) : 0u;
    const int nn = valid ? (int)AUX_NN(aux) : 0, nb = valid ? (int)AUX_NB(aux) : 0;
    const int ti = pi.w & 7, tag = pi.w >> 3;
    S.pos[lane] = pi;
    // start of this lane's entries inside the run
    int inc = nn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    const int off = inc - nn;
    __syncwarp();
    // ---- level 2: the gathers of the first two rounds and of the first two bond partners ----
    const bool l0 = (unsigned)lane < cnt, l1 = (unsigned)(TILE + lane) < cnt;
    const int4 pj0 = __ldg(&posr[l0 ? (int)(e0 & NEIGH_IDX_MASK) : i]);
    const int4 pj1 = __ldg(&posr[l1 ? (int)(e1 & NEIGH_IDX_MASK) : i]);
    const int4 pb0 = __ldg(&posr[0 < nb ? (int)(eb0 & BOND_IDX_MASK) : i]);
    const int4 pb1 = __ldg(&posr[1 < nb ? (int)(eb1 & BOND_IDX_MASK) : i]);

    // ---- pairs: STEP3_CH rounds per pass; the owner adds its terms in list order after each pass ----
    double fx = 0.0, fy = 0.0, fz = 0.0;
    {
      double tx, ty, tz;
      pair3<EV, UNI>(tx, ty, tz, S.pos[(e0 >> NEIGH_IDX_BITS) & 31], pj0, e0, l0, nt, A);
      S.fxy[lane] = make_double2(tx, ty); S.fz[lane] = tz;
      pair3<EV, UNI>(tx, ty, tz, S.pos[(e1 >> NEIGH_IDX_BITS) & 31], pj1, e1, l1, nt, A);
      S.fxy[TILE + lane] = make_double2(tx, ty); S.fz[TILE + lane] = tz;
      if (cnt > 2 * TILE) {                                   // warp-uniform
        const bool l2 = (unsigned)(2 * TILE + lane) < cnt, l3 = (unsigned)(3 * TILE + lane) < cnt;
        const unsigned e2 = l2 ? __ldg(&run[2 * TILE + lane]) : 0u, e3 = l3 ? __ldg(&run[3 * TILE + lane]) : 0u;
        const int4 pj2 = __ldg(&posr[l2 ? (int)(e2 & NEIGH_IDX_MASK) : i]);
        const int4 pj3 = __ldg(&posr[l3 ? (int)(e3 & NEIGH_IDX_MASK) : i]);
        pair3<EV, UNI>(tx, ty, tz, S.pos[(e2 >> NEIGH_IDX_BITS) & 31], pj2, e2, l2, nt, A);
        S.fxy[2 * TILE + lane] = make_double2(tx, ty); S.fz[2 * TILE + lane] = tz;
        if (cnt > 3 * TILE) {
          pair3<EV, UNI>(tx, ty, tz, S.pos[(e3 >> NEIGH_IDX_BITS) & 31], pj3, e3, l3, nt, A);
          S.fxy[3 * TILE + lane] = make_double2(tx, ty); S.fz[3 * TILE + lane] = tz;
        }
      }
      __syncwarp();
      const int hi = min(off + nn, STEP3_CH * TILE);
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (off + q < hi) { const double2 t = S.fxy[off + q]; const double z = S.fz[off + q]; fx += t.x; fy += t.y; fz += z; }
#pragma unroll 1
      for (int k = off + 4; k < hi; k++) { const double2 t = S.fxy[k]; const double z = S.fz[k]; fx += t.x; fy += t.y; fz += z; }
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
      for (int r = 0; r < STEP3_CH; r++) {
        double tx, ty, tz;
        pair3<EV, UNI>(tx, ty, tz, S.pos[(e[r] >> NEIGH_IDX_BITS) & 31], pj[r], e[r], lv[r], nt, A);
        S.fxy[r * TILE + lane] = make_double2(tx, ty); S.fz[r * TILE + lane] = tz;
      }
      __syncwarp();
      const int lo = max(off, (int)base), hi = min(off + nn, (int)base + STEP3_CH * TILE);
      for (int k = lo; k < hi; k++) { const double2 t = S.fxy[k - (int)base]; const double z = S.fz[k - (int)base]; fx += t.x; fy += t.y; fz += z; }
      __syncwarp();
    }

    // ---- bonds, fp64, lane per owner ----
    if (0 < nb) bond3<EV>(fx, fy, fz, ctrl, pi, pb0, eb0, tag, A);
    if (1 < nb) bond3<EV>(fx, fy, fz, ctrl, pi, pb1, eb1, tag, A);
#pragma unroll 1
    for (int m = 2; m < nb; m++) {                             // third and later slots: beads that carry an extruder
      const unsigned e = __ldg(&bondrow[(size_t)m * cap + i]);
      bond3<EV>(fx, fy, fz, ctrl, pi, __ldg(&posr[e & BOND_IDX_MASK]), e, tag, A);
    }
    if (!valid) continue;                                      // (no warp-level operation below this line)
    float4 vi = d.vel[i];
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
      // displacement since the last rebuild (Neighbor::check_distance): an upper bound of this step's path length (its
      // 1-norm, rounded up) joins the bound; pos_hold is consulted only once the bound has reached skin/2
      const float len = __fadd_ru(__fadd_ru(fabsf(__fmul_ru(fabsf((float)dux), c_P.fscale[0])), fabsf(__fmul_ru(fabsf((float)duy), c_P.fscale[1]))),
                                  fabsf(__fmul_ru(fabsf((float)duz), c_P.fscale[2])));
      const unsigned binc = min(__float2uint_ru(__fmul_ru(__fmul_ru(len, 1.000001f), c_P.inv_bound_unit)), AUX_BOUND_MAX);
      const unsigned bound = min(AUX_BOUND(aux) + binc + 1u, AUX_BOUND_MAX);
      if (bound >= AUX_BOUND_ONE) {
        const int4 ph = d.pos_hold[i];
        const float qx = __fmul_rn((float)(int)(nx - (unsigned)ph.x), c_P.fscale[0]);
        const float qy = __fmul_rn((float)(int)(ny - (unsigned)ph.y), c_P.fscale[1]);
        const float qz = __fmul_rn((float)(int)(nz - (unsigned)ph.z), c_P.fscale[2]);
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
    __shared__ double red[STEP4_WARPS][10];
#pragma unroll
    for (int q = 0; q < 10; q++) {
      const double s = warp_sum(acc[q]);
      if (lane == 0) red[wib][q] = s;
    }
    __syncthreads();
    if (wib == 0) {
#pragma unroll
      for (int q = 0; q < 10; q++) {
        double s = (lane < STEP4_WARPS) ? red[lane][q] : 0.0;
        s = warp_sum(s);
        if (lane == 0 && s != 0.0) atomicAdd(&d.thermo[(size_t)a.slot * LE_THERMO_W + q], s);
      }
    }
  }
}
