// le_build3.cuh -- rebuild, part 3: neighbor + bond list build (third generation).
//
// What the round-1 kernel (one thread per atom walking its 3x3 cell columns one after the other, gathering six
// tag-ordered topology tables) cost at 10^6 beads: 215 us, 4214 warp instructions per 32 atoms at 16 active lanes,
// 6.6 x the algorithmic DRAM bytes (profiles/r01_ncu_full_kstep_kbuild.txt).  This kernel keeps "one lane = one
// owned atom" (a dilute chain has ~19 candidates per atom: fewer than a warp has lanes, so "lanes = candidates" would
// idle a third of the warp and pay the window bookkeeping once per atom instead of once per 32) and changes the rest:
//   * the candidate windows of an atom -- slot ranges of the local order, three z cells of one column each -- are
//     fetched three at a time (one x layer) as independent loads; the warp walks a window in lock step (trip count =
//     the longest window among its lanes), two independent position loads per trip, no lane ever leaves a loop early;
//   * a candidate that passes the fp32 distance screen is only queued (shared memory, slot + tag + "in the fp64 sliver");
//     the special-list look-up and the fp64 re-check run afterwards over the queue, every lane busy with its own entries;
//   * topology comes from one 64-byte digest per atom (TopoRec, k_topo_pack) instead of num_bond / bond_atom /
//     bond_type / nspecial / special gathers;
//   * accepted pairs are packed into the tile's flat run (le_common.cuh) by a warp scan of the per-lane counts.
// Pair acceptance is unchanged and bit-exact: fp32 on exact fixed-point differences, the 1e-5 sliver around
// cutneighsq decided by the reference's fp64 arithmetic (npair_half_bin_newton.cpp:98-103); special bits as
// NPair::find_special (src/npair.h:112-136); bond partner rows as NTopoBondAll::build (src/ntopo_bond_all.cpp:39-86).
#pragma once
#include "le_common.cuh"

// ---- topology digest -----------------------------------------------------------------------------------------
__device__ __forceinline__ void topo_pack_one(const Dev &d, int t) {   // t = tag - 1
  const int nb = d.num_bond[t];
  const int *ns = d.nspecial + (size_t)t * 3;
  const int n1 = ns[0], n2 = ns[1], n3 = ns[2];
  const int nscan = c_P.nscan_tier == 0 ? 0 : c_P.nscan_tier == 1 ? n1 : c_P.nscan_tier == 2 ? n2 : n3;
  TopoRec r;
  r.hdr = (unsigned)nb | ((unsigned)nscan << 8) | ((unsigned)n1 << 16) | ((unsigned)n2 << 24);
  unsigned bt = 0;
  const int *ba = d.bond_atom + (size_t)t * d.bpa, *bty = d.bond_type + (size_t)t * d.bpa;
  for (int m = 0; m < nb && m < 8; m++) bt |= ((unsigned)(bty[m] - 1) & 15u) << (4 * m);
  r.btypes = bt;
#pragma unroll
  for (int m = 0; m < 4; m++) r.batom[m] = (m < nb) ? ba[m] : 0;
  const int *sp = d.special + (size_t)t * d.maxspecial;
#pragma unroll
  for (int q = 0; q < TOPO_NSPEC; q++) r.spec[q] = (q < n3) ? sp[q] : 0;
  int4 *out = reinterpret_cast<int4 *>(d.topo + t);
  const int4 *in = reinterpret_cast<const int4 *>(&r);
  out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; out[3] = in[3];
}
__global__ void k_topo_pack(Dev d) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < d.N; t += gridDim.x * blockDim.x) topo_pack_one(d, t);
}

// NPair::find_special (src/npair.h:112-136) on the digest: the first four specials sit in registers
struct SpecCtx { int s0, s1, s2, s3, n1, n2, nscan; const int *rec_spec; const int *row; };
__device__ __forceinline__ int find_special3(const SpecCtx &S, int tag) {
  int k = -1;
  if (S.nscan > 0 && S.s0 == tag) k = 0;
  else if (S.nscan > 1 && S.s1 == tag) k = 1;
  else if (S.nscan > 2 && S.s2 == tag) k = 2;
  else if (S.nscan > 3 && S.s3 == tag) k = 3;
  else
    for (int q = 4; q < S.nscan; q++)
      if ((q < TOPO_NSPEC ? S.rec_spec[q] : S.row[q]) == tag) { k = q; break; }
  if (k < 0) return 0;
  const int tier = (k < S.n1) ? 1 : (k < S.n2) ? 2 : 3;
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

#define BUILD_THREADS 128

// queue entry of a screened candidate: x = slot, y = tag << 1 | "inside the fp64 sliver".
// The warp stays converged through all phases: every loop runs to the warp-wide maximum of its trip count
// (__reduce_max_sync), lanes without work in a trip are predicated off.  (A first version let every lane walk its own
// flattened candidate stream; the lanes left that loop at different times and the compiler reconverged them only at the
// end of the kernel, so the queue pass and the write-out ran 3.8 times per warp with 8 lanes: 5230 warp instructions per
// 32 atoms, profiles/r02_build_step_first.txt.)
template <int QCAP, int MINB, int UNI>
__global__ void __launch_bounds__(BUILD_THREADS, MINB) k_build3(Dev d) {
  __shared__ int2 s_q[QCAP][BUILD_THREADS];
  const unsigned FULL = 0xffffffffu;
  const int cap = d.cap;
  const int t = threadIdx.x, lane = t & 31;
  const int own_end = d.own0 + d.ctrl->nown;
  const int i0 = d.own0 + blockIdx.x * BUILD_THREADS + t;
  if (i0 - lane >= own_end) return;                      // the whole warp (= tile) lies beyond the owned atoms
  const int4 *__restrict__ ph = d.pos_hold;
  const bool active = i0 < own_end;
  const int i = active ? i0 : own_end - 1;               // lanes beyond the end shadow the last atom (loads only)
  const int cur = d.ctrl->cur;
  const int4 pi = ph[i];
  const int tagi = pi.w >> 3, ti = pi.w & 7, nt = c_P.ntypes;
  const float4 vt = d.vel_tmp[i];
  const int imh = d.img_hold[i];
  const TopoRec *__restrict__ tr = d.topo + (tagi - 1);
  const int4 r0 = __ldg(reinterpret_cast<const int4 *>(tr));        // hdr, btypes, batom 0, 1
  const int4 r1 = __ldg(reinterpret_cast<const int4 *>(tr) + 1);    // batom 2, 3, spec 0, 1
  const int4 r2 = __ldg(reinterpret_cast<const int4 *>(tr) + 2);    // spec 2..5

  const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
  const int cx = __umulhi((unsigned)pi.x, (unsigned)ncx);
  const int cy = __umulhi((unsigned)pi.y, (unsigned)ncy);
  const int cz = __umulhi((unsigned)pi.z, (unsigned)ncz);
  const int lx = local_layer(d, cx);
  // per column (lx', cy') the three z-cells are one contiguous range of the local order (a "window"); a column that
  // wraps in z gets its far cell as a second, single-cell window (pass 1)
  const int zlo = d.cell_abs[2] ? 0 : max(cz - 1, 0), zhi = d.cell_abs[2] ? ncz - 1 : min(cz + 1, ncz - 1);
  const int zwrap = d.cell_abs[2] ? -1 : (cz == 0 ? ncz - 1 : (cz == ncz - 1 ? 0 : -1));
  const int npass = __any_sync(FULL, active && zwrap >= 0) ? 2 : 1;
  if (active) {   // the sorted state goes back into the live arrays (the aux word of the velocity follows at the end)
    d.pos[cur][i] = pi;
    d.img[i] = imh;
  }

  const float fsx = c_P.fscale[0], fsy = c_P.fscale[1], fsz = c_P.fscale[2];
  const float hi_u = c_P.cutneigh_hi[0], lo_u = c_P.cutneigh_lo[0];
  SpecCtx S;
  S.nscan = (r0.x >> 8) & 0xff; S.n1 = (r0.x >> 16) & 0xff; S.n2 = (r0.x >> 24) & 0xff;
  S.s0 = r1.z; S.s1 = r1.w; S.s2 = r2.x; S.s3 = r2.y;
  S.rec_spec = tr->spec; S.row = d.special + (size_t)(tagi - 1) * d.maxspecial;
  unsigned *__restrict__ ell = d.nbr_ell + i;          // overflow rows (column i)
  int nq = 0, novf = 0;

  // decide one screened candidate (special bits, fp64 sliver); returns the entry or BUILD_REJECT
  const unsigned BUILD_REJECT = 0xffffffffu;
  auto decide = [&](int j, int tagj, bool sliver, int tj) -> unsigned {
    const int which = find_special3(S, tagj);
    if (which < 0) return BUILD_REJECT;
    if (sliver && !build_border(pi, ph[j], UNI ? 0 : ti * nt + tj)) return BUILD_REJECT;
    return (unsigned)j | ((unsigned)which << 30);
  };
  // fp32 screen of one candidate; a survivor is queued (or, queue full, decided at once and parked in the overflow rows)
  auto screen = [&](const int4 pj, int j, bool live) {
    const float fx = (float)(int)((unsigned)pj.x - (unsigned)pi.x) * fsx;
    const float fy = (float)(int)((unsigned)pj.y - (unsigned)pi.y) * fsy;
    const float fz = (float)(int)((unsigned)pj.z - (unsigned)pi.z) * fsz;
    const float rsqf = fx * fx + fy * fy + fz * fz;
    const int tj = pj.w & 7;
    const float hi = UNI ? hi_u : c_P.cutneigh_hi[ti * nt + tj], lo = UNI ? lo_u : c_P.cutneigh_lo[ti * nt + tj];
    if (live && rsqf <= hi && j != i) {
      const int tagj = pj.w >> 3;
      if (nq < QCAP) s_q[nq][t] = make_int2(j, (tagj << 1) | (rsqf >= lo ? 1 : 0));
      else {
        const unsigned e = decide(j, tagj, rsqf >= lo, tj);
        if (e != BUILD_REJECT) {
          if (QCAP + novf >= d.maxneigh) le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, d.maxneigh);
          else { ell[(size_t)novf * cap] = e; novf++; }
        }
      }
      nq++;
    }
  };

  // ---- phase 1: fp32 screen, window by window; two independent candidate loads per trip ----
  for (int pass = 0; pass < npass; pass++) {
    const int za = pass ? zwrap : zlo, zb = pass ? zwrap : zhi;
    const bool on = active && (pass == 0 || zwrap >= 0);
    for (int ox = 0; ox < d.cell_span[0]; ox++) {
      int xc = d.cell_abs[0] ? ox : lx - 1 + ox;
      if (d.nranks == 1) { if (xc < 0) xc += ncx; else if (xc >= ncx) xc -= ncx; }   // one GPU: the slab is the whole box
      int wlo[3], whi[3];
#pragma unroll
      for (int oy = 0; oy < 3; oy++) {
        wlo[oy] = whi[oy] = 0;
        if (oy < d.cell_span[1] && on) {
          int yc = d.cell_abs[1] ? oy : cy - 1 + oy;
          if (yc < 0) yc += ncy; else if (yc >= ncy) yc -= ncy;
          const int base = cell_slot(d, xc, yc, 0);
          wlo[oy] = __ldg(&d.cell_start[base + za]); whi[oy] = __ldg(&d.cell_start[base + zb + 1]);
        }
      }
#pragma unroll
      for (int oy = 0; oy < 3; oy++) {
        const int lo = wlo[oy], hi = whi[oy];
        const int maxlen = __reduce_max_sync(FULL, hi - lo);
        for (int k = 0; k < maxlen; k += 2) {
          const int j0 = lo + k, j1 = j0 + 1;
          const bool v0 = j0 < hi, v1 = j1 < hi;
          const int4 p0 = __ldg(&ph[v0 ? j0 : i]), p1 = __ldg(&ph[v1 ? j1 : i]);
          screen(p0, j0, v0);
          screen(p1, j1, v1);
        }
      }
    }
  }
  // ---- phase 2: decide the queued candidates; the accepted ones are compacted to the front of the queue ----
  int na = 0;
  {
    const int nqq = min(nq, QCAP);
    const int maxq = __reduce_max_sync(FULL, nqq);
    for (int q = 0; q < maxq; q++) {
      if (q < nqq) {
        const int2 e = s_q[q][t];
        int tj = 0;
        if (!UNI && (e.y & 1)) tj = ph[e.x].w & 7;
        const unsigned r = decide(e.x, e.y >> 1, (e.y & 1) != 0, tj);
        if (r != BUILD_REJECT) { s_q[na][t].x = (int)r; na++; }
      }
    }
  }
  int n = active ? na + novf : 0;
  if (n > d.maxneigh) {                                    // neigh_modify one: the tile's run holds 32 * maxneigh entries
    le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, n, d.maxneigh);
    n = d.maxneigh; if (na > n) na = n;
  }

  // ---- bond partner rows (the partners' slots come from the tag map written by k_permute / k_ghost_map) ----
  const int nb = r0.x & 0xff;
  if (active) {
    bool missing = false;
    const int bp[4] = {r0.z, r0.w, r1.x, r1.y};
#pragma unroll
    for (int m = 0; m < 4; m++)
      if (m < nb) {
        const int jb = __ldg(&d.map[bp[m] - 1]);
        if (jb < 0) missing = true;
        else d.bondrow[(size_t)m * cap + i] = (unsigned)jb | ((((unsigned)r0.y >> (4 * m)) & 15u) << 28);
      }
    for (int m = 4; m < nb; m++) {
      const int pt = d.bond_atom[(size_t)(tagi - 1) * d.bpa + m];
      const int bt = d.bond_type[(size_t)(tagi - 1) * d.bpa + m];
      const int jb = d.map[pt - 1];
      if (jb < 0) { missing = true; continue; }
      d.bondrow[(size_t)m * cap + i] = (unsigned)jb | ((unsigned)(bt - 1) << 28);
    }
    if (missing) le_raise(d.ctrl, LE_DERR_MISSING_ATOM, tagi, nb);
    if (n > 255) { le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, n); n = 255; if (na > 255) na = 255; }
    d.vel[i] = make_float4(vt.x, vt.y, vt.z, __uint_as_float(AUX_PACK(n, nb, 0)));
  }
  // ---- pack the tile's run: exclusive scan of the counts over the warp, entries grouped by owner ----
  int inc = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += v;
  }
  const int off = inc - n;
  const int total = __shfl_sync(FULL, inc, 31);
  const int tile = (i0 - lane - d.own0) >> 5;
  unsigned *__restrict__ run = d.nbr + (size_t)tile * d.tcap;
  const unsigned own = (unsigned)lane << NEIGH_IDX_BITS;
  const int maxn = __reduce_max_sync(FULL, n);
  for (int k = 0; k < maxn; k++)
    if (k < n) run[off + k] = (k < na ? (unsigned)s_q[k][t].x : d.nbr_ell[(size_t)(k - na) * cap + i]) | own;
  if (lane == 0) d.tile_cnt[tile] = (unsigned)total;
}

// list statistics on demand: out[0] unused, out[1] = full-list entries of the owned tiles
__global__ void k_count_pairs(Dev d, unsigned long long *out) {
  unsigned long long f = 0;
  const int ntiles = (d.ctrl->nown + TILE - 1) / TILE;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) f += d.tile_cnt[t];
  for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&out[1], f);
}
