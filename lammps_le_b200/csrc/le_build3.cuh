// le_build3.cuh -- rebuild, part 3: neighbor + bond list build (third generation).
//
// What the round-1 kernel (one thread per atom walking its 3x3 cell columns one after the other, gathering six
// tag-ordered topology tables) cost at 10^6 beads: 215 us, 4214 warp instructions per 32 atoms at 16 active lanes,
// 6.6 x the algorithmic DRAM bytes (profiles/r01_ncu_full_kstep_kbuild.txt).  This kernel keeps "one lane = one
// owned atom" (a dilute chain has ~19 candidates per atom: fewer than a warp has lanes, so "lanes = candidates" would
// idle a third of the warp and pay the window bookkeeping once per atom instead of once per 32) and changes the rest:
//   * the 9 (18 with a z wrap) candidate windows of an atom -- slot ranges of the local order, three z cells of one
//     column each -- are fetched up front as one batch of independent loads and kept in shared memory;
//   * the windows are walked as ONE flattened candidate stream, four independent position loads per trip: a warp makes
//     max-over-lanes(candidates)/4 trips instead of sum-over-columns(max-over-lanes) ones;
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
__global__ void k_topo_pack(Dev d) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < d.N; t += gridDim.x * blockDim.x) {
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
#define BUILD_MAXWIN 18

// queue entry of a screened candidate: x = slot, y = tag << 1 | "inside the fp64 sliver"
template <int QCAP, int MINB>
__global__ void __launch_bounds__(BUILD_THREADS, MINB) k_build3(Dev d) {
  __shared__ int2 s_q[QCAP][BUILD_THREADS];
  __shared__ int2 s_win[BUILD_MAXWIN][BUILD_THREADS];
  const int cap = d.cap;
  const int t = threadIdx.x, lane = t & 31;
  const int own_end = d.own0 + d.ctrl->nown;
  const int i = d.own0 + blockIdx.x * BUILD_THREADS + t;
  if (i - lane >= own_end) return;                       // the whole warp (= tile) lies beyond the owned atoms
  const int4 *__restrict__ ph = d.pos_hold;
  const bool active = i < own_end;
  int n = 0;                                             // accepted neighbors of this atom
  if (active) {
    const int cur = d.ctrl->cur;
    const int4 pi = ph[i];
    const int tagi = pi.w >> 3, ti = pi.w & 7, nt = c_P.ntypes;
    const float4 vt = d.vel_tmp[i];
    const int imh = d.img_hold[i];
    const TopoRec *__restrict__ tr = d.topo + (tagi - 1);
    const int4 r0 = __ldg(reinterpret_cast<const int4 *>(tr));        // hdr, btypes, batom 0, 1
    const int4 r1 = __ldg(reinterpret_cast<const int4 *>(tr) + 1);    // batom 2, 3, spec 0, 1
    const int4 r2 = __ldg(reinterpret_cast<const int4 *>(tr) + 2);    // spec 2..5

    // ---- the candidate windows: one batch of independent cell_start loads ----
    const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
    const int cx = __umulhi((unsigned)pi.x, (unsigned)ncx);
    const int cy = __umulhi((unsigned)pi.y, (unsigned)ncy);
    const int cz = __umulhi((unsigned)pi.z, (unsigned)ncz);
    const int lx = local_layer(d, cx);
    // per column (lx', cy') the three z-cells are one contiguous range of the local order; a column that wraps in z
    // gets its far cell as a second, single-cell window
    const int zlo = d.cell_abs[2] ? 0 : max(cz - 1, 0), zhi = d.cell_abs[2] ? ncz - 1 : min(cz + 1, ncz - 1);
    const int zwrap = d.cell_abs[2] ? -1 : (cz == 0 ? ncz - 1 : (cz == ncz - 1 ? 0 : -1));
    int nw = 0;
    for (int ox = 0; ox < d.cell_span[0]; ox++) {
      int xc = d.cell_abs[0] ? ox : lx - 1 + ox;
      if (d.nranks == 1) { if (xc < 0) xc += ncx; else if (xc >= ncx) xc -= ncx; }   // one GPU: the slab is the whole box
      for (int oy = 0; oy < d.cell_span[1]; oy++) {
        int yc = d.cell_abs[1] ? oy : cy - 1 + oy;
        if (yc < 0) yc += ncy; else if (yc >= ncy) yc -= ncy;
        const int base = cell_slot(d, xc, yc, 0);
        s_win[nw++][t] = make_int2(__ldg(&d.cell_start[base + zlo]), __ldg(&d.cell_start[base + zhi + 1]));
        if (zwrap >= 0) s_win[nw++][t] = make_int2(__ldg(&d.cell_start[base + zwrap]), __ldg(&d.cell_start[base + zwrap + 1]));
      }
    }
    // the sorted state goes back into the live arrays (the aux word of the velocity follows at the end)
    d.pos[cur][i] = pi;
    d.img[i] = imh;

    const float fsx = c_P.fscale[0], fsy = c_P.fscale[1], fsz = c_P.fscale[2];
    const bool uni = c_P.pair_uniform != 0;
    const float hi_u = c_P.cutneigh_hi[0], lo_u = c_P.cutneigh_lo[0];
    SpecCtx S;
    S.nscan = (r0.x >> 8) & 0xff; S.n1 = (r0.x >> 16) & 0xff; S.n2 = (r0.x >> 24) & 0xff;
    S.s0 = r1.z; S.s1 = r1.w; S.s2 = r2.x; S.s3 = r2.y;
    S.rec_spec = tr->spec; S.row = d.special + (size_t)(tagi - 1) * d.maxspecial;
    unsigned *__restrict__ ell = d.nbr_ell + i;          // overflow rows (column i)
    int nq = 0, novf = 0;

    // decide one screened candidate (special bits, fp64 sliver); returns the entry or -1
    auto decide = [&](int j, int tagj, bool sliver, int tj) -> int {
      const int which = find_special3(S, tagj);
      if (which < 0) return -1;
      if (sliver && !build_border(pi, ph[j], uni ? 0 : ti * nt + tj)) return -1;
      return (int)((unsigned)j | ((unsigned)which << 30));
    };

    // ---- phase 1: fp32 screen of the flattened candidate stream, four independent loads per trip ----
    int w = 0, j = 0, jend = 0;
    for (;;) {
      int c[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        while (j >= jend && w < nw) { const int2 q = s_win[w++][t]; j = q.x; jend = q.y; }
        c[u] = (j < jend) ? j++ : -1;
      }
      if (c[0] < 0) break;
      int4 p[4];
#pragma unroll
      for (int u = 0; u < 4; u++) p[u] = __ldg(&ph[max(c[u], 0)]);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float fx = (float)(int)((unsigned)p[u].x - (unsigned)pi.x) * fsx;
        const float fy = (float)(int)((unsigned)p[u].y - (unsigned)pi.y) * fsy;
        const float fz = (float)(int)((unsigned)p[u].z - (unsigned)pi.z) * fsz;
        const float rsqf = fx * fx + fy * fy + fz * fz;
        const int tj = p[u].w & 7;
        const float hi = uni ? hi_u : c_P.cutneigh_hi[ti * nt + tj], lo = uni ? lo_u : c_P.cutneigh_lo[ti * nt + tj];
        if (rsqf <= hi && c[u] >= 0 && c[u] != i) {
          const int tagj = p[u].w >> 3;
          if (nq < QCAP) { s_q[nq][t] = make_int2(c[u], (tagj << 1) | (rsqf >= lo ? 1 : 0)); nq++; }
          else {
            // the queue is full (dense systems): decide at once, park the entry in the per-atom overflow rows
            const int e = decide(c[u], tagj, rsqf >= lo, tj);
            if (e >= 0) {
              if (QCAP + novf >= d.maxneigh) le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, d.maxneigh);
              else { ell[(size_t)novf * cap] = (unsigned)e; novf++; }
            }
          }
        }
      }
    }
    // ---- phase 2: decide the queued candidates; the accepted ones are compacted to the front of the queue ----
    int na = 0;
    for (int q = 0; q < nq; q++) {
      const int2 e = s_q[q][t];
      const int tagj = e.y >> 1;
      int tj = 0;
      if (!uni && (e.y & 1)) tj = ph[e.x].w & 7;
      const int r = decide(e.x, tagj, (e.y & 1) != 0, tj);
      if (r >= 0) { s_q[na][t].x = r; na++; }
    }
    n = na + novf;

    // ---- bond partner rows (the partners' slots come from the tag map written by k_permute / k_ghost_map) ----
    const int nb = r0.x & 0xff;
    bool missing = false;
    {
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
    }
    if (missing) le_raise(d.ctrl, LE_DERR_MISSING_ATOM, tagi, nb);
    if (n > 255) { le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, n); n = 255; }
    d.vel[i] = make_float4(vt.x, vt.y, vt.z, __uint_as_float(AUX_PACK(n, nb, 0)));
    // stash for the write-out below
    s_win[0][t] = make_int2(na, novf);
  }
  __syncwarp();
  // ---- pack the tile's run: exclusive scan of the counts over the warp, entries grouped by owner ----
  int inc = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  const int off = inc - n;
  const int total = __shfl_sync(0xffffffffu, inc, 31);
  const int tile = (i - lane - d.own0) >> 5;
  unsigned *__restrict__ run = d.nbr + (size_t)tile * d.tcap;
  if (active) {
    const int na = s_win[0][t].x;
    const unsigned own = (unsigned)lane << NEIGH_IDX_BITS;
    for (int k = 0; k < na; k++) run[off + k] = (unsigned)s_q[k][t].x | own;
    for (int k = na; k < n; k++) run[off + k] = d.nbr_ell[(size_t)(k - na) * cap + i] | own;
  }
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
