// le_build4.cuh -- rebuild, part 3: neighbor + bond list build (fourth generation: warp-cooperative work list).
//
// k_build3 kept "one lane = one owned atom" through the distance screen: the warp walked every candidate window in lock
// step for as many trips as its longest lane needed, 25 instructions per trip -- 4100 warp instructions per tile of 32
// atoms at 10^6 beads for 600 useful distance tests (profiles/r02_ncu_build3_step3.txt).  Here a warp still owns one
// tile, but the expensive parts run on dense chunks of 32 work items whatever the spread between the atoms:
//   1. every lane fetches the candidate windows of its atom (slot ranges of the local order: the 3 z-cells of each of
//      the 3x3 neighboring cell columns, plus single far cells where a column wraps in z) as independent loads and
//      parks the non-empty ones in shared memory;
//   2. EXPANSION: the windows are flattened into the warp's work list (shared memory), entry = owner lane | candidate
//      slot, owner-major; a lane writes its own entries at the offset a warp scan gave it -- the only loop whose trip
//      count is the maximum over the lanes, and its body is a store;
//   3. SCREEN: lanes = work-list entries, 64 per trip: owner position from shared memory, candidate position gathered
//      (neighboring entries are neighboring slots), fp32 distance test on exact fixed-point differences; survivors are
//      ballot-compacted into a small ring;
//   4. DECIDE: lanes = survivors, 32 per trip: NPair::find_special (src/npair.h:112-136) on the owner's topology digest
//      (shared memory), the reference's fp64 arithmetic for the 1e-5 sliver around cutneighsq
//      (npair_half_bin_newton.cpp:98-103), ballot-compacted append to the tile's run in global memory.
// Work-list order = owner-major, windows in k_build3's order, slots ascending: the tile's run comes out grouped by
// owner in the same per-atom order as before (the step kernel sums an atom's pair terms in run order).
// Bond partner rows as NTopoBondAll::build (src/ntopo_bond_all.cpp:39-86), one lane per owner, from the digest.
#pragma once
#include "le_build3.cuh"

#define B4_THREADS 128
#define B4_WARPS (B4_THREADS / 32)
#define B4_WL 768            // work-list entries per warp
#define B4_SVQ 128           // survivor ring (power of two; at most 31 + 64 entries wait in it)
#define B4_WMAX 20           // candidate windows per atom (18 needed: 9 columns, 9 far cells)
#define B4_WLEN_MAX 127      // window length field: 7 bits above the 25 slot bits

struct __align__(16) B4Smem {
  int4 pos[TILE];                        // pos_hold of the tile's atoms
  uint2 sv[B4_SVQ];                      // survivors: x = work-list entry | sliver << 30, y = candidate's w word (tag << 3 | type)
  unsigned hdr[TILE];                    // TopoRec::hdr
  int spec[TOPO_NSPEC][TILE];            // TopoRec::spec, entry-major (lanes that look at different owners hit different banks)
  unsigned win[B4_WMAX][TILE];           // non-empty windows: first slot | length << 25
  unsigned wl[B4_WL];                    // work list: candidate slot | owner lane << 25
  int cnt[TILE];                         // accepted entries per owner
};

template <int UNI>
__global__ void __launch_bounds__(B4_THREADS, 6) k_build4(Dev d) {
  __shared__ B4Smem s_all[B4_WARPS];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  B4Smem &S = s_all[wib];
  const int cap = d.cap;
  Ctrl *__restrict__ ctrl = d.ctrl;
  const int own_end = d.own0 + ctrl->nown;
  const int tile = blockIdx.x * B4_WARPS + wib;
  const int i0 = d.own0 + tile * TILE;
  if (i0 >= own_end) return;                               // the whole warp (= tile) lies beyond the owned atoms
  const int4 *__restrict__ ph = d.pos_hold;
  const bool active = i0 + lane < own_end;
  const int i = active ? i0 + lane : own_end - 1;          // lanes beyond the end shadow the last atom (loads only)
  const int cur = ctrl->cur;
  const int4 pi = ph[i];
  const int tagi = pi.w >> 3, nt = c_P.ntypes;
  const float4 vt = d.vel_tmp[i];
  const int imh = d.img_hold[i];
  const int4 *__restrict__ tr = reinterpret_cast<const int4 *>(d.topo + (tagi - 1));
  const int4 r0 = __ldg(tr), r1 = __ldg(tr + 1), r2 = __ldg(tr + 2), r3 = __ldg(tr + 3);
  S.pos[lane] = pi;
  S.hdr[lane] = (unsigned)r0.x;
  S.spec[0][lane] = r1.z; S.spec[1][lane] = r1.w;
  S.spec[2][lane] = r2.x; S.spec[3][lane] = r2.y; S.spec[4][lane] = r2.z; S.spec[5][lane] = r2.w;
  S.spec[6][lane] = r3.x; S.spec[7][lane] = r3.y; S.spec[8][lane] = r3.z; S.spec[9][lane] = r3.w;
  S.cnt[lane] = 0;
  if (active) {   // the sorted state goes back into the live arrays (the aux word of the velocity follows at the end)
    d.pos[cur][i] = pi;
    d.img[i] = imh;
  }

  // ---- 1. candidate windows ----
  const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
  const int cx = __umulhi((unsigned)pi.x, (unsigned)ncx);
  const int cy = __umulhi((unsigned)pi.y, (unsigned)ncy);
  const int cz = __umulhi((unsigned)pi.z, (unsigned)ncz);
  const int lx = local_layer(d, cx);
  const int zlo = d.cell_abs[2] ? 0 : max(cz - 1, 0), zhi = d.cell_abs[2] ? ncz - 1 : min(cz + 1, ncz - 1);
  const int zwrap = d.cell_abs[2] ? -1 : (cz == 0 ? ncz - 1 : (cz == ncz - 1 ? 0 : -1));
  const int npass = __any_sync(FULL, active && zwrap >= 0) ? 2 : 1;
  int nw = 0, c = 0;
  bool toolong = false;
  for (int pass = 0; pass < npass; pass++) {
    const int za = pass ? zwrap : zlo, zb = pass ? zwrap : zhi;
    const bool on = active && (pass == 0 || zwrap >= 0);
    int wlo[9], whi[9];
#pragma unroll
    for (int ox = 0; ox < 3; ox++) {
      int xc = d.cell_abs[0] ? ox : lx - 1 + ox;
      if (d.nranks == 1) { if (xc < 0) xc += ncx; else if (xc >= ncx) xc -= ncx; }   // one GPU: the slab is the whole box
#pragma unroll
      for (int oy = 0; oy < 3; oy++) {
        wlo[ox * 3 + oy] = whi[ox * 3 + oy] = 0;
        if (ox < d.cell_span[0] && oy < d.cell_span[1] && on) {
          int yc = d.cell_abs[1] ? oy : cy - 1 + oy;
          if (yc < 0) yc += ncy; else if (yc >= ncy) yc -= ncy;
          const int base = cell_slot(d, xc, yc, 0);
          wlo[ox * 3 + oy] = __ldg(&d.cell_start[base + za]); whi[ox * 3 + oy] = __ldg(&d.cell_start[base + zb + 1]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 9; q++) {
      const int len = whi[q] - wlo[q];
      if (len > 0) {
        if (len > B4_WLEN_MAX || nw >= B4_WMAX) toolong = true;
        else { S.win[nw][lane] = (unsigned)wlo[q] | ((unsigned)len << NEIGH_IDX_BITS); nw++; c += len; }
      }
    }
  }
  if (toolong) le_raise(ctrl, LE_DERR_CELL_OVERFLOW, tagi, nw);

  // ---- 2..4: expansion, screen, decide ----
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += v;
  }
  const int off = inc - c;                                  // flat index of this lane's first work item
  const int T = __shfl_sync(FULL, inc, 31);
  const float fsx = c_P.fscale[0], fsy = c_P.fscale[1], fsz = c_P.fscale[2];
  const float hi_u = c_P.cutneigh_hi[0], lo_u = c_P.cutneigh_lo[0];
  const int maxn = min(d.maxneigh, 255);
  unsigned *__restrict__ run = d.nbr + (size_t)tile * d.tcap;
  const unsigned own = (unsigned)lane << NEIGH_IDX_BITS;
  int w = 0, cj = 0, rem = 0;
  if (nw > 0) { const unsigned x = S.win[0][lane]; cj = (int)(x & NEIGH_IDX_MASK); rem = (int)(x >> NEIGH_IDX_BITS); }
  int done = 0;                                             // work items this lane has written so far
  int fill = 0;                                             // items waiting at the front of the work list (< 32)
  int svh = 0, svt = 0;                                     // survivor ring
  int outn = 0;                                             // entries of the tile's run so far
  __syncwarp();

  // 4. one chunk of survivors (n <= 32)
  auto decide_chunk = [&](int n) {
    const bool v = lane < n;
    const uint2 sv = S.sv[(svh + lane) & (B4_SVQ - 1)];
    const int owner = (sv.x >> NEIGH_IDX_BITS) & 31, j = (int)(sv.x & NEIGH_IDX_MASK);
    const int tagj = (int)sv.y >> 3;
    const unsigned hdr = S.hdr[owner];
    const int nscan = (hdr >> 8) & 0xff, n1 = (hdr >> 16) & 0xff, n2 = (hdr >> 24) & 0xff;
    const int nsc = v ? min(nscan, TOPO_NSPEC) : 0;
    const int mx = __reduce_max_sync(FULL, nsc);
    int k = -1;
    for (int q = 0; q < mx; q++)
      if (q < nsc && k < 0 && S.spec[q][owner] == tagj) k = q;
    if (v && k < 0 && nscan > TOPO_NSPEC) {                 // (rare: more specials than the digest holds)
      const int *row = d.special + (size_t)((S.pos[owner].w >> 3) - 1) * d.maxspecial;
      for (int q = TOPO_NSPEC; q < nscan; q++)
        if (row[q] == tagj) { k = q; break; }
    }
    bool acc = v;
    unsigned which = 0;
    if (k >= 0) {
      const int tier = (k < n1) ? 1 : (k < n2) ? 2 : 3;
      const int f = c_P.special_flag[tier];
      if (f == 0) acc = false;
      else if (f != 1) which = (unsigned)tier;
    }
    if (acc && (sv.x & (1u << 30))) {
      const int4 po = S.pos[owner];
      acc = build_border(po, ph[j], UNI ? 0 : (po.w & 7) * nt + ((int)sv.y & 7)) != 0;
    }
    if (acc && atomicAdd(&S.cnt[owner], 1) >= maxn) {
      le_raise(ctrl, LE_DERR_NEIGH_OVERFLOW, S.pos[owner].w >> 3, maxn);
      acc = false;
    }
    const unsigned m = __ballot_sync(FULL, acc);
    if (acc) run[outn + __popc(m & lt_mask)] = (sv.x & ~(1u << 30)) | (which << 30);
    outn += __popc(m);
    svh += n;
  };
  // 3. fp32 screen of one work item; a survivor is appended to the ring
  auto screen = [&](unsigned e, const int4 pj, bool v) {
    const int owner = (e >> NEIGH_IDX_BITS) & 31;
    const int4 po = S.pos[owner];
    const float fx = (float)(int)((unsigned)pj.x - (unsigned)po.x) * fsx;
    const float fy = (float)(int)((unsigned)pj.y - (unsigned)po.y) * fsy;
    const float fz = (float)(int)((unsigned)pj.z - (unsigned)po.z) * fsz;
    const float rsqf = fx * fx + fy * fy + fz * fz;
    const int tp = UNI ? 0 : (po.w & 7) * nt + (pj.w & 7);
    const float hi = UNI ? hi_u : c_P.cutneigh_hi[tp], lo = UNI ? lo_u : c_P.cutneigh_lo[tp];
    const bool s = v && rsqf <= hi && (int)(e & NEIGH_IDX_MASK) != i0 + owner;
    const unsigned m = __ballot_sync(FULL, s);
    if (s) S.sv[(svt + __popc(m & lt_mask)) & (B4_SVQ - 1)] = make_uint2(e | (rsqf >= lo ? (1u << 30) : 0u), (unsigned)pj.w);
    svt += __popc(m);
  };

  for (int base = 0; base < T;) {
    // 2. expansion: flat items [base, lim) go to wl[fill ..)
    const int lim = min(T, base + (B4_WL - fill));
    const int first = off + done;                           // flat index of this lane's next unwritten item (>= base)
    const int m = max(0, min(off + c, lim) - first);
    const int p0 = fill + (first - base);
    const int trips = __reduce_max_sync(FULL, m);
    for (int t = 0; t < trips; t++) {
      if (t < m) {
        S.wl[p0 + t] = (unsigned)cj | own;
        cj++;
        if (--rem == 0 && ++w < nw) { const unsigned x = S.win[w][lane]; cj = (int)(x & NEIGH_IDX_MASK); rem = (int)(x >> NEIGH_IDX_BITS); }
      }
    }
    done += m;
    const int avail = fill + (lim - base);
    base = lim;
    const bool last = base >= T;
    __syncwarp();
    int head = 0;
    while (head < avail && (last || avail - head >= 32)) {
      const int left = avail - head;
      const int n0 = min(32, left);
      const int n1 = (last || left >= 64) ? min(32, left - n0) : 0;
      const bool v0 = lane < n0, v1 = lane < n1;
      const unsigned e0 = v0 ? S.wl[head + lane] : 0u, e1 = v1 ? S.wl[head + 32 + lane] : 0u;
      const int4 q0 = __ldg(&ph[v0 ? (int)(e0 & NEIGH_IDX_MASK) : i]), q1 = __ldg(&ph[v1 ? (int)(e1 & NEIGH_IDX_MASK) : i]);
      screen(e0, q0, v0);
      screen(e1, q1, v1);
      head += n0 + n1;
      __syncwarp();
      while (svt - svh >= 32) { decide_chunk(32); __syncwarp(); }
    }
    // items that do not fill a chunk wait at the front for the next round
    fill = avail - head;
    if (fill > 0) {
      const unsigned e = lane < fill ? S.wl[head + lane] : 0u;
      __syncwarp();
      if (lane < fill) S.wl[lane] = e;
    }
    __syncwarp();
  }
  if (svt > svh) { decide_chunk(svt - svh); __syncwarp(); }

  // ---- bond partner rows (the partners' slots come from the tag map written by k_permute / k_ghost_map) ----
  const int nb = r0.x & 0xff;
  if (active) {
    int n = min(S.cnt[lane], maxn);
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
    if (missing) le_raise(ctrl, LE_DERR_MISSING_ATOM, tagi, nb);
    d.vel[i] = make_float4(vt.x, vt.y, vt.z, __uint_as_float(AUX_PACK(n, nb, 0)));
  }
  if (lane == 0) d.tile_cnt[tile] = (unsigned)outn;
}
