// le_build6.cuh -- rebuild, part 3: neighbor + bond list build, tile-centred form (LE_BUILD_VARIANT=6).
//
// k_build3 (one lane = one owned atom, every lane walks ITS 9 windows in lock step with the warp) spends its
// instructions on window bookkeeping: a window of one atom holds 2-3 candidates, the warp-wide trip count is the longest
// of 32 windows, 129 trips for 28 candidates (profiles/r02_build_variants.txt).  Here the warp looks at its tile of 32
// consecutive owned slots as a whole.  The tile's atoms are sorted by cell (x slowest, z fastest), so they are one -- at
// a column end two, rarely more -- SEGMENT: atoms of one (x, y) cell column with z cells zmin..zmax.  The candidates of
// a segment are then 9 long windows (one per neighbor column, z cells zmin-1..zmax+1: ~35 slots each) plus the two
// periodic-wrap layers, 27 windows at most, instead of 9 short windows per atom.  The windows are concatenated into one
// stream (warp scan of their lengths); a round takes 32 consecutive candidates of the stream, one per lane: all lanes
// busy, loads coalesced.  A candidate in z cell c can only pair with the segment's atoms in cells c-1..c+1, and those
// are the contiguous slot range [cell_start[own column, c-1], cell_start[own column, c+2]) clipped to the segment: the
// lane tests exactly these (2-3 atoms, positions staged in shared memory), no search.  Survivors of the fp32 screen are
// appended to the OWNER's queue in shared memory (shared-memory atomic counter); afterwards every lane is the owner of
// its atom again and decides its queue as k_build3 does (find_special on the digest, fp64 sliver check).
// Appending by atomics leaves the queue order to chance; the step kernel adds pair terms in list order and the engine
// is bit-reproducible, so every entry carries the key (window number, slot) = k_build3's visiting order and the
// accepted entries (2 per atom on average) are insertion-sorted by it: the lists are IDENTICAL to k_build3's, word for
// word (tests/test_gpu_step3.py::test_build_variants_give_identical_lists).
// Pair acceptance, special bits and bond rows: as le_build3.cuh (npair_half_bin_newton.cpp:98-103, npair.h:112-136,
// ntopo_bond_all.cpp:39-86).  Needs three or more cells in y and z (and in x on one GPU); the host falls back to k_build3.
#pragma once
#include <type_traits>
#include "le_build3.cuh"

#define B6_THREADS 128
#define B6_WARPS (B6_THREADS / 32)
#define B6_KEY_SHIFT NEIGH_IDX_BITS           // entry while it is being built: slot | window << 25 | which << 30
#define B6_SORT_MASK 0x3fffffffu

template <int QCAP, int MINB, int UNI>
__global__ void __launch_bounds__(B6_THREADS, MINB) k_build6(Dev d, int ell_rows) {
  __shared__ int2 s_q[QCAP][B6_THREADS];       // x: key (window << 25 | slot), y: tag << 1 | "inside the fp64 sliver"
  __shared__ int4 s_pos[B6_THREADS];
  __shared__ int s_cnt[B6_THREADS];
  __shared__ int s_wadj[B6_WARPS][32], s_wend[B6_WARPS][32];
  const unsigned FULL = 0xffffffffu;
  const int cap = d.cap;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, wt = t - lane;
  const int own_end = d.own0 + d.ctrl->nown;
  const int i0 = d.own0 + blockIdx.x * B6_THREADS + t;
  const int tile_lo = i0 - lane;
  if (tile_lo >= own_end) return;                        // the whole warp (= tile) lies beyond the owned atoms
  const int4 *__restrict__ ph = d.pos_hold;
  const int *__restrict__ cstart = d.cell_start;
  const bool active = i0 < own_end;
  const int nact = min(32, own_end - tile_lo);
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
  s_pos[t] = pi;
  s_cnt[t] = 0;

  const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
  const int cy = __umulhi((unsigned)pi.y, (unsigned)ncy);
  const int cz = __umulhi((unsigned)pi.z, (unsigned)ncz);
  const int lx = local_layer(d, __umulhi((unsigned)pi.x, (unsigned)ncx));
  const int colid = lx * ncy + cy;
  if (active) {   // the sorted state goes back into the live arrays (the aux word of the velocity follows at the end)
    d.pos[cur][i] = pi;
    d.img[i] = imh;
  }
  const float fsx = c_P.fscale[0], fsy = c_P.fscale[1], fsz = c_P.fscale[2];
  const float hi_u = c_P.cutneigh_hi[0], lo_u = c_P.cutneigh_lo[0];
  __syncwarp();

  // ---- phase 1: fp32 screen, segment by segment; lanes = candidates of the segment's window stream ----
  for (int s = 0; s < nact;) {
    const int col = __shfl_sync(FULL, colid, s);
    const int len = __popc(__ballot_sync(FULL, colid == col && lane >= s && lane < nact));   // sorted by cell: one run from s
    const int slx = __shfl_sync(FULL, lx, s), scy = __shfl_sync(FULL, cy, s);
    const int zmin = __shfl_sync(FULL, cz, s), zmax = __shfl_sync(FULL, cz, s + len - 1);
    const int seg_lo = tile_lo + s, seg_hi = seg_lo + len;
    // window `lane`: group 0 = the 3 x 3 columns over z cells zmin-1..zmax+1, group 1 = their cell ncz-1 for the atoms in
    // cell 0, group 2 = their cell 0 for the atoms in cell ncz-1; window number = k_build3's visiting order (pass, ox, oy)
    int wl = 0, wn = 0;
    if (lane < 27) {
      const int grp = lane / 9, r = lane - grp * 9, ox = r / 3, oy = r - ox * 3;
      int xc = slx - 1 + ox;
      if (d.nranks == 1) { if (xc < 0) xc += ncx; else if (xc >= ncx) xc -= ncx; }   // one GPU: the slab is the whole box
      int yc = scy - 1 + oy;
      if (yc < 0) yc += ncy; else if (yc >= ncy) yc -= ncy;
      const int base = cell_slot(d, xc, yc, 0);
      const int za = grp == 0 ? max(zmin - 1, 0) : grp == 1 ? ncz - 1 : 0;
      const int zb = grp == 0 ? min(zmax + 1, ncz - 1) : za;
      const bool on = grp == 0 || (grp == 1 ? zmin == 0 : zmax == ncz - 1);
      if (on) { wl = __ldg(&cstart[base + za]); wn = __ldg(&cstart[base + zb + 1]) - wl; }
    }
    int wend = wn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, wend, o);
      if (lane >= o) wend += v;
    }
    const int ctotal = __shfl_sync(FULL, wend, 31);
    s_wadj[warp][lane] = wl - (wend - wn);                 // candidate c of the stream is slot s_wadj[w] + c
    s_wend[warp][lane] = wend;
    __syncwarp();
    const int own_base = cell_slot(d, slx, scy, 0);
    int w = 0;                                             // the window candidate c falls into: only ever moves forward
    for (int c0 = 0; c0 < ctotal; c0 += 32) {
      const int c = c0 + lane;
      const bool live = c < ctotal;
      for (;;) {                                           // (windows are ~35 slots long: one or two passes per round)
        const bool adv = live && s_wend[warp][w] <= c;
        if (!__any_sync(FULL, adv)) break;
        if (adv) w++;
      }
      const int j = live ? s_wadj[warp][w] + c : i;
      const int4 pj = __ldg(&ph[j]);
      const int czj = __umulhi((unsigned)pj.z, (unsigned)ncz);
      // the segment's atoms this candidate can pair with: z cells czj-1..czj+1 (no wrap), or the end cell of a wrap window
      const int za = w < 9 ? max(czj - 1, 0) : w < 18 ? 0 : ncz - 1;
      const int zb = w < 9 ? min(czj + 2, ncz) : za + 1;
      const int ilo = max(__ldg(&cstart[own_base + za]), seg_lo);
      const int ihi = live ? min(__ldg(&cstart[own_base + zb]), seg_hi) : ilo;
      const int trips = __reduce_max_sync(FULL, ihi - ilo);
      const int tj = pj.w & 7;
      // branch-free screen: bit k of `hit` = atom ilo + k passed, of `sliv` = it lies in the fp64 sliver
      unsigned hit = 0, sliv = 0;
      const int rel = wt - tile_lo, last = seg_hi - 1;
      for (int k = 0; k < trips; k++) {
        const int ii = ilo + k;
        const int4 pa = s_pos[rel + min(ii, last)];
        const float fx = (float)(int)((unsigned)pj.x - (unsigned)pa.x) * fsx;
        const float fy = (float)(int)((unsigned)pj.y - (unsigned)pa.y) * fsy;
        const float fz = (float)(int)((unsigned)pj.z - (unsigned)pa.z) * fsz;
        const float rsqf = fx * fx + fy * fy + fz * fz;
        float hi = hi_u, lo = lo_u;
        if (!UNI) { const int tp = (pa.w & 7) * nt + tj; hi = c_P.cutneigh_hi[tp]; lo = c_P.cutneigh_lo[tp]; }
        hit |= (unsigned)(ii < ihi && ii != j && rsqf <= hi) << k;
        sliv |= (unsigned)(rsqf >= lo) << k;
      }
      // survivors go to their owner's queue
      const unsigned key = ((unsigned)w << B6_KEY_SHIFT) | (unsigned)j;
      const int tag2 = (pj.w >> 3) << 1;
      while (__any_sync(FULL, hit != 0)) {
        if (hit) {
          const int k = __ffs(hit) - 1;
          hit &= hit - 1;
          const int ii = ilo + k, il = rel + ii;
          const int q = atomicAdd(&s_cnt[il], 1);
          const int val = tag2 | (int)((sliv >> k) & 1u);
          if (q < QCAP) s_q[q][il] = make_int2((int)key, val);
          else {                                           // queue full: two scratch rows per entry
            const int ro = 2 * (q - QCAP);
            if (ro + 1 < ell_rows) { d.nbr_ell[(size_t)ro * cap + ii] = key; d.nbr_ell[(size_t)(ro + 1) * cap + ii] = (unsigned)val; }
          }
        }
      }
    }
    s += len;
    __syncwarp();
  }

  // ---- phase 2: every lane decides the queue of its own atom; accepted entries are compacted to the front ----
  SpecCtx S;
  S.nscan = (r0.x >> 8) & 0xff; S.n1 = (r0.x >> 16) & 0xff; S.n2 = (r0.x >> 24) & 0xff;
  S.s0 = r1.z; S.s1 = r1.w; S.s2 = r2.x; S.s3 = r2.y;
  S.rec_spec = tr->spec; S.row = d.special + (size_t)(tagi - 1) * d.maxspecial;
  unsigned *__restrict__ ell = d.nbr_ell + i;          // scratch rows (column i)
  int nq = active ? s_cnt[t] : 0;
  if (nq > QCAP + ell_rows / 2) { le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, nq, d.maxneigh); nq = QCAP + ell_rows / 2; }
  // SMALL: no queue of this tile went beyond shared memory (the rule in a dilute system) -> no scratch-row branches
  auto decide_and_sort = [&](auto SMALL) -> int {
    constexpr bool small = decltype(SMALL)::value;
    auto put = [&](int k, unsigned v) { if (small || k < QCAP) s_q[k][t].x = (int)v; else ell[(size_t)(k - QCAP) * cap] = v; };
    auto get = [&](int k) -> unsigned { return (small || k < QCAP) ? (unsigned)s_q[k][t].x : ell[(size_t)(k - QCAP) * cap]; };
    int n = 0;
    const int maxq = __reduce_max_sync(FULL, nq);
    for (int q = 0; q < maxq; q++) {
      if (q < nq) {
        unsigned key; int val;
        if (small || q < QCAP) { const int2 e = s_q[q][t]; key = (unsigned)e.x; val = e.y; }
        else { key = ell[(size_t)(2 * (q - QCAP)) * cap]; val = (int)ell[(size_t)(2 * (q - QCAP) + 1) * cap]; }
        const int j = (int)(key & NEIGH_IDX_MASK);
        const int which = find_special3(S, val >> 1);
        bool ok = which >= 0;
        if (ok && (val & 1)) {
          int tp = 0;
          if (!UNI) tp = ti * nt + (ph[j].w & 7);
          ok = build_border(pi, ph[j], tp);
        }
        if (ok) { put(n, key | ((unsigned)which << 30)); n++; }
      }
    }
    if (n > d.maxneigh) {                                    // neigh_modify one: the tile's run holds 32 * maxneigh entries
      le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, n, d.maxneigh);
      n = d.maxneigh;
    }
    if (n > 255) { le_raise(d.ctrl, LE_DERR_NEIGH_OVERFLOW, tagi, n); n = 255; }
    // k_build3's order: by window, then by slot
    for (int a = 1; a < n; a++) {
      const unsigned v = get(a);
      int b = a - 1;
      while (b >= 0) {
        const unsigned u = get(b);
        if ((u & B6_SORT_MASK) <= (v & B6_SORT_MASK)) break;
        put(b + 1, u);
        b--;
      }
      put(b + 1, v);
    }
    __syncwarp();
    return n;
  };
  const bool all_small = __all_sync(FULL, nq <= QCAP);
  const int n = all_small ? decide_and_sort(std::true_type{}) : decide_and_sort(std::false_type{});

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
  const int tile = (tile_lo - d.own0) >> 5;
  unsigned *__restrict__ run = d.nbr + (size_t)tile * d.tcap;
  const unsigned own = (unsigned)lane << NEIGH_IDX_BITS;
  const unsigned keep = ~(31u << B6_KEY_SHIFT);
  const int maxn = __reduce_max_sync(FULL, n);
  if (all_small) {
    for (int k = 0; k < maxn; k++)
      if (k < n) run[off + k] = ((unsigned)s_q[k][t].x & keep) | own;
  } else {
    for (int k = 0; k < maxn; k++)
      if (k < n) run[off + k] = ((k < QCAP ? (unsigned)s_q[k][t].x : ell[(size_t)(k - QCAP) * cap]) & keep) | own;
  }
  if (lane == 0) d.tile_cnt[tile] = (unsigned)total;
}
