// le_build5.cuh -- rebuild, part 3: neighbor + bond list build (fifth generation).
//
// k_build3 (one lane = one owned atom, nine candidate windows walked in lock step) spent 2600 of its 4170 warp
// instructions per tile in 129 screening trips per lane, although an atom of the 10^6-bead chain has only 28 candidates:
// the trip count was the SUM over the nine windows of the longest window among the 32 lanes.  Two changes:
//   * geometric pruning: a neighbor cell whose closest point is farther from the atom than the list cutoff cannot hold a
//     neighbor.  From the atom's position inside its own cell the kernel drops whole diagonal columns and trims the
//     first / last z cell of a window (corner cells go 48 % of the time, edge cells 21 %): 28 -> ~21 candidates per atom;
//   * one flattened candidate stream per lane: the surviving windows are parked in shared memory and every lane walks
//     its own windows back to back; the warp stays converged (trip count = the largest candidate total among the lanes,
//     lanes that are done are predicated off), so the trip count is max(sum) instead of sum(max).
// Candidate order per atom, pair acceptance (fp32 screen on exact fixed-point differences, the 1e-5 sliver around
// cutneighsq decided by the reference's fp64 arithmetic, npair_half_bin_newton.cpp:98-103), special bits
// (NPair::find_special, src/npair.h:112-136), bond partner rows (NTopoBondAll::build, src/ntopo_bond_all.cpp:39-86) and
// the packed tile run are those of k_build3: the lists are identical entry for entry.
#pragma once
#include "le_build3.cuh"

#define B5_WMAX 18           // windows per atom: 9 columns + 9 far cells of columns that wrap in z

template <int QCAP, int MINB, int UNI>
__global__ void __launch_bounds__(BUILD_THREADS, MINB) k_build5(Dev d) {
  __shared__ int2 s_q[QCAP][BUILD_THREADS];
  __shared__ unsigned s_win[B5_WMAX][BUILD_THREADS];       // first slot | length << 25 (lengths above 127 are split)
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
  if (active) {   // the sorted state goes back into the live arrays (the aux word of the velocity follows at the end)
    d.pos[cur][i] = pi;
    d.img[i] = imh;
  }

  // ---- candidate windows, pruned by the distance from the atom to the neighbor cells ----
  const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
  const unsigned long long mx = (unsigned long long)(unsigned)pi.x * (unsigned)ncx;
  const unsigned long long my = (unsigned long long)(unsigned)pi.y * (unsigned)ncy;
  const unsigned long long mz = (unsigned long long)(unsigned)pi.z * (unsigned)ncz;
  const int cx = (int)(mx >> 32), cy = (int)(my >> 32), cz = (int)(mz >> 32);
  const int lx = local_layer(d, cx);
  // distance to the lower / upper face of the own cell per dimension (0 where every cell of the dimension is visited)
  const float ux = c_P.fscale[0] / (float)ncx, uy = c_P.fscale[1] / (float)ncy, uz = c_P.fscale[2] / (float)ncz;
  // (a dimension with fewer than four cells is not pruned: its lower neighbor cell is also an upper neighbor)
  const bool px = ncx >= 4, py = ncy >= 4, pz = ncz >= 4;
  const float dxl = px ? (float)(unsigned)mx * ux : 0.f, dxh = px ? (float)(~(unsigned)mx) * ux : 0.f;
  const float dyl = py ? (float)(unsigned)my * uy : 0.f, dyh = py ? (float)(~(unsigned)my) * uy : 0.f;
  const float dzl = pz ? (float)(unsigned)mz * uz : 0.f, dzh = pz ? (float)(~(unsigned)mz) * uz : 0.f;
  const float thr = c_P.cutneighmaxsq_f * 1.0005f;       // (fp32 rounding of the face distances: 1e-7 relative)
  const float dx2[3] = {dxl * dxl * 0.9995f, 0.f, dxh * dxh * 0.9995f};
  const float dy2[3] = {dyl * dyl * 0.9995f, 0.f, dyh * dyh * 0.9995f};
  const float dzl2 = dzl * dzl * 0.9995f, dzh2 = dzh * dzh * 0.9995f;
  int nw = 0, ctot = 0;
  bool toolong = false;
  auto push = [&](int lo, int hi) {
    int len = hi - lo;
    while (len > 0) {                                      // (one trip unless a window is longer than 127 slots)
      const int l = min(len, 127);
      if (nw < B5_WMAX) { s_win[nw][t] = (unsigned)lo | ((unsigned)l << NEIGH_IDX_BITS); nw++; ctot += l; }
      else toolong = true;
      lo += l; len -= l;
    }
  };
  if (active) {
#pragma unroll
    for (int ox = 0; ox < 3; ox++) {
      if (ox >= d.cell_span[0]) break;
      int xc = d.cell_abs[0] ? ox : lx - 1 + ox;
      if (d.nranks == 1) { if (xc < 0) xc += ncx; else if (xc >= ncx) xc -= ncx; }   // one GPU: the slab is the whole box
      const float ax = dx2[ox];
#pragma unroll
      for (int oy = 0; oy < 3; oy++) {
        if (oy >= d.cell_span[1]) break;
        int yc = d.cell_abs[1] ? oy : cy - 1 + oy;
        if (yc < 0) yc += ncy; else if (yc >= ncy) yc -= ncy;
        const float axy = ax + dy2[oy];
        if (axy > thr) continue;                           // the whole column is out of reach
        const int base = cell_slot(d, xc, yc, 0);
        if (d.cell_abs[2]) { push(__ldg(&d.cell_start[base]), __ldg(&d.cell_start[base + ncz])); continue; }
        const bool lo_in = axy + dzl2 <= thr, hi_in = axy + dzh2 <= thr;     // the cells below / above cz
        // contiguous part: cells max(cz - 1, 0) .. min(cz + 1, ncz - 1), trimmed
        const int za = (lo_in && cz > 0) ? cz - 1 : cz, zb = (hi_in && cz < ncz - 1) ? cz + 1 : cz;
        push(__ldg(&d.cell_start[base + za]), __ldg(&d.cell_start[base + zb + 1]));
      }
    }
    // far cells of columns that wrap in z (k_build3's second pass: after all contiguous windows, same column order)
    if (!d.cell_abs[2] && (cz == 0 || cz == ncz - 1)) {
      const int zw = cz == 0 ? ncz - 1 : 0;
      const float az = cz == 0 ? dzl2 : dzh2;
#pragma unroll
      for (int ox = 0; ox < 3; ox++) {
        if (ox >= d.cell_span[0]) break;
        int xc = d.cell_abs[0] ? ox : lx - 1 + ox;
        if (d.nranks == 1) { if (xc < 0) xc += ncx; else if (xc >= ncx) xc -= ncx; }
        const float ax = dx2[ox];
#pragma unroll
        for (int oy = 0; oy < 3; oy++) {
          if (oy >= d.cell_span[1]) break;
          int yc = d.cell_abs[1] ? oy : cy - 1 + oy;
          if (yc < 0) yc += ncy; else if (yc >= ncy) yc -= ncy;
          if (ax + dy2[oy] + az > thr) continue;
          const int base = cell_slot(d, xc, yc, 0);
          push(__ldg(&d.cell_start[base + zw]), __ldg(&d.cell_start[base + zw + 1]));
        }
      }
    }
  }
  if (toolong) le_raise(d.ctrl, LE_DERR_CELL_OVERFLOW, tagi, nw);
  __syncwarp();

  const float fsx = c_P.fscale[0], fsy = c_P.fscale[1], fsz = c_P.fscale[2];
  const float hi_u = c_P.cutneigh_hi[0], lo_u = c_P.cutneigh_lo[0];
  SpecCtx S;
  S.nscan = (r0.x >> 8) & 0xff; S.n1 = (r0.x >> 16) & 0xff; S.n2 = (r0.x >> 24) & 0xff;
  S.s0 = r1.z; S.s1 = r1.w; S.s2 = r2.x; S.s3 = r2.y;
  S.rec_spec = tr->spec; S.row = d.special + (size_t)(tagi - 1) * d.maxspecial;
  unsigned *__restrict__ ell = d.nbr_ell + i;          // overflow rows (column i)
  int nq = 0, novf = 0;

  const unsigned BUILD_REJECT = 0xffffffffu;
  auto decide = [&](int j, int tagj, bool sliver, int tj) -> unsigned {
    const int which = find_special3(S, tagj);
    if (which < 0) return BUILD_REJECT;
    if (sliver && !build_border(pi, ph[j], UNI ? 0 : ti * nt + tj)) return BUILD_REJECT;
    return (unsigned)j | ((unsigned)which << 30);
  };
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

  // ---- phase 1: fp32 screen along the lane's flattened stream; two candidates per trip ----
  {
    int w = 0, cj = 0, rem = 0, left = ctot;
    if (nw > 0) { const unsigned x = s_win[0][t]; cj = (int)(x & NEIGH_IDX_MASK); rem = (int)(x >> NEIGH_IDX_BITS); }
    auto next = [&]() -> int {                             // the lane's next candidate slot (valid while left > 0)
      const int j = cj;
      cj++; left--;
      if (--rem == 0 && ++w < nw) { const unsigned x = s_win[w][t]; cj = (int)(x & NEIGH_IDX_MASK); rem = (int)(x >> NEIGH_IDX_BITS); }
      return j;
    };
    const int trips = __reduce_max_sync(FULL, ctot);
    for (int k = 0; k < trips; k += 2) {
      const bool v0 = left > 0;
      const int j0 = v0 ? next() : i;
      const bool v1 = left > 0;
      const int j1 = v1 ? next() : i;
      const int4 p0 = __ldg(&ph[j0]), p1 = __ldg(&ph[j1]);
      screen(p0, j0, v0);
      screen(p1, j1, v1);
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
