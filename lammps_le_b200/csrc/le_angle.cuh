// le_angle.cuh -- angle_style cosine (E = K (1 + cos theta), src/MOLECULE/angle_cosine.cpp:47-140): chain stiffness of chromatin
// decks.  Not part of the fused step kernel: a side kernel in front of it writes every owned atom's angle
// force into `fang`, the step kernel adds it to the pair + bond force (StepArgs::angles).
//
// Angles are held by tag, replicated like the bond tables: ang[A] = {type, a1, a2 (centre), a3}, and per atom the list of the
// angles it takes part in (ang_cnt / ang_idx, ascending angle id).  One thread per owned atom evaluates ITS angles and keeps the
// share of the force that falls on it -- every angle is computed by its three atoms, no atomics, a fixed summation order, so
// trajectories stay bit-reproducible.  Arms are differences of 32-bit fixed-point coordinates: the closest image for free.
// The energy / virial tally (EV) gives every atom a third of each of its angles (Angle::ev_tally, src/angle.cpp:236-270).
#pragma once
#include "le_common.cuh"

#define LE_ANGLE_NONE 0
#define LE_ANGLE_COSINE 1

template <int EV>
__global__ void __launch_bounds__(256) k_angle(Dev d, int rdp1, int slot) {
  const int cur = rdp1 ? rdp1 - 1 : d.ctrl->cur;
  const int4 *__restrict__ pos = d.pos[cur];
  const int lo = d.own0, hi = d.own0 + (d.nranks > 1 ? d.ctrl->nown : d.N);
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};       // eangle, virial xx yy zz xy xz yz
  for (int k = lo + blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += gridDim.x * blockDim.x) {
    const int t = pos[k].w >> 3;
    const int na = d.ang_cnt[t - 1];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int m = 0; m < na; m++) {
      const int4 a = __ldg(&d.ang[d.ang_idx[(size_t)(t - 1) * d.apa + m]]);
      const int s1 = d.map[a.y - 1], s2 = d.map[a.z - 1], s3 = d.map[a.w - 1];
      if ((s1 | s2 | s3) < 0) { le_raise(d.ctrl, LE_DERR_MISSING_ATOM, t, a.y, a.z, a.w); continue; }
      const int4 p1 = pos[s1], p2 = pos[s2], p3 = pos[s3];
      const double dx1 = (double)(int)((unsigned)p1.x - (unsigned)p2.x) * c_P.scale[0];
      const double dy1 = (double)(int)((unsigned)p1.y - (unsigned)p2.y) * c_P.scale[1];
      const double dz1 = (double)(int)((unsigned)p1.z - (unsigned)p2.z) * c_P.scale[2];
      const double dx2 = (double)(int)((unsigned)p3.x - (unsigned)p2.x) * c_P.scale[0];
      const double dy2 = (double)(int)((unsigned)p3.y - (unsigned)p2.y) * c_P.scale[1];
      const double dz2 = (double)(int)((unsigned)p3.z - (unsigned)p2.z) * c_P.scale[2];
      const double rsq1 = dx1 * dx1 + dy1 * dy1 + dz1 * dz1, rsq2 = dx2 * dx2 + dy2 * dy2 + dz2 * dz2;
      const double r1 = sqrt(rsq1), r2 = sqrt(rsq2);
      double c = (dx1 * dx2 + dy1 * dy2 + dz1 * dz2) / (r1 * r2);
      c = fmin(1.0, fmax(-1.0, c));
      const double kk = c_P.ak_d[a.x - 1];
      const double a11 = kk * c / rsq1, a12 = -kk / (r1 * r2), a22 = kk * c / rsq2;
      const double f1x = a11 * dx1 + a12 * dx2, f1y = a11 * dy1 + a12 * dy2, f1z = a11 * dz1 + a12 * dz2;
      const double f3x = a22 * dx2 + a12 * dx1, f3y = a22 * dy2 + a12 * dy1, f3z = a22 * dz2 + a12 * dz1;
      if (t == a.y) { fx += f1x; fy += f1y; fz += f1z; }
      if (t == a.z) { fx -= f1x + f3x; fy -= f1y + f3y; fz -= f1z + f3z; }
      if (t == a.w) { fx += f3x; fy += f3y; fz += f3z; }
      if (EV) {
        const double w = 1.0 / 3.0;
        acc[0] += w * kk * (1.0 + c);
        acc[1] += w * (dx1 * f1x + dx2 * f3x); acc[2] += w * (dy1 * f1y + dy2 * f3y); acc[3] += w * (dz1 * f1z + dz2 * f3z);
        acc[4] += w * (dx1 * f1y + dx2 * f3y); acc[5] += w * (dx1 * f1z + dx2 * f3z); acc[6] += w * (dy1 * f1z + dy2 * f3z);
      }
    }
    d.fang[3 * (size_t)k] = fx; d.fang[3 * (size_t)k + 1] = fy; d.fang[3 * (size_t)k + 2] = fz;
  }
  if (EV) {
    __shared__ double red[8][7];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 7; q++) { const double s = warp_sum(acc[q]); if (lane == 0) red[w][q] = s; }
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int q = 0; q < 7; q++) {
        double s = lane < 8 ? red[lane][q] : 0.0;
        s = warp_sum(s);
        // thermo slot: 10 = eangle, 3..8 = virial (joins the pair + bond virial)
        if (lane == 0 && s != 0.0) atomicAdd(&d.thermo[(size_t)slot * LE_THERMO_W + (q == 0 ? 10 : 2 + q)], s);
      }
    }
  }
}
