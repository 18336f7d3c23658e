// le_min.cuh -- device helpers of `minimize` (min_style cg with the quadratic line search): the vectors of the conjugate
// gradient iteration live on the device in TAG order (the list rebuilds that every energy evaluation starts with permute
// the local order), the host only steers the iteration with a handful of dot products per force evaluation.
//   reference: MinCG::iterate src/min_cg.cpp:35-200, MinLineSearch::linemin_quadratic src/min_linesearch.cpp:325-505,
//   MinLineSearch::alpha_step :640-690, Min::setup / run / cleanup src/min.cpp:180-520.
// Setup step either side of the hot path, one GPU.
#pragma once
#include "le_common.cuh"

// x0[t], img0[t] <- current position / image flags of tag t+1 (start point of a line search)
__global__ void k_min_save(Dev d, int4 *__restrict__ x0, int *__restrict__ img0) {
  const int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < d.N; k += gridDim.x * blockDim.x) {
    const int4 p = pos[k];
    x0[(p.w >> 3) - 1] = p;
    img0[(p.w >> 3) - 1] = d.img[k];
  }
}

// x <- x0 + alpha h (MinLineSearch::alpha_step); a wrap of the 32-bit coordinate is a periodic crossing (image flags)
__global__ void k_min_move(Dev d, const int4 *__restrict__ x0, const int *__restrict__ img0, const double *__restrict__ h, double alpha) {
  int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < d.N; t += gridDim.x * blockDim.x) {
    const int k = d.map[t];
    const int4 p = x0[t];
    long long du[3];
    for (int q = 0; q < 3; q++) du[q] = llrint(alpha * h[3 * t + q] / c_P.scale[q]);
    const unsigned u[3] = {(unsigned)p.x, (unsigned)p.y, (unsigned)p.z};
    unsigned nu[3]; int w[3];
    for (int q = 0; q < 3; q++) {
      const long long s = (long long)u[q] + du[q];
      w[q] = (int)(s >> 32);                        // floor(s / 2^32): box crossings
      nu[q] = (unsigned)s;
    }
    const int im = img0[t];
    const int ix = (im & 1023) - 512 + w[0], iy = ((im >> 10) & 1023) - 512 + w[1], iz = ((im >> 20) & 1023) - 512 + w[2];
    d.img[k] = ((ix + 512) & 1023) | (((iy + 512) & 1023) << 10) | (((iz + 512) & 1023) << 20);
    pos[k] = make_int4((int)nu[0], (int)nu[1], (int)nu[2], p.w);
  }
}

// out[0] = f.f, out[1] = f.g, out[2] = f.h, out[3] = g.h, out[4] = max |h| (as the bits of a non-negative double),
// out[5] = max |f|
__global__ void k_min_dots(int n3, const double *__restrict__ f, const double *__restrict__ g, const double *__restrict__ h, double *out) {
  double ff = 0, fg = 0, fh = 0, gh = 0, hm = 0, fm = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x) {
    const double a = f[i], b = g[i], c = h[i];
    ff += a * a; fg += a * b; fh += a * c; gh += b * c;
    hm = fmax(hm, fabs(c)); fm = fmax(fm, fabs(a));
  }
  ff = warp_sum(ff); fg = warp_sum(fg); fh = warp_sum(fh); gh = warp_sum(gh);
  for (int o = 16; o > 0; o >>= 1) { hm = fmax(hm, __shfl_xor_sync(0xffffffffu, hm, o)); fm = fmax(fm, __shfl_xor_sync(0xffffffffu, fm, o)); }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], ff); atomicAdd(&out[1], fg); atomicAdd(&out[2], fh); atomicAdd(&out[3], gh);
    atomicMax((unsigned long long *)&out[4], (unsigned long long)__double_as_longlong(hm));
    atomicMax((unsigned long long *)&out[5], (unsigned long long)__double_as_longlong(fm));
  }
}

// g <- f; h <- g + beta h  (beta = 0 with first != 0: h <- g)
__global__ void k_min_dir(int n3, const double *__restrict__ f, double *__restrict__ g, double *__restrict__ h, double beta) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x) {
    const double a = f[i];
    g[i] = a;
    h[i] = a + beta * h[i];
  }
}
__global__ void k_min_copy(int n3, const double *__restrict__ g, double *__restrict__ h) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x) h[i] = g[i];
}
