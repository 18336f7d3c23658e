// le_sort.cuh -- rebuild, part 1b: the cell sort of the owned atoms (after k_cell_count / k_inbox of le_md.cuh).
//
//   k_scan_cells    exclusive scan of the per-cell counts -> cell_start, ONE launch (4096-cell tiles handed out by
//                   ticket, every tile sums the published sums of its predecessors; the round-1 engine took three
//                   launches and 20 us for it)
//   k_cell_scatter  old slot i -> record {i, tag, cell start, cell end} at cell_start[c] + (arrival rank in the cell)
//   k_permute       record k -> final slot = cell start + rank of its TAG among the cell's members (the local order must
//                   not depend on the order in which the atomics of k_cell_count happened to be served), and the move
//                   of position / velocity / image flags into the new order; refreshes the tag map
// Binning follows NBinStandard::bin_atoms (src/nbin_standard.cpp:192-232) with cells of one neighbor cutoff instead of
// half of one; the reference's linked lists per bin become contiguous slot ranges.
#pragma once
#include "le_common.cuh"

#define SCAN_BLOCK 1024
#define SCAN_MAXITEMS 8                       // rows of 32 cells per warp: chosen by the host so that all tiles run in ONE wave when they can

__device__ __forceinline__ int own_cell_first(const Dev &d) { return cell_slot(d, d.halo, 0, 0); }
__device__ __forceinline__ int own_cell_count(const Dev &d) { return (d.nlx - 2 * d.halo) * d.ncell[1] * d.ncell[2]; }

// block-wide exclusive scan of one value per thread (SCAN_BLOCK threads): warp shuffles + one shared pass
__device__ __forceinline__ int block_excl_scan(int v, int *total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const int base = warp ? wsum[warp - 1] : 0;
  if (total) *total = wsum[31];
  __syncthreads();
  return base + inc - v;
}

// state word of one scan tile: epoch << 32 | the tile's own sum.  The epoch (launches of this kernel so far) makes the
// words of earlier rebuilds read as "not there yet": nothing is reset between rebuilds.
__device__ __forceinline__ unsigned long long ld_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_gpu(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// One launch, one wave: every block (a tile of SCAN_BLOCK * items cells, handed out by ticket so that a block only ever
// waits for tiles that are already running) publishes the sum of its tile, then adds up the sums of ALL its
// predecessors -- one thread per predecessor, one memory round trip -- instead of chaining inclusive prefixes from tile
// to tile.  `items` is sized by the host so that the grid fits the GPU in a single wave (a second wave of a few blocks
// doubled the kernel's duration).
__global__ void __launch_bounds__(SCAN_BLOCK, 2) k_scan_cells(Dev d, int items) {
  __shared__ int s_tile;
  __shared__ unsigned s_epoch;
  __shared__ int s_wtot[32];
  Ctrl *c = d.ctrl;
  if (threadIdx.x == 0) {
    s_epoch = c->nbuilds_scan + 1u;                         // read before this launch's last block bumps it
    s_tile = (int)atomicAdd(&c->scan_ticket, 1u);
  }
  __syncthreads();
  const int tile = s_tile;
  const unsigned epoch = s_epoch;
  const int n = own_cell_count(d), first = own_cell_first(d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // a warp owns 32 * items consecutive cells as `items` rows of 32: every load and store of a row is one coalesced 128-byte
  // line (the first form gave every thread `items` consecutive cells: each of its loads and stores touched 32 lines)
  const int wbase = (tile * SCAN_BLOCK + warp * 32) * items;
  int ex[SCAN_MAXITEMS], wtot = 0;                          // ex[q]: exclusive prefix of this lane's cell of row q inside the warp
#pragma unroll
  for (int q = 0; q < SCAN_MAXITEMS; q++) {
    const int idx = wbase + q * 32 + lane;
    ex[q] = (q < items && idx < n) ? d.cell_count[first + idx] : 0;
  }
#pragma unroll
  for (int q = 0; q < SCAN_MAXITEMS; q++) {
    const int v = ex[q];
    int x = v;
    if (q < items) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
      }
    }
    ex[q] = wtot + x - v;
    wtot += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) s_wtot[warp] = wtot;
  __syncthreads();
  if (warp == 0) {                                          // exclusive scan of the 32 warp totals; lane 31 keeps the tile's sum
    int w = s_wtot[lane], x = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += t;
    }
    s_wtot[lane] = x - w;
    if (lane == 31) st_gpu(&d.scan_state[tile], ((unsigned long long)epoch << 32) | (unsigned)x);
  }
  __syncthreads();
  const int warp_excl = s_wtot[warp];
  // the sums of the tiles before this one
  int before = 0;
  for (int p = threadIdx.x; p < tile; p += SCAN_BLOCK) {
    unsigned long long w;
    do { w = ld_gpu(&d.scan_state[p]); } while ((unsigned)(w >> 32) != epoch);
    before += (int)(unsigned)w;
  }
  int excl;
  block_excl_scan(before, &excl);
  const int base = d.own0 + excl + warp_excl;
#pragma unroll
  for (int q = 0; q < SCAN_MAXITEMS; q++) {
    const int idx = wbase + q * 32 + lane;
    if (q < items && idx < n) {
      d.cell_start[first + idx] = base + ex[q];
      d.cell_count[first + idx] = 0;
    }
  }
  if (tile == (int)gridDim.x - 1 && threadIdx.x == SCAN_BLOCK - 1) {
    // the owned population after migration; the sentinel slot behind the owned region closes its last cell
    const int total = warp_excl + wtot;                     // (warp 31: the tile's sum)
    c->nown = excl + total;
    d.cell_start[first + n] = d.own0 + excl + total;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&c->scan_done, 1u) == gridDim.x - 1) { c->scan_done = 0; c->scan_ticket = 0; c->nbuilds_scan++; }
  }
}

__global__ void k_cell_scatter(Dev d) {
  const int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  const int lo = d.own0, hi = d.own0 + (d.nranks > 1 ? d.ctrl->nown_unsorted : d.N);
  for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
    const int c = d.cellid[i];
    if (c < 0) continue;
    const int s = d.cell_start[c], e = d.cell_start[c + 1];
    d.order2[s + d.slot[i]] = make_int4(i, pos[i].w >> 3, s, e);
  }
}

__global__ void k_permute(Dev d) {
  const int4 *__restrict__ pos = d.pos[d.ctrl->cur];
  const int4 *__restrict__ ord = d.order2;
  const int lo = d.own0, hi = d.own0 + d.ctrl->nown;
  for (int k = lo + blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += gridDim.x * blockDim.x) {
    const int4 r = ord[k];
    const int4 p = pos[r.x];
    const float4 v = d.vel[r.x];
    const int im = d.img[r.x];
    int dest = r.z;
    for (int a = r.z; a < r.w; a++) dest += (ord[a].y < r.y);     // cells hold a handful of atoms
    d.pos_hold[dest] = p;
    d.vel_tmp[dest] = v;
    d.img_hold[dest] = im;
    d.map[r.y - 1] = dest;
  }
}
