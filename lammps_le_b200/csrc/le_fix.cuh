// le_fix.cuh -- the three USER-LE fixes as device kernels.
//
// The reference code (src/USER-LE/fix_extrusion.cpp:256-872, fix_ex_load.cpp:329-655,
// fix_ex_unload.cpp:172-372) is a set of SEQUENTIAL loops over the bond list / the half neighbor
// list / the local atoms whose iterations communicate through per-atom scratch (to_add, to_remove,
// distsq, partner) and which consume one sequential Marsaglia stream.  To stay bit-exact yet
// parallel we use three facts:
//   1. the number of draws an iteration consumes depends only on the state BEFORE the loop, so an
//      exclusive scan gives every iteration its offset into the stream, which is generated in bulk
//      (k_ranmars_chunks: the linear lag-97/33 recurrence is jumped ahead chunk by chunk);
//   2. every iteration reads and writes scratch only at a few beads around its extruder (its "touch
//      set"); two iterations whose touch sets are disjoint commute;
//   3. so the loop is run by an ORDERED EXECUTOR: in rounds, every pending iteration claims its
//      beads with atomicMin(priority = position in the sequential order); an iteration that holds
//      all its beads has no earlier pending conflicting iteration and executes.  Extruders are thus
//      resolved in chain-position order with a neighbor-exclusion pass, deterministically.
// Everything here is indexed by tag (t-1); positions are reached through the tag map.
#pragma once
#include "le_common.cuh"
#include "le_build3.cuh"   // topo_pack_one
#include <algorithm>
#include <vector>

#define LE_EXEC_THREADS 1024
#define LE_COPYMAX 192

struct RngDev {
  int s[97];          // ring of the last 97 raw values, 24-bit integers (RanMars::u, src/random_mars.cpp)
  int head;           // index of the oldest value (the one written 97 draws ago)
  int c;              // RanMars::c * 2^24
  long long consumed; // draws handed out so far
};

struct RngHost { int seeded; int seed; };

struct LeFixDev {
  int N;
  int *bondcount, *to_add, *to_remove, *final_add, *final_remove, *partner;
  int *bc_keep;         // fix bond/create: its bond counts between events (bondcount is scratch the other fixes recount)
  double *distsq, *prob;
  unsigned long long *claim;   // ordered executor: (~round << 32) | priority of the best claimant of a bead
  int *exec_rem;               // [LE_EXEC_ROUNDS + 1] "tasks were left over after round r"
  int *flag, *scan, *scan2, *ndraw, *tasks, *blocksum;
  unsigned char *done;
  int *counters;        // [16] device counters
  int *mark_list;       // [2][mark_cap] tags of the end points of the bonds an event broke ([0]) / created ([1])
  int *mark_n;          // [0], [1]: entries of the two lists; [2]: stamp of the current topology sweep
  int *infl_stamp;      // [N] last sweep that looked at the atom (every candidate is tested once per sweep)
  int mark_cap;
  double *draws;
  int draws_cap;
  int *rm_base;         // [193] raw values ahead of the current state (k_ranmars_base)
  unsigned *rm_jump;    // [(nchunks-1)][97] jump polynomials
  RngDev *rngdev;       // [3]
  RngHost rng[3];
  void *scratch64;
  double *geo;          // [N][LE_GEO_D] position-dependent inputs of the current event, by tag (in the peer arena)
  int *geo_i;           // [N][LE_GEO_I]
};

struct ExtrusionArgs { int btype, neutral, left, right, lr; double p; };
struct LoadArgs { int btype, itype, jtype, imax, inew, jmax, jnew; double cutsq, fraction; };
struct UnloadArgs { int btype; double cutsq, fraction; };

enum { CNT_NTASK = 0, CNT_NDRAW = 1, CNT_NBREAK = 2, CNT_NCREATE = 3, CNT_TOTAL = 4, CNT_NLIST = 5 };

static int le_fix_alloc(LeFixDev &f, int n, int maxspecial, std::vector<void *> &allocs, cudaStream_t st) {
  f.N = n;
  auto A = [&](void **p, size_t bytes) -> int {
    if (cudaMalloc(p, bytes) != cudaSuccess) return -1;
    cudaMemsetAsync(*p, 0, bytes, st);
    allocs.push_back(*p);
    return 0;
  };
  const size_t n1 = (size_t)n + 2;
  int r = 0;
  r |= A((void **)&f.bondcount, n1 * 4); r |= A((void **)&f.bc_keep, n1 * 4); r |= A((void **)&f.to_add, n1 * 4); r |= A((void **)&f.to_remove, n1 * 4);
  r |= A((void **)&f.final_add, n1 * 4); r |= A((void **)&f.final_remove, n1 * 4); r |= A((void **)&f.partner, n1 * 4);
  r |= A((void **)&f.distsq, n1 * 8); r |= A((void **)&f.prob, n1 * 8);
  r |= A((void **)&f.claim, n1 * 8); r |= A((void **)&f.exec_rem, 64 * 4); r |= A((void **)&f.flag, n1 * 4); r |= A((void **)&f.scan, n1 * 4);
  r |= A((void **)&f.scan2, n1 * 4); r |= A((void **)&f.ndraw, n1 * 4);
  r |= A((void **)&f.tasks, n1 * 4); r |= A((void **)&f.blocksum, (n1 / 1024 + 2) * 4);
  r |= A((void **)&f.done, n1);
  r |= A((void **)&f.counters, 16 * 4);
  f.mark_cap = 2 * n + 16;
  r |= A((void **)&f.mark_list, (size_t)2 * f.mark_cap * 4); r |= A((void **)&f.mark_n, 4 * 4); r |= A((void **)&f.infl_stamp, n1 * 4);
  f.draws_cap = 2 * n + 64;
  r |= A((void **)&f.draws, (size_t)f.draws_cap * 8);
  r |= A((void **)&f.rngdev, 3 * sizeof(RngDev));

  r |= A(&f.scratch64, 64);
  (void)maxspecial;
  return r;
}

// ------------------------------------------------------------------------------------------------
// RanMars (src/random_mars.cpp:29-95) in exact 24-bit integer arithmetic, generated in parallel.
//   raw_n = (raw_{n-97} - raw_{n-33}) mod 2^24;  c_n = c_{n-1} - cd (+cm if negative);  out = (raw_n - c_n) mod 2^24
//   every value is a multiple of 2^-24, so the reference's doubles are reproduced exactly.
// The raw recurrence is linear over Z/2^24 with characteristic polynomial x^97 + x^64 - 1, so the state K
// draws ahead is a fixed linear map of the current one: x^(wK) mod p(x) (97 coefficients, precomputed on the host
// for every chunk w) applied to 193 consecutive raw values.  Warp w jumps to draw w*K and produces K draws, 32
// per step (the shortest lag is 33); c_n is an arithmetic progression mod cm and needs no jump table.
// ------------------------------------------------------------------------------------------------
#define RM_CHUNK 1024
#define RM_CD 7654321LL
#define RM_CM 16777213LL

// x^k mod (x^97 + x^64 - 1) over Z/2^24 for k = K, 2K, ..., (nchunks-1)K; row w-1 holds chunk w
static void ranmars_jump_table(int nchunks, std::vector<unsigned> &table) {
  const unsigned M = 0xffffffu;
  auto mulmod = [&](const std::vector<unsigned> &a, const std::vector<unsigned> &b) {
    std::vector<unsigned long long> t(193, 0);
    for (int i = 0; i < 97; i++) {
      if (!a[i]) continue;
      for (int j = 0; j < 97; j++) t[i + j] = (t[i + j] + (unsigned long long)a[i] * b[j]) & M;
    }
    for (int dgr = 192; dgr >= 97; dgr--) {          // x^d = x^(d-97) * (1 - x^64)
      const unsigned long long cf = t[dgr];
      if (!cf) continue;
      t[dgr - 97] = (t[dgr - 97] + cf) & M;
      t[dgr - 33] = (t[dgr - 33] + (16777216ULL - cf)) & M;
      t[dgr] = 0;
    }
    std::vector<unsigned> r(97);
    for (int i = 0; i < 97; i++) r[i] = (unsigned)t[i];
    return r;
  };
  std::vector<unsigned> xk(97, 0), base(97, 0);
  base[1] = 1;                                        // x
  xk[0] = 1;
  for (int bit = 0, k = RM_CHUNK; k; k >>= 1, bit++) { if (k & 1) xk = mulmod(xk, base); base = mulmod(base, base); }
  table.assign((size_t)std::max(nchunks - 1, 1) * 97, 0);
  std::vector<unsigned> cur = xk;
  for (int w = 1; w < nchunks; w++) {
    for (int i = 0; i < 97; i++) table[(size_t)(w - 1) * 97 + i] = cur[i];
    cur = mulmod(cur, xk);
  }
}

static int le_fix_alloc_rng(LeFixDev &f, std::vector<void *> &allocs, cudaStream_t st) {
  const int nchunks = (f.draws_cap + RM_CHUNK - 1) / RM_CHUNK;
  std::vector<unsigned> table;
  ranmars_jump_table(nchunks, table);
  if (cudaMalloc((void **)&f.rm_base, 256 * 4) != cudaSuccess) return -1;
  allocs.push_back(f.rm_base);
  if (cudaMalloc((void **)&f.rm_jump, table.size() * 4) != cudaSuccess) return -1;
  allocs.push_back(f.rm_jump);
  cudaMemcpyAsync(f.rm_jump, table.data(), table.size() * 4, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  return 0;
}

// raw values 0..192 relative to the current state (the state is raw 0..96), into `base`
__global__ void k_ranmars_base(const RngDev *st, int *base, const int *n_ptr, int cap, Ctrl *ctrl) {
  const int lane = threadIdx.x;
  const int n = n_ptr ? *n_ptr : 0;
  if (n <= 0) return;
  if (n > cap) { if (lane == 0) le_raise(ctrl, LE_DERR_RNG_OVERFLOW, n, cap); return; }
  __shared__ int ring[224];
  for (int k = lane; k < 97; k += 32) ring[k] = st->s[(st->head + k) % 97];
  __syncwarp();
  for (int w = 97; w < 193; w += 32) {
    int raw = ring[w + lane - 97] - ring[w + lane - 33];
    if (raw < 0) raw += 16777216;
    __syncwarp();
    ring[w + lane] = raw;
    __syncwarp();
  }
  for (int k = lane; k < 193; k += 32) base[k] = ring[k];
  if (lane == 0) base[200] = st->c;   // the chunk warps must not read st: the last chunk rewrites it
}

// one warp per chunk of RM_CHUNK draws
__global__ void __launch_bounds__(128) k_ranmars_chunks(RngDev *st, const int *__restrict__ base, const unsigned *__restrict__ jump,
                                                        double *out, const int *n_ptr, int cap) {
  __shared__ int rings[4][256];
  __shared__ int sbase[193];
  const int n = n_ptr ? *n_ptr : 0;
  if (n <= 0 || n > cap) return;
  const int nchunks = (n + RM_CHUNK - 1) / RM_CHUNK;
  if ((int)blockIdx.x * 4 >= nchunks) return;
  for (int k = threadIdx.x; k < 193; k += blockDim.x) sbase[k] = base[k];
  __syncthreads();
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int w = blockIdx.x * 4 + wl;
  if (w >= nchunks) return;
  int *ring = rings[wl];
  // start state of this chunk: raw_{wK+i} = sum_j a_j raw_{i+j}
  if (w == 0) {
    for (int i = lane; i < 97; i += 32) ring[i] = sbase[i];
  } else {
    const unsigned *a = jump + (size_t)(w - 1) * 97;
    for (int i = lane; i < 97; i += 32) {
      unsigned long long acc = 0;
      for (int j = 0; j < 97; j++) acc += (unsigned long long)__ldg(&a[j]) * (unsigned)sbase[i + j];
      ring[i] = (int)(acc & 0xffffffu);
    }
  }
  __syncwarp();
  const long long first = (long long)w * RM_CHUNK;          // draws before this chunk
  const long long cst = base[200];
  long long c0 = (cst - (first % RM_CM) * RM_CD % RM_CM) % RM_CM;
  if (c0 < 0) c0 += RM_CM;
  const int count = min(RM_CHUNK, n - (int)first);
  int wp = 97;
  for (int b = 0; b < count; b += 32) {
    int raw = ring[(wp + lane - 97) & 255] - ring[(wp + lane - 33) & 255];
    if (raw < 0) raw += 16777216;
    long long cc = (c0 - ((long long)(b + lane + 1) * RM_CD) % RM_CM) % RM_CM;
    if (cc < 0) cc += RM_CM;
    int v = raw - (int)cc;
    if (v < 0) v += 16777216;
    __syncwarp();
    ring[(wp + lane) & 255] = raw;
    if (b + lane < count) out[first + b + lane] = (double)v * (1.0 / 16777216.0);
    __syncwarp();
    wp += min(32, count - b);
  }
  // the chunk that holds the last draw leaves the new generator state: raw_n .. raw_{n+96}
  if (w == nchunks - 1) {
    __syncwarp();
    long long cn = (cst - ((long long)n % RM_CM) * RM_CD % RM_CM) % RM_CM;
    if (cn < 0) cn += RM_CM;
    __syncwarp();
    for (int k = lane; k < 97; k += 32) st->s[k] = ring[(wp - 97 + k) & 255];
    if (lane == 0) { st->head = 0; st->c = (int)cn; st->consumed += n; }
  }
}

// ------------------------------------------------------------------------------------------------
// generic exclusive scan over n ints (three launches) and flag compaction
// ------------------------------------------------------------------------------------------------
__global__ void k_iscan_partial(const int *in, int n, int *blocksum) {
  __shared__ int sh[32];
  const int idx = blockIdx.x * 1024 + threadIdx.x;
  int s = (idx < n) ? in[idx] : 0;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = sh[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) k_iscan_blocks(int *blocksum, int nblocks, int *total) {
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int idx = base + threadIdx.x;
    const int v = (idx < nblocks) ? blocksum[idx] : 0;
    int tot;
    const int ex = block_excl_scan(v, &tot);
    if (idx < nblocks) blocksum[idx] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void __launch_bounds__(1024) k_iscan_apply(const int *in, int *out, int n, const int *blocksum) {
  const int idx = blockIdx.x * 1024 + threadIdx.x;
  const int v = (idx < n) ? in[idx] : 0;
  const int ex = block_excl_scan(v, nullptr);
  if (idx < n) out[idx] = blocksum[blockIdx.x] + ex;
}
// tasks[scan[i]] = i for flagged i (flag value > 0)
__global__ void k_compact(const int *flag, const int *scan, int n, int *tasks) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (flag[i] > 0) tasks[scan[i]] = i;
}

// ------------------------------------------------------------------------------------------------
// helpers: tag-indexed access to the live state
// ------------------------------------------------------------------------------------------------
struct LeView {
  Dev d; LeFixDev f;
  __device__ __forceinline__ int cur() const { return d.ctrl->cur; }   // position buffer holding the current coordinates
};

// the coordinate the reference holds for an OWNED atom between reneighborings: wrapped at the last
// rebuild (Domain::pbc runs only then, src/verlet.cpp:272), drifting freely since.  The number of box
// crossings since the rebuild is the wrap of the 32-bit coordinate relative to pos_hold, which is also
// available for ghosts.  Returns false if the atom is not on this GPU (halo too thin).
__device__ __forceinline__ bool raw_xyz(const LeView &V, int tag, double x[3]) {
  const int k = V.d.map[tag - 1];
  if (k < 0) { le_raise(V.d.ctrl, LE_DERR_MISSING_ATOM, tag, 0, 1); x[0] = x[1] = x[2] = 0.0; return false; }
  const int4 p = V.d.pos[V.cur()][k], h = V.d.pos_hold[k];
  const unsigned u[3] = {(unsigned)p.x, (unsigned)p.y, (unsigned)p.z};
  const unsigned uh[3] = {(unsigned)h.x, (unsigned)h.y, (unsigned)h.z};
#pragma unroll
  for (int q = 0; q < 3; q++) {
    double v = le_deq(u[q], q);
    const int di = le_image_shift(uh[q], u[q]);
    if (di) v = __dadd_rn(v, (double)di * c_P.L[q]);
    x[q] = v;
  }
  return true;
}
// image of bond partner p closest to atom t at the last rebuild, as shifts -1/0/+1 per dimension packed 2 bits
// each (+1 bias; 21 = same image): Domain::closest_image picks a ghost exactly when a shift is non-zero, and then
// the reference's bondlist holds the bond twice (src/ntopo_bond_all.cpp:65-66).  0 = partner not on this GPU.
__device__ __forceinline__ int bond_cross_code(const Dev &d, int t, int p) {
  const int kt = d.map[t - 1], kp = d.map[p - 1];
  if (kt < 0 || kp < 0) { le_raise(d.ctrl, LE_DERR_MISSING_ATOM, t, p, 2); return 0; }
  const int4 a = d.pos_hold[kt], b = d.pos_hold[kp];
  const int h0 = le_image_shift((unsigned)a.x, (unsigned)b.x);
  const int h1 = le_image_shift((unsigned)a.y, (unsigned)b.y);
  const int h2 = le_image_shift((unsigned)a.z, (unsigned)b.z);
  return (h0 + 1) | ((h1 + 1) << 2) | ((h2 + 1) << 4);
}
__device__ __forceinline__ double dist2(const double a[3], const double b[3]) {
  const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ int type_of(const LeView &V, int tag) { return V.d.type_tag[tag - 1]; }

// ------------------------------------------------------------------------------------------------
// position-dependent inputs of an event ("geometry records").  The decision logic of the three fixes is
// replicated: every GPU runs it on the whole (small) set of extruders with identical inputs, so the
// result is independent of the decomposition and bit-identical to the 1-rank reference.  What a GPU
// cannot do alone is measure distances between beads it does not hold; so the owner of a bead computes
// the few distances / image codes its extruder needs and stores them, by tag, into EVERY GPU's record
// arrays (peer stores over NVLink), stamped with the event number.  One flag round (k_le_exchange)
// later all GPUs hold all records.
//   geo_i[t][0]  extrusion: bond_cross_code of t's extruder bond | unload: partner tag | load: eligibility flag
//   geo_i[t][1]  stamp (Ctrl::le_epoch of the event that wrote the record; anything else = no record)
//   geo[t][0..2] extrusion (t = lower end a): |L-R|^2, |L-b|^2, |a-R|^2 | load: |lo-hi|^2 in [0]
// ------------------------------------------------------------------------------------------------
__global__ void k_le_begin(Dev d) { d.ctrl->le_epoch++; }

__device__ __forceinline__ void geo_store(const Dev &d, int i, int v0, int stamp, int nd, const double *r) {
  for (int p = 0; p < d.nranks; p++) {
    const PeerView &pv = d.peer[p];
    for (int q = 0; q < nd; q++) pv.geo[(size_t)i * LE_GEO_D + q] = r[q];
    pv.geo_i[(size_t)i * LE_GEO_I] = v0;
    pv.geo_i[(size_t)i * LE_GEO_I + 1] = stamp;
  }
}
__device__ __forceinline__ bool geo_fresh(const LeView &V, int i) {
  return V.f.geo_i[(size_t)i * LE_GEO_I + 1] == (int)V.d.ctrl->le_epoch;
}

// all records of this event are out: tell every peer, wait for every peer (one thread)
__global__ void k_le_exchange(Dev d) {
  if (d.nranks <= 1) return;
  Ctrl *c = d.ctrl;
  __threadfence_system();
  const unsigned long long e = (unsigned long long)c->le_epoch;
  const int slot = FLAG_LE + (int)(e & 1) * LE_MAXRANKS;
  for (int p = 0; p < d.nranks; p++)
    if (p != d.rank) st_sys(&d.peer[p].flags[slot + d.rank], e);
  for (int p = 0; p < d.nranks; p++)
    if (p != d.rank) le_wait_flag(c, &d.flags[slot + p], e, 0);
  __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// ordered executor (one block).  Task must provide:
//   int  ntasks;  unsigned prio(int k);  int beads(int k, int b[4]) (tag-1 indices, may repeat);
//   void exec(int k)
// ------------------------------------------------------------------------------------------------
// Rounds 0..LE_EXEC_ROUNDS-1 run as two grid-wide kernels each (claim, run); a claim is the 64-bit word
// (~round << 32) | priority, so a later round's claim beats every stale one and nothing has to be reset between
// rounds.  Whatever is still pending after the last grid round (a conflict chain longer than LE_EXEC_ROUNDS
// extruders) is finished by one block that loops to completion.
#define LE_EXEC_ROUNDS 6
#define LE_EXEC_BASE_STRIDE (1u << 20)

__device__ __forceinline__ unsigned long long exec_key(unsigned round, unsigned prio) {
  return ((unsigned long long)(0xffffffffu - round) << 32) | prio;
}

template <class Task>
__global__ void k_exec_claim(Task T, unsigned long long *claim, unsigned char *done, const int *rem, int round, unsigned base) {
  if (round > 0 && rem[round - 1] == 0) return;
  const int n = T.ntasks();
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    if (round == 0) done[k] = 0;
    else if (done[k]) continue;
    int b[4];
    const int nb = T.beads(k, b);
    const unsigned long long key = exec_key(base + round, T.prio(k));
    for (int q = 0; q < nb; q++) atomicMin(&claim[b[q]], key);
  }
}

template <class Task>
__global__ void k_exec_run(Task T, unsigned long long *claim, unsigned char *done, int *rem, int round, unsigned base) {
  if (round > 0 && rem[round - 1] == 0) return;
  const int n = T.ntasks();
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    if (done[k]) continue;
    int b[4];
    const int nb = T.beads(k, b);
    const unsigned long long key = exec_key(base + round, T.prio(k));
    bool mine = true;
    for (int q = 0; q < nb; q++) mine = mine && (claim[b[q]] == key);
    if (mine) { T.exec(k); done[k] = 1; }
    else rem[round] = 1;
  }
}

template <class Task>
__global__ void __launch_bounds__(LE_EXEC_THREADS) k_exec_finish(Task T, unsigned long long *claim, unsigned char *done, const int *rem, unsigned base) {
  if (rem[LE_EXEC_ROUNDS - 1] == 0) return;
  __shared__ int remaining;
  const int n = T.ntasks();
  for (unsigned round = LE_EXEC_ROUNDS; round < LE_EXEC_BASE_STRIDE; round++) {
    if (threadIdx.x == 0) remaining = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      if (done[k]) continue;
      int b[4];
      const int nb = T.beads(k, b);
      const unsigned long long key = exec_key(base + round, T.prio(k));
      for (int q = 0; q < nb; q++) atomicMin(&claim[b[q]], key);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      if (done[k]) continue;
      int b[4];
      const int nb = T.beads(k, b);
      const unsigned long long key = exec_key(base + round, T.prio(k));
      bool mine = true;
      for (int q = 0; q < nb; q++) mine = mine && (claim[b[q]] == key);
      if (mine) { T.exec(k); done[k] = 1; }
      else remaining = 1;
    }
    __syncthreads();
    if (!remaining) break;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// shared pieces: bond count, slot deletion, special-list edits (identical in the three fixes)
// ------------------------------------------------------------------------------------------------
// delete the first slot of atom t whose partner is p (fix_extrusion.cpp:656-668, fix_ex_unload.cpp:293-302)
__device__ __forceinline__ void delete_bond_slot(const Dev &d, int t, int p) {
  int *ba = d.bond_atom + (size_t)(t - 1) * d.bpa, *bt = d.bond_type + (size_t)(t - 1) * d.bpa;
  const int nb = d.num_bond[t - 1];
  for (int m = 0; m < nb; m++) {
    if (ba[m] == p) {
      for (int k = m; k < nb - 1; k++) { ba[k] = ba[k + 1]; bt[k] = bt[k + 1]; }
      d.num_bond[t - 1] = nb - 1;
      break;
    }
  }
}
// remove p from the 1-2 part of t's special list (fix_extrusion.cpp:673-683)
__device__ __forceinline__ void special_remove12(const Dev &d, int t, int p) {
  int *sl = d.special + (size_t)(t - 1) * d.maxspecial;
  int *ns = d.nspecial + (size_t)(t - 1) * 3;
  const int n1 = ns[0], n3 = ns[2];
  int m;
  for (m = 0; m < n1; m++) if (sl[m] == p) break;
  for (; m < n3 - 1; m++) sl[m] = sl[m + 1];
  ns[0]--; ns[1]--; ns[2]--;
}
// make p a 1-2 neighbor of t (fix_extrusion.cpp:748-771, fix_ex_load.cpp:570-588)
__device__ __forceinline__ void special_insert12(const Dev &d, int t, int p) {
  int *sl = d.special + (size_t)(t - 1) * d.maxspecial;
  int *ns = d.nspecial + (size_t)(t - 1) * 3;
  int n1 = ns[0], n2 = ns[1], n3 = ns[2];
  int m;
  for (m = n1; m < n3; m++) if (sl[m] == p) break;
  if (m < n3) {
    for (int n = m; n < n3 - 1; n++) sl[n] = sl[n + 1];
    n3--;
    if (m < n2) n2--;
  }
  if (n3 == d.maxspecial) { le_raise(d.ctrl, LE_DERR_SPECIAL_OVERFLOW, t, p); return; }
  for (m = n3; m > n1; m--) sl[m] = sl[m - 1];
  sl[n1] = p;
  ns[0] = n1 + 1; ns[1] = n2 + 1; ns[2] = n3 + 1;
}

// FixExtrusion::dedup (fix_extrusion.cpp:1116-1135)
__device__ __forceinline__ int le_dedup(int nstart, int nstop, int *copy) {
  int m = nstart;
  while (m < nstop) {
    int i;
    for (i = 0; i < m; i++)
      if (copy[i] == copy[m]) { copy[m] = copy[nstop - 1]; nstop--; break; }
    if (i == m) m++;
  }
  return nstop;
}

// rebuild_special_one (fix_extrusion.cpp:1045-1108): 1-3 = 1-2 of 1-2, 1-4 = 1-2 of 1-3
__device__ void rebuild_special_one(const Dev &d, int t) {
  int copy[LE_COPYMAX];
  int *sl = d.special + (size_t)(t - 1) * d.maxspecial;
  int *ns = d.nspecial + (size_t)(t - 1) * 3;
  const int cn1 = ns[0];
  for (int i = 0; i < cn1; i++) copy[i] = sl[i];
  int cn2 = cn1;
  for (int i = 0; i < cn1; i++) {
    const int n = copy[i];
    const int *s2 = d.special + (size_t)(n - 1) * d.maxspecial;
    const int m1 = d.nspecial[(size_t)(n - 1) * 3];
    for (int j = 0; j < m1; j++)
      if (s2[j] != t) { if (cn2 >= LE_COPYMAX) { le_raise(d.ctrl, LE_DERR_SPECIAL_OVERFLOW, t, cn2); return; } copy[cn2++] = s2[j]; }
  }
  cn2 = le_dedup(cn1, cn2, copy);
  if (cn2 > d.maxspecial) { le_raise(d.ctrl, LE_DERR_SPECIAL_OVERFLOW, t, cn2); return; }
  int cn3 = cn2;
  for (int i = cn1; i < cn2; i++) {
    const int n = copy[i];
    const int *s2 = d.special + (size_t)(n - 1) * d.maxspecial;
    const int m1 = d.nspecial[(size_t)(n - 1) * 3];
    for (int j = 0; j < m1; j++)
      if (s2[j] != t) { if (cn3 >= LE_COPYMAX) { le_raise(d.ctrl, LE_DERR_SPECIAL_OVERFLOW, t, cn3); return; } copy[cn3++] = s2[j]; }
  }
  cn3 = le_dedup(cn2, cn3, copy);
  if (cn3 > d.maxspecial) { le_raise(d.ctrl, LE_DERR_SPECIAL_OVERFLOW, t, cn3); return; }
  // the 1-2 part is unchanged (and is being read by other threads): write only tiers 2 and 3
  ns[1] = cn2; ns[2] = cn3;
  for (int i = cn1; i < cn3; i++) sl[i] = copy[i];
}

// an end point of a broken (which 0) / created (which 1) bond joins the event's list of marked atoms
__device__ __forceinline__ void le_mark(const LeFixDev &f, int which, int tag) {
  const int q = atomicAdd(&f.mark_n[which], 1);
  if (q < f.mark_cap) f.mark_list[(size_t)which * f.mark_cap + q] = tag;
}

// update_topology sweeps (fix_extrusion.cpp:924-1002, fix_ex_load.cpp:700-751, fix_ex_unload.cpp:417-463).
// mode 0: broken bonds, marks = final_remove: influenced if endpoint, or BOTH ends of one broken bond are in
//         the full special list;  mode 1: created bonds, marks = final_add: endpoint, or EITHER end among 1-2/1-3.
// Two kernels: k_le_topo_detect collects the influenced atoms in a list, then rebuild_special_one runs over that list (the
// rebuilds commute: each writes only its own atom's 1-3/1-4 tiers and reads only 1-2 tiers, which no rebuild
// touches).
// the influence test of one atom (tag i + 1), exactly as the reference's sweeps state it
__device__ __forceinline__ bool le_influenced(const Dev &d, const int *marks, int mode, int i) {
  if (marks[i] != 0) return true;
  const int *sl = d.special + (size_t)i * d.maxspecial;
  if (mode == 0) {
    const int n = d.nspecial[(size_t)i * 3 + 2];
    for (int k = 0; k < n; k++) {
      const int s = sl[k];
      const int p = marks[s - 1];
      if (p == 0) continue;
      int found = 0;
      for (int q = 0; q < n; q++) if (sl[q] == s || sl[q] == p) found++;
      if (found == 2) return true;
    }
  } else {
    const int n = d.nspecial[(size_t)i * 3 + 1];
    for (int k = 0; k < n; k++) if (marks[sl[k] - 1] != 0) return true;
  }
  return false;
}
// Detection without a sweep over all tags: an atom can only be influenced if a marked atom (an end point of a broken /
// created bond) is on its special list, and special lists are mutual -- so the candidates are the marked atoms, the atoms
// on their special lists and (a margin for lists that an earlier sweep of the same event has already rebuilt on one side
// only) the atoms on THOSE lists.  Every candidate is put to the reference's own test once (stamp per sweep).  At the
// 10^6-bead bench an event marks ~10^3 atoms: 10^5 tests instead of a pass over N (x GPUs, the logic is replicated).
__global__ void k_le_topo_detect(Dev d, LeFixDev f, const int *marks, int mode, const int *gate, int *list, int *nlist) {
  if (gate && *gate == 0) return;
  const int nm = min(f.mark_n[mode], f.mark_cap), stamp = f.mark_n[2];
  const int *ml = f.mark_list + (size_t)mode * f.mark_cap;
  auto visit = [&](int tag) {
    if (atomicExch(&f.infl_stamp[tag - 1], stamp) == stamp) return;      // somebody has tested this atom in this sweep
    if (le_influenced(d, marks, mode, tag - 1)) list[atomicAdd(nlist, 1)] = tag;
  };
  // one thread per (marked atom, entry of its special list): the thread visits that atom and walks ITS special list
  const int per = d.maxspecial + 1;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < nm * per; w += gridDim.x * blockDim.x) {
    const int t = ml[w / per], a = w % per - 1;
    if (a < 0) { visit(t); continue; }
    if (a >= d.nspecial[(size_t)(t - 1) * 3 + 2]) continue;
    const int u = d.special[(size_t)(t - 1) * d.maxspecial + a];
    visit(u);
    const int *s2 = d.special + (size_t)(u - 1) * d.maxspecial;
    const int n2 = d.nspecial[(size_t)(u - 1) * 3 + 2];
    for (int b = 0; b < n2; b++) visit(s2[b]);
  }
}
__global__ void k_le_topo_rebuild(Dev d, const int *list, int *nlist) {
  const int n = *nlist;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    rebuild_special_one(d, list[k]);
    // the atom's digest for the list build follows its tables at once: every atom whose bond or special tables an event
    // changed is on one of the event's lists (the end points of a broken / created bond carry a mark themselves), so no
    // sweep over all tags is needed after an event
    topo_pack_one(d, list[k] - 1);
  }
}
__global__ void k_le_topo_reset(int *nlist, int *mark_n) { *nlist = 0; mark_n[2]++; }

// ------------------------------------------------------------------------------------------------
// fix extrusion
// ------------------------------------------------------------------------------------------------
// E0a: recount bondcount, clear scratch (fix_extrusion.cpp:281-295,339-348)
__global__ void k_ext_init(LeView V, int btype) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    int bc = 0;
    const int nb = d.num_bond[i];
    for (int m = 0; m < nb; m++) if (d.bond_type[(size_t)i * d.bpa + m] == btype) bc++;
    if (bc > 1) le_raise(d.ctrl, LE_DERR_BONDCOUNT, i + 1, bc);
    f.bondcount[i] = bc;
    f.to_add[i] = 0; f.to_remove[i] = 0; f.final_add[i] = 0; f.final_remove[i] = 0;
    f.distsq[i] = LE_BIG; f.claim[i] = ~0ull; f.partner[i] = 0;
    if (i < 16) f.counters[i] = 0;
    if (i < 2) f.mark_n[i] = 0;
  }
}

// side test of the candidate pass; returns whether the clauses before the random draws hold and how many
// draws the reference's short-circuit chain consumes (fix_extrusion.cpp:406-416 / :420-429)
__device__ __forceinline__ bool ext_side_base(const LeView &V, const ExtrusionArgs &A, int bead, int barrier, int &ndraw) {
  ndraw = 0;
  if (bead < 1 || bead > V.d.N) return false;
  const int nb = V.d.num_bond[bead - 1], bc = V.f.bondcount[bead - 1];
  if (!(nb - bc == 2 && bc == 0)) return false;
  const int ty = type_of(V, bead);
  if (!(ty == A.left || ty == A.right || ty == A.lr || ty == A.neutral)) return false;
  if (ty == barrier || ty == A.lr) ndraw = 1;
  return true;
}

// E0g: geometry records of the extruders anchored on atoms this GPU owns
__global__ void k_ext_geo(LeView V, ExtrusionArgs A) {
  const Dev &d = V.d;
  const int lo = d.own0, hi = d.own0 + d.ctrl->nown, stamp = (int)d.ctrl->le_epoch;
  const int4 *__restrict__ pos = d.pos[V.cur()];
  for (int k = lo + blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += gridDim.x * blockDim.x) {
    const int i = (pos[k].w >> 3) - 1;
    const int nb = d.num_bond[i];
    for (int m = 0; m < nb; m++) {
      if (d.bond_type[(size_t)i * d.bpa + m] != A.btype) continue;
      const int p = d.bond_atom[(size_t)i * d.bpa + m];
      const int code = bond_cross_code(d, i + 1, p);
      double r[3] = {0.0, 0.0, 0.0};
      int nd = 0;
      if (i + 1 < p) {                 // the lower end measures the three candidate bonds (fix_extrusion.cpp:431-434,451-454,492-495)
        const int a = i + 1, b = p, L = a - 1, R = b + 1;
        double xa[3], xb[3], xl[3], xr[3];
        const bool hl = L >= 1, hr = R <= d.N;
        raw_xyz(V, a, xa); raw_xyz(V, b, xb);
        if (hl) raw_xyz(V, L, xl);
        if (hr) raw_xyz(V, R, xr);
        if (hl && hr) r[0] = dist2(xl, xr);
        if (hl) r[1] = dist2(xl, xb);
        if (hr) r[2] = dist2(xa, xr);
        nd = 3;
      }
      geo_store(d, i, code, stamp, nd, r);
      break;                           // bondcount <= 1 (checked by k_ext_init)
    }
  }
}

// E0b: which atoms own a bondlist visit of an extruder, and how many draws that visit consumes
__global__ void k_ext_visits(LeView V, ExtrusionArgs A) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    int visit = 0, nd = 0;
    const int nb = d.num_bond[i];
    for (int m = 0; m < nb; m++) {
      if (d.bond_type[(size_t)i * d.bpa + m] != A.btype) continue;
      const int p = d.bond_atom[(size_t)i * d.bpa + m];
      // NTopoBondAll::build, newton_bond off: listed from atom i iff i < closest image of p
      if (!geo_fresh(V, i)) { le_raise(d.ctrl, LE_DERR_MISSING_ATOM, i + 1, p, 3); continue; }
      if (!((i + 1) < p || f.geo_i[(size_t)i * LE_GEO_I] != 21)) continue;
      const int a = min(i + 1, p), b = max(i + 1, p);
      const int nba = d.num_bond[a - 1], nbb = d.num_bond[b - 1];
      if (nba == 1 || nbb == 1 || nba == 0 || nbb == 0 || f.bondcount[a - 1] != 1 || f.bondcount[b - 1] != 1) continue;
      visit = 1;
      int n1, n2;
      ext_side_base(V, A, a - 1, A.left, n1);
      ext_side_base(V, A, b + 1, A.right, n2);
      nd = n1 + n2;
      f.partner[i] = p;   // remembered for the executor
      break;              // bondcount <= 1: at most one extruder slot
    }
    f.flag[i] = visit;
    f.ndraw[i] = nd;
  }
}

struct VisitTask {
  LeView V; ExtrusionArgs A; const int *ntask_ptr;
  __device__ int ntasks() const { return *ntask_ptr; }
  __device__ unsigned prio(int k) const { return (unsigned)V.f.tasks[k]; }
  __device__ int beads(int k, int b[4]) const {
    const int i = V.f.tasks[k];
    const int p = V.f.partner[i];
    const int lo = min(i + 1, p), hi = max(i + 1, p);
    int n = 0;
    if (lo - 1 >= 1) b[n++] = lo - 2;
    b[n++] = lo - 1; b[n++] = hi - 1;
    if (hi + 1 <= V.d.N) b[n++] = hi;
    return n;
  }
  __device__ void exec(int k) {
    const LeFixDev &f = V.f;
    const int i = f.tasks[k];
    const int p = f.partner[i];
    const int a = min(i + 1, p), b = max(i + 1, p);       // tags of i1 < i2
    const int L = a - 1, R = b + 1;
    int off = f.scan2[i];   // exclusive scan of the per-visit draw counts = position in the stream
    int n1, n2;
    bool okL = ext_side_base(V, A, L, A.left, n1);
    if (okL && n1) okL = A.p > f.draws[off++];
    bool okR = ext_side_base(V, A, R, A.right, n2);
    if (okR && n2) okR = A.p > f.draws[off++];
    const double *g = f.geo + (size_t)(a - 1) * LE_GEO_D;     // measured by the owner of a (k_ext_geo)
    if (okL && okR) {
      const double rsq = g[0];
      if (rsq >= f.distsq[L - 1] && rsq >= f.distsq[R - 1]) return;
      if (rsq < f.distsq[L - 1]) { f.distsq[L - 1] = rsq; f.to_add[L - 1] = R; }
      if (rsq < f.distsq[R - 1]) { f.distsq[R - 1] = rsq; f.to_add[R - 1] = L; }
      f.to_remove[a - 1] = b; f.to_remove[b - 1] = a;
    } else if (okL) {
      const double rsq = g[1];
      if (rsq >= f.distsq[L - 1]) return;
      f.distsq[L - 1] = rsq; f.to_add[L - 1] = b;
      if (f.distsq[b - 1] == LE_BIG) { f.distsq[b - 1] = rsq; f.to_add[b - 1] = L; }
      f.to_remove[a - 1] = b; f.to_remove[b - 1] = a;
    } else if (okR) {
      const double rsq = g[2];
      if (rsq >= f.distsq[R - 1]) return;
      if (f.distsq[a - 1] == LE_BIG) { f.distsq[a - 1] = rsq; f.to_add[a - 1] = R; }
      f.distsq[R - 1] = rsq; f.to_add[R - 1] = a;
      f.to_remove[a - 1] = b; f.to_remove[b - 1] = a;
    }
  }
};


// flag the atoms the next sequential pass iterates over
__global__ void k_ext_flag(LeView V, int which) {   // which 0: to_add != 0, 1: to_remove != 0
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x)
    f.flag[i] = (which == 0 ? f.to_add[i] : f.to_remove[i]) != 0;
}

// reconciliation of proposals that lost a contest (fix_extrusion.cpp:517-599)
struct ReconcileTask {
  LeView V; const int *ntask_ptr;
  __device__ int ntasks() const { return *ntask_ptr; }
  __device__ unsigned prio(int k) const { return (unsigned)V.f.tasks[k]; }
  __device__ void bounds(int k, int &ti, int &tj, int &lb, int &rb, bool &left) const {
    ti = V.f.tasks[k] + 1; tj = V.f.to_add[ti - 1];
    if (ti < tj) { lb = ti + 1; rb = tj - 1; left = true; } else { lb = tj + 1; rb = ti - 1; left = false; }
  }
  __device__ int beads(int k, int b[4]) const {
    int ti, tj, lb, rb; bool left; bounds(k, ti, tj, lb, rb, left);
    b[0] = ti - 1; b[1] = tj - 1; b[2] = lb - 1; b[3] = rb - 1;
    return 4;
  }
  __device__ void exec(int k) {
    const LeFixDev &f = V.f;
    int ti, tj, lb, rb; bool left; bounds(k, ti, tj, lb, rb, left);
    if (f.to_add[tj - 1] == ti) return;   // reciprocated: nothing to undo
    int *tr = f.to_remove;
    if (left) {
      if (lb == tr[rb - 1] && tr[lb - 1] == rb) { tr[lb - 1] = 0; tr[rb - 1] = 0; }
      else if (ti == tr[rb - 1] && tr[ti - 1] == rb) { tr[ti - 1] = 0; tr[rb - 1] = 0; }
      else if (lb == tr[tj - 1] && tr[lb - 1] == tj) { tr[lb - 1] = 0; tr[tj - 1] = 0; }
      else if (ti == tr[tj - 1] && tj == tr[ti - 1]) { tr[ti - 1] = 0; tr[tj - 1] = 0; }
    } else {
      if (lb == tr[rb - 1] && tr[lb - 1] == rb) { tr[lb - 1] = 0; tr[rb - 1] = 0; }
      else if (ti == tr[lb - 1] && tr[ti - 1] == lb) { tr[ti - 1] = 0; tr[lb - 1] = 0; }
      else if (rb == tr[tj - 1] && tr[rb - 1] == tj) { tr[rb - 1] = 0; tr[tj - 1] = 0; }
      else if (ti == tr[tj - 1] && tj == tr[ti - 1]) { tr[ti - 1] = 0; tr[tj - 1] = 0; }
    }
  }
};

// break loop (fix_extrusion.cpp:618-692)
struct BreakTask {
  LeView V; const int *ntask_ptr;
  __device__ int ntasks() const { return *ntask_ptr; }
  __device__ unsigned prio(int k) const { return (unsigned)V.f.tasks[k]; }
  __device__ int beads(int k, int b[4]) const {
    const int ti = V.f.tasks[k] + 1, r = V.f.to_remove[ti - 1];
    const int lb = min(ti, r), rb = max(ti, r);
    int n = 0;
    if (lb - 1 >= 1) b[n++] = lb - 2;
    b[n++] = lb - 1; b[n++] = rb - 1;
    if (rb + 1 <= V.d.N) b[n++] = rb;
    return n;
  }
  __device__ int ta(int tag) const { return (tag >= 1 && tag <= V.d.N) ? V.f.to_add[tag - 1] : 0; }
  __device__ void exec(int k) {
    const Dev &d = V.d; const LeFixDev &f = V.f;
    const int ti = f.tasks[k] + 1, r = f.to_remove[ti - 1];
    if (f.to_remove[r - 1] != ti) return;
    const int lb = min(ti, r), rb = max(ti, r);
    if (ta(lb - 1) == rb && ta(rb) == lb - 1 && ta(lb) == rb + 1 && ta(rb + 1) == lb) {
      f.to_add[lb - 2] = rb + 1; f.to_add[rb] = lb - 1; f.to_add[lb - 1] = 0; f.to_add[rb - 1] = 0;
    }
    if ((ta(lb - 1) == rb && ta(rb) == lb - 1) || (ta(lb - 1) == rb + 1 && ta(rb + 1) == lb - 1) ||
        (ta(lb) == rb + 1 && ta(rb + 1) == lb)) {
      delete_bond_slot(d, ti, r);
      special_remove12(d, ti, r);
      f.final_remove[ti - 1] = r; f.final_remove[r - 1] = ti;
      le_mark(f, 0, ti); le_mark(f, 0, r);
      if (ti < r) atomicAdd(&f.counters[CNT_NBREAK], 1);
    }
  }
};

// create loop (fix_extrusion.cpp:699-786); iterations are independent
__global__ void k_ext_create(LeView V, int btype) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const int ti = i + 1, tj = f.to_add[i];
    if (tj == 0) continue;
    if (f.to_add[tj - 1] != ti) continue;
    if (d.num_bond[i] == d.bpa) continue;
    const int lb = min(ti, tj), rb = max(ti, tj);
    auto tr = [&](int tag) { return (tag >= 1 && tag <= d.N) ? f.to_remove[tag - 1] : 0; };
    if ((tr(lb + 1) == rb && tr(rb) == lb + 1) || (tr(lb + 1) == rb - 1 && tr(rb - 1) == lb + 1) ||
        (tr(lb) == rb - 1 && tr(rb - 1) == lb)) {
      const int nb = d.num_bond[i];
      d.bond_type[(size_t)i * d.bpa + nb] = btype;
      d.bond_atom[(size_t)i * d.bpa + nb] = tj;
      d.num_bond[i] = nb + 1;
      special_insert12(d, ti, tj);
      f.final_add[i] = tj; f.final_add[tj - 1] = ti;
      le_mark(f, 1, ti); le_mark(f, 1, tj);
      if (ti < tj) atomicAdd(&f.counters[CNT_NCREATE], 1);
    }
  }
}

// bookkeeping after an event: counters, forced rebuild, sanity (fix_extrusion.cpp:788-809)
__global__ void k_le_finish(Dev d, LeFixDev f, int which) {
  const int nbreak = f.counters[CNT_NBREAK], ncreate = f.counters[CNT_NCREATE];
  Ctrl *c = d.ctrl;
  if (which == 1) {          // extrusion
    c->le_count[0] = nbreak;
    c->le_count[4] += nbreak;
    if (nbreak != ncreate) le_raise(c, LE_DERR_COUNT_MISMATCH, ncreate, nbreak);
    if (nbreak || ncreate) c->forced = 1;
  } else if (which == 2) {   // unload
    c->le_count[1] = nbreak;
    c->le_count[5] += nbreak;
    c->nbonds -= nbreak;
    if (nbreak) c->forced = 1;
  } else {                   // load
    c->le_count[2] = ncreate;
    c->le_count[6] += ncreate;
    c->nbonds += ncreate;
    if (ncreate) c->forced = 1;
  }
  f.counters[CNT_TOTAL] = nbreak + ncreate;   // gate of the topology sweeps
}

// ------------------------------------------------------------------------------------------------
// fix ex_unload (fix_ex_unload.cpp:172-372)
// ------------------------------------------------------------------------------------------------
// candidate pass: every extruder bond longer than the cutoff makes its two atoms each other's partner.
// The distance is the one the bondlist visit computes: owned-owned for a bond inside the box, owned-ghost
// (partner shifted by the image found at the last rebuild) for one straddling the boundary.  Measured by the
// owner of each atom (k_unl_geo), collected by tag on every GPU (k_unl_candidates).
__global__ void k_unl_geo(LeView V, UnloadArgs A) {
  const Dev &d = V.d;
  const int lo = d.own0, hi = d.own0 + d.ctrl->nown, stamp = (int)d.ctrl->le_epoch;
  const int4 *__restrict__ pos = d.pos[V.cur()];
  for (int k = lo + blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += gridDim.x * blockDim.x) {
    const int i = (pos[k].w >> 3) - 1;
    int partner = 0; double best = 0.0;
    const int nb = d.num_bond[i];
    for (int m = 0; m < nb; m++) {
      if (d.bond_type[(size_t)i * d.bpa + m] != A.btype) continue;
      const int p = d.bond_atom[(size_t)i * d.bpa + m];
      const int code = bond_cross_code(d, i + 1, p);
      double xi[3], xp[3];
      double rsq;
      if (code == 21) {           // one visit, from the lower tag: (x[lo] - x[hi])
        const int a = min(i + 1, p), b = max(i + 1, p);
        raw_xyz(V, a, xi); raw_xyz(V, b, xp);
        rsq = dist2(xi, xp);
      } else {                    // visited from this atom with the partner's ghost image
        raw_xyz(V, i + 1, xi); raw_xyz(V, p, xp);
        const int sh[3] = {(code & 3) - 1, ((code >> 2) & 3) - 1, ((code >> 4) & 3) - 1};
        for (int q = 0; q < 3; q++) if (sh[q]) xp[q] = __dadd_rn(xp[q], (double)sh[q] * c_P.L[q]);
        rsq = dist2(xi, xp);
      }
      if (rsq <= A.cutsq) continue;
      if (rsq > best) { best = rsq; partner = p; }
    }
    if (partner) geo_store(d, i, partner, stamp, 0, nullptr);
  }
}

__global__ void k_unl_candidates(LeView V, UnloadArgs A) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const int partner = geo_fresh(V, i) ? f.geo_i[(size_t)i * LE_GEO_I] : 0;
    f.partner[i] = partner;
    f.final_remove[i] = 0; f.final_add[i] = 0;
    f.flag[i] = partner != 0;
    if (i < 16) f.counters[i] = 0;
    if (i < 2) f.mark_n[i] = 0;
  }
}

// probability[i] = uniform() for atoms with a partner, ascending (fix_ex_unload.cpp:256-259)
__global__ void k_le_assign_draws(LeFixDev f, int n, const int *scan, double fraction) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (f.flag[i] > 0 && fraction < 1.0) f.prob[i] = f.draws[scan[i]];
}

__global__ void k_unl_break(LeView V, UnloadArgs A) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const int ti = i + 1, p = f.partner[i];
    if (p == 0) continue;
    if (f.partner[p - 1] != ti) continue;
    if (A.fraction < 1.0) {
      const double pr = (ti < p) ? f.prob[i] : f.prob[p - 1];
      if (pr >= A.fraction) continue;
    }
    delete_bond_slot(d, ti, p);
    special_remove12(d, ti, p);
    f.final_remove[i] = p; f.final_remove[p - 1] = ti;
    le_mark(f, 0, ti); le_mark(f, 0, p);
    if (ti < p) atomicAdd(&f.counters[CNT_NBREAK], 1);
  }
}

// ------------------------------------------------------------------------------------------------
// fix ex_load (fix_ex_load.cpp:329-655)
// ------------------------------------------------------------------------------------------------
// recount: fix ex_load counts the bonds of its type afresh in every event (fix_ex_load.cpp:352-366); fix bond/create counts
// once, in the setup of its first run, and from then on only adds what it creates itself (fix_bond_create.cpp:302-345, 576)
__global__ void k_load_init(LeView V, int btype, int recount) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    if (recount) {
      int bc = 0;
      const int nb = d.num_bond[i];
      for (int m = 0; m < nb; m++) if (d.bond_type[(size_t)i * d.bpa + m] == btype) bc++;
      f.bondcount[i] = bc;
    }
    f.partner[i] = 0; f.final_add[i] = 0; f.final_remove[i] = 0;
    f.distsq[i] = LE_BIG; f.claim[i] = ~0ull;
    if (i < 16) f.counters[i] = 0;
    if (i < 2) f.mark_n[i] = 0;
  }
}

// static part of the half-list scan for the pair (t, t+2), t = i+1 (fix_ex_load.cpp:447-494).
// The list fix ex_load walks is the pair list of the LAST rebuild, so membership and
// the atom the pair is stored on come from pos_hold: in the list iff rsq(pos_hold) <= cutneighsq (and not a 1-2
// special, which is re-checked below exactly as the reference does); stored on the lower tag iff
// le_pair_stored_on_i says so.  If the stored neighbor is a periodic ghost the reference reads the ghost's
// num_bond, which is never communicated (src/MOLECULE/atom_vec_bond.cpp:41) and is 0: never eligible.
// flag[i] = 0 not eligible, 1 eligible and stored on the upper tag, 2 eligible and stored on the lower tag.
// Measured for the pair (lo, lo+2) by the GPU that owns lo (k_load_geo); a partner outside its halo is farther
// away than any list cutoff.  Collected by tag on every GPU (k_load_eligible).
__global__ void k_load_geo(LeView V, LoadArgs A) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  const int own_lo = d.own0, own_hi = d.own0 + d.ctrl->nown, stamp = (int)d.ctrl->le_epoch;
  const int4 *__restrict__ pos = d.pos[V.cur()];
  for (int k = own_lo + blockIdx.x * blockDim.x + threadIdx.x; k < own_hi; k += gridDim.x * blockDim.x) {
    const int i = (pos[k].w >> 3) - 1;
    int ok = 0;
    double rsq_out = 0.0;
    if (i + 2 < d.N && d.map[i + 2] >= 0) {
      const int lo = i + 1, hi = i + 3, mid = i + 2;
      if (d.num_bond[lo - 1] == 2 && d.num_bond[hi - 1] == 2 && d.num_bond[mid - 1] == 2) {
        const int4 plo = d.pos_hold[d.map[lo - 1]], phi = d.pos_hold[d.map[hi - 1]];
        const unsigned ulo[3] = {(unsigned)plo.x, (unsigned)plo.y, (unsigned)plo.z};
        const unsigned uhi[3] = {(unsigned)phi.x, (unsigned)phi.y, (unsigned)phi.z};
        const int tlo = plo.w & 7, thi = phi.w & 7;
        if (le_pair_rsq_ref(c_P, ulo, uhi) <= c_P.cutneighsq[tlo * c_P.ntypes + thi]) {
          int ghost;
          const bool on_lo = le_pair_stored_on_i(c_P, ulo, uhi, lo, hi, &ghost);
          if (!ghost) {
            const int ti = on_lo ? lo : hi, tj = on_lo ? hi : lo;     // list owner / stored neighbor
            const int itype = type_of(V, ti), jtype = type_of(V, tj);
            int possible = 0;
            if (itype == A.itype && jtype == A.jtype) {
              if ((A.imax == 0 || f.bondcount[ti - 1] < A.imax) && (A.jmax == 0 || f.bondcount[tj - 1] < A.jmax)) possible = 1;
            } else if (itype == A.jtype && jtype == A.itype) {
              if ((A.jmax == 0 || f.bondcount[ti - 1] < A.jmax) && (A.imax == 0 || f.bondcount[tj - 1] < A.imax)) possible = 1;
            }
            if (possible) {
              const int *sl = d.special + (size_t)(ti - 1) * d.maxspecial;
              const int n1 = d.nspecial[(size_t)(ti - 1) * 3];
              for (int k = 0; k < n1; k++) if (sl[k] == tj) possible = 0;
            }
            if (possible) {
              double xi[3], xj[3];
              raw_xyz(V, ti, xi); raw_xyz(V, tj, xj);
              const double rsq = dist2(xi, xj);
              if (rsq < A.cutsq) { ok = on_lo ? 2 : 1; rsq_out = rsq; }
            }
          }
        }
      }
    }
    if (ok) geo_store(d, i, ok, stamp, 1, &rsq_out);
  }
}

__global__ void k_load_eligible(LeView V) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const bool fresh = geo_fresh(V, i);
    f.flag[i] = fresh ? f.geo_i[(size_t)i * LE_GEO_I] : 0;
    if (fresh) f.prob[i] = f.geo[(size_t)i * LE_GEO_D];
    f.done[i] = 0;
  }
}

// The half-list scan itself (fix_ex_load.cpp:433-505) is sequential: a pair is skipped when its middle bead has
// already been claimed as somebody's partner.  Pair i touches beads i, i+1, i+2 only, so two eligible pairs
// interact only if they are at most 2 apart: maximal chains of eligible pairs with gaps <= 2 ("runs") are
// independent of each other.  One thread replays one run in the reference's traversal order -- owner atom
// ascending (ilist), the pair stored on the upper tag before the one stored on the lower tag.
__device__ __forceinline__ unsigned load_prio(int i, int flag) {
  const int owner = (flag == 2) ? i + 1 : i + 3;
  return (unsigned)owner * 2u + ((flag == 2) ? 1u : 0u);
}
__global__ void k_load_scan_runs(LeView V) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  const int N = d.N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    if (f.flag[i] <= 0) continue;
    if ((i >= 1 && f.flag[i - 1] > 0) || (i >= 2 && f.flag[i - 2] > 0)) continue;   // not the head of its run
    for (;;) {
      // the pending pair of this run that the sequential scan reaches first
      int best = -1; unsigned bp = 0xffffffffu;
      int p = i;
      while (p >= 0) {
        if (!f.done[p]) { const unsigned pr = load_prio(p, f.flag[p]); if (pr < bp) { bp = pr; best = p; } }
        if (p + 1 < N && f.flag[p + 1] > 0) p = p + 1;
        else if (p + 2 < N && f.flag[p + 2] > 0) p = p + 2;
        else p = -1;
      }
      if (best < 0) break;
      f.done[best] = 1;
      if (f.partner[best + 1] != 0) continue;     // middle bead already claimed (fix_ex_load.cpp:471-476,484)
      const double rsq = f.prob[best];
      const int lo = best + 1, hi = best + 3;
      if (rsq < f.distsq[lo - 1]) { f.partner[lo - 1] = hi; f.distsq[lo - 1] = rsq; }
      if (rsq < f.distsq[hi - 1]) { f.partner[hi - 1] = lo; f.distsq[hi - 1] = rsq; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fix bond/create (src/MC/fix_bond_create.cpp:349-629), the ancestor of fix ex_load: the same event without the loop-extrusion
// rules (|tag difference| == 2, free middle bead, two backbone bonds), so ANY listed pair within the cutoff can bond and the
// partner of an atom is simply its closest eligible neighbor -- a minimum, not a traversal: no runs, no ordered executor.
// The list the fix walks is the pair list of the last rebuild (an occasional copy) with the
// CURRENT coordinates.  One warp per tile of the full list, lane = owned atom: for every entry the lane works out which end
// the reference's HALF list stores the pair on (le_pair_stored_on_i: the type / bond-count test of fix_bond_create.cpp:432-441
// is not symmetric when both types are equal but the limits differ, and the 1-2 check :447-449 reads the list owner's
// specials), measures the distance as that visit does -- owner minus stored neighbor, a periodic ghost shifted by the image
// found at the last rebuild -- and keeps the smallest (ties, which need two bit-equal squared distances, go to the lower
// partner tag; the reference would keep the pair it visits first).  Measured by the owner of each atom, collected by tag on
// every GPU like the records of the other fixes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bcreate_geo(LeView V, LoadArgs A) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  const int lane = threadIdx.x & 31;
  const int nown = d.ctrl->nown, own_end = d.own0 + nown, stamp = (int)d.ctrl->le_epoch;
  const int ntiles = (nown + TILE - 1) / TILE;
  const int nt = c_P.ntypes;
  for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles; tile += (gridDim.x * blockDim.x) >> 5) {
    const int k = d.own0 + tile * TILE + lane;
    const bool valid = k < own_end;
    const int nn = valid ? (int)AUX_NN(__float_as_uint(d.vel[k].w)) : 0;
    int inc = nn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (!valid || nn == 0) continue;
    const unsigned *__restrict__ run = d.nbr + (size_t)tile * d.tcap + (inc - nn);
    const int4 hi_ = d.pos_hold[k];
    const int ti = hi_.w >> 3;
    const unsigned ui[3] = {(unsigned)hi_.x, (unsigned)hi_.y, (unsigned)hi_.z};
    double xi[3];
    raw_xyz(V, ti, xi);
    const int type_i = type_of(V, ti), bc_i = f.bondcount[ti - 1];
    int best = 0; double best_rsq = LE_BIG;
    for (int q = 0; q < nn; q++) {
      const int j = (int)(run[q] & NEIGH_IDX_MASK);
      const int4 hj = d.pos_hold[j];
      const int tj = hj.w >> 3;
      const unsigned uj[3] = {(unsigned)hj.x, (unsigned)hj.y, (unsigned)hj.z};
      int ghost;
      const bool on_i = le_pair_stored_on_i(c_P, ui, uj, ti, tj, &ghost);
      const int type_j = type_of(V, tj), bc_j = f.bondcount[tj - 1];
      // (itype, bondcount) of the list owner / of the stored neighbor
      const int ot = on_i ? type_i : type_j, nty = on_i ? type_j : type_i, obc = on_i ? bc_i : bc_j, nbc = on_i ? bc_j : bc_i;
      int possible = 0;
      if (ot == A.itype && nty == A.jtype) {
        if ((A.imax == 0 || obc < A.imax) && (A.jmax == 0 || nbc < A.jmax)) possible = 1;
      } else if (ot == A.jtype && nty == A.itype) {
        if ((A.jmax == 0 || obc < A.jmax) && (A.imax == 0 || nbc < A.imax)) possible = 1;
      }
      if (!possible) continue;
      const int otag = on_i ? ti : tj, ntag = on_i ? tj : ti;
      const int *sl = d.special + (size_t)(otag - 1) * d.maxspecial;
      const int n1 = d.nspecial[(size_t)(otag - 1) * 3];
      for (int m = 0; m < n1; m++) if (sl[m] == ntag) possible = 0;
      if (!possible) continue;
      double xj[3];
      raw_xyz(V, tj, xj);
      double rsq;
      if (!ghost) rsq = on_i ? dist2(xi, xj) : dist2(xj, xi);
      else {
        // the stored neighbor is a periodic ghost: its coordinate is the owned atom's plus the shift of the image that was
        // closest to the list owner at the last rebuild (AtomVec::pack_border / pack_comm: x + pbc * prd)
        double xo[3], xn[3];
        for (int a = 0; a < 3; a++) {
          xo[a] = on_i ? xi[a] : xj[a];
          xn[a] = on_i ? xj[a] : xi[a];
          const int sh = on_i ? le_image_shift(ui[a], uj[a]) : le_image_shift(uj[a], ui[a]);
          if (sh) xn[a] = __dadd_rn(xn[a], (double)sh * c_P.L[a]);
        }
        rsq = dist2(xo, xn);
      }
      if (rsq >= A.cutsq) continue;
      if (rsq < best_rsq || (rsq == best_rsq && tj < best)) { best = tj; best_rsq = rsq; }
    }
    if (best) geo_store(d, ti - 1, best, stamp, 1, &best_rsq);
  }
  (void)nt;
}

__global__ void k_bcreate_collect(LeView V) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const bool fresh = geo_fresh(V, i);
    f.partner[i] = fresh ? f.geo_i[(size_t)i * LE_GEO_I] : 0;
    f.distsq[i] = fresh ? f.geo[(size_t)i * LE_GEO_D] : LE_BIG;
  }
}

__global__ void k_load_flag_partners(LeFixDev f, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    f.flag[i] = f.partner[i] != 0;
    f.prob[i] = LE_BIG;      // `probability` overlays distsq in the reference; only drawn entries are read
  }
}

__global__ void k_load_create(LeView V, LoadArgs A) {
  const Dev &d = V.d; const LeFixDev &f = V.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const int ti = i + 1, p = f.partner[i];
    if (p == 0) continue;
    if (f.partner[p - 1] != ti) continue;
    if (A.fraction < 1.0) {
      const double pr = (ti < p) ? f.prob[i] : f.prob[p - 1];
      if (pr >= A.fraction) continue;
    }
    const int nb = d.num_bond[i];
    if (nb == d.bpa) { le_raise(d.ctrl, LE_DERR_BOND_OVERFLOW, ti, nb); continue; }
    d.bond_type[(size_t)i * d.bpa + nb] = A.btype;
    d.bond_atom[(size_t)i * d.bpa + nb] = p;
    d.num_bond[i] = nb + 1;
    special_insert12(d, ti, p);
    const int bc = f.bondcount[i] + 1;
    f.bondcount[i] = bc;
    // type[i] = inewtype / jnewtype (fix_ex_load.cpp:594-598): the replicated table, and the copy of the atom this
    // GPU holds (owned or ghost); the forced rebuild that follows refreshes pos_hold from it
    const int ty = d.type_tag[i];
    int nty = ty;
    if (ty == A.itype) { if (bc == A.imax) nty = A.inew; }
    else { if (bc == A.jmax) nty = A.jnew; }
    if (nty != ty) {
      d.type_tag[i] = nty;
      const int k = d.map[i];
      if (k >= 0) { int4 *pp = &d.pos[V.cur()][k]; pp->w = (pp->w & ~7) | (nty - 1); }
    }
    f.final_add[i] = p; f.final_add[p - 1] = ti;
    le_mark(f, 1, ti); le_mark(f, 1, p);
    if (ti < p) atomicAdd(&f.counters[CNT_NCREATE], 1);
  }
}
