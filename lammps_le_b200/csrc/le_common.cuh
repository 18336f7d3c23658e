// le_common.cuh -- device data layout and small helpers shared by every kernel.
//
// Layout in HBM (N atoms in the whole system, `cap` local slots on this GPU; all arrays device-resident):
//   local order (index k = slot after the last cell sort; rewritten at every rebuild).  One GPU owns the x-slab of
//   cells [X0, X1) and keeps `halo` layers of ghost cells on either side; slots are laid out by region
//     [0, own0)            ghosts of the left neighbor's last `halo` layers   (written by that GPU over NVLink)
//     [own0, own0 + nown)  owned atoms, sorted by cell (x slowest, z fastest), by tag inside a cell
//     [gr0, cap)           ghosts of the right neighbor's first `halo` layers
//   (one GPU: own0 = 0, nown = N, no ghosts: the periodic wrap comes from the fixed-point differences)
//     pos[2][cap]   int4   {ux,uy,uz: 32-bit fixed-point box fractions, w: tag<<3 | type-1}   double-buffered
//     vel[cap]      float4 {vx,vy,vz, w: aux word -- list counts and the displacement bound, see AUX_* below}
//     pos_hold[cap] int4   positions at the last rebuild (Neighbor::xhold, src/neighbor.cpp:2048-2052)
//     img[cap], img_hold[cap]  LAMMPS-packed image flags now / at the last rebuild (owned atoms)
//     bondrow[bpa][cap]     ELL bond partner rows; entry = k_j | (bondtype-1)<<28
//   neighbor list: the owned slots are cut into TILES of 32 consecutive slots (one warp of the step kernel).  The full
//   list of a tile is ONE flat run of entries, grouped by owner (slot order), each pair (i,j) present in the run of
//   i's tile and in the run of j's tile:
//     tile_cnt[ntiles]            entries of the tile
//     nbr[ntiles][32 * maxneigh]  entry = k_j | owner_lane<<25 | which<<30   (k_j < 2^25 local slots per GPU)
//     nbr_ell[maxneigh][cap]      scratch of the list build (per-atom rows before they are packed into the tile run)
//   tag order (index t-1; what the reference's Atom class holds, src/atom.h) -- replicated on every GPU
//     num_bond, bond_type[N][bpa], bond_atom[N][bpa], nspecial[N][3], special[N][maxspecial]
//     topo[N]     64-byte digest of the five tables above for the list build (TopoRec), refreshed when they change
//     map[N]      tag-1 -> local slot, -1 if the atom is neither owned nor a ghost here (Atom::map, src/atom.h:354-358)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LE_MAXT 8    // atom types
#define LE_MAXB 8    // bond types
#define LE_BIG 1.0e20
#define NEIGH_IDX_BITS 25
#define NEIGH_IDX_MASK 0x01ffffffu
#define BOND_IDX_MASK 0x0fffffffu
#define TILE 32

// aux word of an owned atom (vel[k].w, as bits): neighbor count, bond count, and an upper bound of the distance the atom
// has moved since the last rebuild in units of (skin/2)/65536 (saturating).  The step kernel adds each step's path
// length to the bound (rounded up) and looks at pos_hold -- the exact test of Neighbor::check_distance -- only for the
// few atoms whose bound has reached skin/2: the 16-byte pos_hold read leaves the per-step working set.
#define AUX_NN(a) ((a) & 0xffu)
#define AUX_NB(a) (((a) >> 8) & 0xfu)
#define AUX_BOUND(a) ((a) >> 12)
#define AUX_BOUND_ONE 65536u         // bound value that means "skin/2"
#define AUX_BOUND_MAX 0xfffffu
#define AUX_PACK(nn, nb, bound) ((unsigned)(nn) | ((unsigned)(nb) << 8) | ((unsigned)(bound) << 12))

// digest of one atom's topology for the list build, tag order.  bond slots beyond 4 / specials beyond 10 are read from
// the full tables (rare)
struct __align__(16) TopoRec {
  unsigned hdr;        // num_bond | nscan<<8 | n1<<16 | n2<<24   (nscan: special entries find_special must scan)
  unsigned btypes;     // bond type - 1 of slots 0..7, 4 bits each
  int batom[4];        // bond partner tags of slots 0..3
  int spec[10];        // special[0..9]
};
#define TOPO_NSPEC 10

struct Params {
  // box and fixed-point mapping: x = lo + u*scale, scale = L/2^32
  double lo[3], hi[3], L[3], scale[3], half[3];
  float fscale[3], inv_fscale[3];
  int periodic[3];
  // the reference's neighbor bins (NBinStandard::setup_bins, src/nbin_standard.cpp:53-186); used only to
  // reproduce which atom of a pair stores it in the half list
  double bininv[3];
  int nbin[3];
  // pair lj/cut (PairLJCut::init_one, src/pair_lj_cut.cpp:512-535)
  int ntypes, nbondtypes;
  int pair_uniform;     // every type pair has the same lj/cut coefficients -> table index 0
  double cutneighsq[LE_MAXT * LE_MAXT];
  float cutneighmaxsq_f;
  float cutsq[LE_MAXT * LE_MAXT], lj1[LE_MAXT * LE_MAXT], lj2[LE_MAXT * LE_MAXT];
  float lj3[LE_MAXT * LE_MAXT], lj4[LE_MAXT * LE_MAXT], offset[LE_MAXT * LE_MAXT];
  float cutsq_screen[LE_MAXT * LE_MAXT];   // fp32 screen, slightly above cutsq
  double cutsq_d[LE_MAXT * LE_MAXT], lj1_d[LE_MAXT * LE_MAXT], lj2_d[LE_MAXT * LE_MAXT];
  double lj3_d[LE_MAXT * LE_MAXT], lj4_d[LE_MAXT * LE_MAXT], offset_d[LE_MAXT * LE_MAXT];
  float special_lj[4];
  int special_flag[4];
  int nscan_tier;       // how many special tiers find_special must scan (0..3)
  float mass[LE_MAXT];
  // bonds
  int bstyle[LE_MAXB];
  float bk[LE_MAXB], br0[LE_MAXB], beps[LE_MAXB], bsig[LE_MAXB];
  double bk_d[LE_MAXB], br0_d[LE_MAXB], beps_d[LE_MAXB], bsig_d[LE_MAXB];
  double br0sq_d[LE_MAXB], binvr0sq_d[LE_MAXB], bsig2_d[LE_MAXB], bcore_d[LE_MAXB];  // R0^2, 1/R0^2, sigma^2, 2^(1/3) sigma^2
  double beps48_d[LE_MAXB];   // 48 epsilon (k_step2)
  int astyle[LE_MAXB];        // angle styles by angle type (0 none, 1 cosine)
  double ak_d[LE_MAXB];       // angle cosine: K
  // fp32 brackets around cutneighsq: below lo a pair is certainly listed, above hi certainly not; only the
  // sliver in between needs the reference's fp64 arithmetic (k_build)
  float cutneigh_lo[LE_MAXT * LE_MAXT], cutneigh_hi[LE_MAXT * LE_MAXT];
  float t_start, t_stop, tsqrt_const;
  float inv_bound_unit;  // 65536 / (skin/2) (aux displacement bound)
  float dtfm[LE_MAXT];   // dtf / mass
  // integration / thermostat
  float dt, dtf;
  float triggersq;
  float vlimitsq;       // fix nve/limit: (xmax/dt)^2, 0 = off
  int nve_on, langevin_on;
  float gfac1[LE_MAXT], gfac2[LE_MAXT];
  unsigned seed_lo, seed_hi;
  // neighbor policy
  int every, delay, check;
};

struct Ctrl {
  int moved;          // some atom moved more than skin/2 since the last rebuild
  int forced;         // a fix asked for a rebuild on this step (Fix::next_reneighbor)
  int rebuild_now;    // decision of the last k_decide
  int ago;
  int err;            // first run-time error code (0 = none)
  int err_info[4];
  long long nbuilds, ndanger;
  long long fene_warn;
  // USER-LE counters
  int le_count[8];
  long long nbonds;
  // device-side run state, so that captured graphs are identical from step to step
  int cur;                    // which of pos[0]/pos[1] holds the current coordinates
  int pad0;
  long long step;             // timestep of the next force evaluation (Update::ntimestep)
  long long run_begin, run_end;  // Update::beginstep / endstep of the current run (Langevin ramp)
  // local population (changes at every rebuild when atoms migrate between GPUs)
  int nown;                   // owned atoms: slots [own0, own0 + nown)
  int nghl, nghr;             // ghosts in [0, nghl) and [gr0, gr0 + nghr)
  int send_l_end;             // owned slots [own0, send_l_end) are the left neighbor's right ghosts
  int send_r_beg;             // owned slots [send_r_beg, own0 + nown) are the right neighbor's left ghosts
  int nown_unsorted;          // owned + arrived atoms before the sort of a rebuild
  unsigned blocks_done;       // k_step blocks that have finished their stores (last one signals the peers)
  int pad1;
  long long epoch;            // force evaluations so far: the value the per-step peer flags carry
  long long rebuild_epoch;    // rebuilds so far (all GPUs rebuild on the same steps)
  long long le_epoch;         // USER-LE exchange rounds so far
  unsigned scan_ticket;       // k_scan_cells: tile tickets (reset by the block that finishes last)
  unsigned scan_done;
  unsigned nbuilds_scan;      // launches of k_scan_cells so far (the epoch of its state words)
  unsigned pad3;
};

enum {
  LE_DERR_NONE = 0,
  LE_DERR_BAD_FENE = 1,
  LE_DERR_NEIGH_OVERFLOW = 2,
  LE_DERR_BONDCOUNT = 3,
  LE_DERR_SPECIAL_OVERFLOW = 4,
  LE_DERR_BOND_OVERFLOW = 5,
  LE_DERR_COUNT_MISMATCH = 6,
  LE_DERR_MISSING_ATOM = 7,
  LE_DERR_CELL_OVERFLOW = 8,
  LE_DERR_RNG_OVERFLOW = 9,
  LE_DERR_PEER_TIMEOUT = 10,
  LE_DERR_LOCAL_OVERFLOW = 11
};

#define LE_MAXRANKS 8
#define LE_GEO_D 3    // USER-LE geometry record, doubles per tag (see le_fix.cuh)
#define LE_GEO_I 2    // ... ints per tag
// what one GPU sees of another GPU's arena (CUDA IPC mapping); all offsets are identical on every rank
struct PeerView {
  int4 *pos[2];
  int4 *pos_hold;
  int *cell_start;
  int4 *in_pos; float4 *in_vel; int *in_img;       // migration inbox [2][inbox_cap]: side 0 = from its left neighbor
  unsigned long long *flags;                        // see FLAG_* below
  double *geo;                                      // USER-LE geometry records, tag order
  int *geo_i;
};
// flag words in every arena, written by peers with system-scope stores:
//   [FLAG_STEP + src]      (epoch << 1) | moved          after src's k_step of that epoch has stored its halo
//   [FLAG_INBOX + side]    (rebuild_epoch << 24) | count  migrants written into inbox `side`
//   [FLAG_GHOST + side]    (rebuild_epoch << 24) | count  ghost slice `side` (0 = left ghosts) written
//   [FLAG_LE + src]        le_epoch                        src's USER-LE geometry records stored
enum { FLAG_STEP = 0, FLAG_INBOX = 16, FLAG_GHOST = 24, FLAG_LE = 32, FLAG_WORDS = 64 };

struct Dev {
  int N;              // atoms in the whole system = length of the tag-indexed arrays
  int cap;            // local slots (owned + ghosts) = stride of the ELL rows
  int own0, gr0;      // first owned slot / first right-ghost slot
  int bpa, maxspecial, maxneigh;
  // domain decomposition: x-slabs of cells
  int nranks, rank, halo;
  int X0, X1, nlx;    // owned global x-cells [X0, X1); local layers nlx = X1 - X0 + 2 halo
  int nlx_left, nlx_right;   // the neighbors' layer counts (slab widths may differ by one)
  int inbox_cap;
  PeerView peer[LE_MAXRANKS];   // peer[rank] is this GPU's own arena
  unsigned long long *flags;    // this GPU's flag words
  int4 *in_pos; float4 *in_vel; int *in_img;
  int4 *pos[2];
  int4 *pos_hold;
  float4 *vel, *vel_tmp;
  int *img, *img_hold;
  unsigned *bondrow;
  unsigned *tile_cnt, *nbr, *nbr_ell;
  int tcap;           // entries per tile run = 32 * maxneigh
  TopoRec *topo;
  int4 *order2;       // cell sort: {old slot, tag, cell start, cell end} in cell order
  unsigned long long *scan_state;   // single-pass scan of the cell counts (decoupled look-back)
  // tag order
  int *num_bond, *bond_type, *bond_atom, *nspecial, *special, *map;
  int *type_tag;      // atom type by tag (the USER-LE fixes read and change types of atoms that may live on another GPU)
  // cell sort scratch
  int *cell_count, *cell_start, *cellid, *slot;
  int *ghost_tag;     // [2 own0] tags of the current ghosts (left, then right)
  int ncell[3];       // global cell grid
  int ncells;         // local cell slots incl. region sentinels: nlx*ncy*ncz + 3
  int nscanblocks;
  int cell_span[3], cell_abs[3];  // per-dim stencil span / "visit all cells" (k_build)
  // angles (tag order, replicated): ang[nangles] = {type, a1, a2, a3}; per atom the angles it takes part in
  int nangles, apa;
  int4 *ang; int *ang_cnt, *ang_idx;
  double *fang;     // [cap][3] angle force of the owned atoms (k_angle -> step kernel)
  Ctrl *ctrl;
  double *thermo;   // [slots][LE_THERMO_W]
  double *fout;     // [N][3] optional force output (tag order)
};

#define LE_THERMO_W 24
// thermo slot layout: 0 ke(sum m v^2) 1 evdwl 2 ebond 3..8 virial 9 fene warnings 10 eangle | snapshot at that force evaluation (not
// summed over GPUs: the USER-LE state is replicated): 16 atom->nbonds, 17..19 f_ID[1] of extrusion / ex_unload / ex_load
// (bonds of the fix's last event), 20..22 f_ID[2] (cumulative)

// one translation unit (le_engine.cu) includes this header; the block is refreshed before every use
__constant__ Params c_P;

// slot of local cell (lx, cy, cz) in cell_start / cell_count.  x is the slowest index so that a slab's boundary
// layers are contiguous slices of the local order; one sentinel slot follows each region (left ghosts / owned /
// right ghosts) so that cell_start[c + 1] is always the end of cell c.
__host__ __device__ __forceinline__ int cell_slot(const Dev &d, int lx, int cy, int cz) {
  return (lx * d.ncell[1] + cy) * d.ncell[2] + cz + (lx >= d.halo) + (lx >= d.nlx - d.halo);
}
// local layer of a global x-cell: [halo, nlx - halo) owned, [0, halo) left halo, [nlx - halo, nlx) right halo,
// -1 if the cell is neither in this GPU's slab nor in its halo
__host__ __device__ __forceinline__ int local_layer(const Dev &d, int cx) {
  int dd = cx - d.X0;
  if (dd < 0) dd += d.ncell[0];
  if (dd < d.X1 - d.X0 + d.halo) return dd + d.halo;
  if (dd >= d.ncell[0] - d.halo) return dd - d.ncell[0] + d.halo;
  return -1;
}

__device__ __forceinline__ void le_raise(Ctrl *c, int code, int a = 0, int b = 0, int e = 0, int f = 0) {
  if (atomicCAS(&c->err, 0, code) == 0) {
    c->err_info[0] = a; c->err_info[1] = b; c->err_info[2] = e; c->err_info[3] = f;
  }
}

// dequantised coordinate exactly as the host does it: one multiply, one add, no fma contraction
__device__ __forceinline__ double le_deq(unsigned u, int d) {
  return __dadd_rn(c_P.lo[d], __dmul_rn((double)u, c_P.scale[d]));
}

// Philox4x32-7 (Salmon et al. 2011: seven rounds already pass BigCrush) -- counter-based generator for the
// Langevin noise, keyed by (seed), counted by (tag, timestep)
__device__ __forceinline__ void philox4x32_7(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                              unsigned k0, unsigned k1, unsigned out[4]) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; r++) {
    unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---- reference arithmetic shared by the device list build and the host-side list download ----
#ifdef __CUDA_ARCH__
#define LE_DADD(a, b) __dadd_rn(a, b)
#define LE_DSUB(a, b) __dsub_rn(a, b)
#define LE_DMUL(a, b) __dmul_rn(a, b)
#else
#define LE_DADD(a, b) ((a) + (b))
#define LE_DSUB(a, b) ((a) - (b))
#define LE_DMUL(a, b) ((a) * (b))
#endif

// periodic shift of j's closest image relative to i in one dimension: (uj - ui)_wrapped - (uj - ui)_raw = s * 2^32
__host__ __device__ __forceinline__ int le_image_shift(unsigned ui, unsigned uj) {
  const int w = (int)(uj - ui);
  return (int)(((long long)w - ((long long)uj - (long long)ui)) >> 32);
}

// dequantised coordinate with explicit Params (one multiply, one add, no fma contraction)
__host__ __device__ __forceinline__ double le_deq_p(const Params &P, unsigned u, int d) {
  return LE_DADD(P.lo[d], LE_DMUL((double)u, P.scale[d]));
}

// NBin::coord2bin for one dimension (src/nbin.cpp:120-150)
__host__ __device__ __forceinline__ int le_ref_bin(const Params &P, double x, int dim) {
  const double lo = P.lo[dim], hi = P.hi[dim], inv = P.bininv[dim];
  const int nb = P.nbin[dim];
  int ix;
  if (x >= hi) ix = (int)LE_DMUL(LE_DSUB(x, hi), inv) + nb;
  else if (x >= lo) { ix = (int)LE_DMUL(LE_DSUB(x, lo), inv); if (ix > nb - 1) ix = nb - 1; }
  else ix = (int)LE_DMUL(LE_DSUB(x, lo), inv) - 1;
  return ix;
}

// coordinates of i and of the image of j closest to i, as the reference holds them (owned atom / ghost made by
// AtomVec::pack_border, x + pbc*prd)
__host__ __device__ __forceinline__ void le_pair_coords(const Params &P, const unsigned ui[3], const unsigned uj[3],
                                                        double xi[3], double xj[3], int sh[3]) {
  for (int q = 0; q < 3; q++) {
    xi[q] = le_deq_p(P, ui[q], q);
    sh[q] = le_image_shift(ui[q], uj[q]);
    double v = le_deq_p(P, uj[q], q);
    if (sh[q]) v = LE_DADD(v, (double)sh[q] * P.L[q]);
    xj[q] = v;
  }
}

// squared distance with the reference's operation order (npair_half_bin_newton.cpp:98-102)
__host__ __device__ __forceinline__ double le_pair_rsq_ref(const Params &P, const unsigned ui[3], const unsigned uj[3]) {
  double xi[3], xj[3]; int sh[3];
  le_pair_coords(P, ui, uj, xi, xj, sh);
  const double dx = LE_DSUB(xi[0], xj[0]), dy = LE_DSUB(xi[1], xj[1]), dz = LE_DSUB(xi[2], xj[2]);
  return LE_DADD(LE_DADD(LE_DMUL(dx, dx), LE_DMUL(dy, dy)), LE_DMUL(dz, dz));
}

// would NPairHalfBinNewton::build store the pair (i, j) in the list of atom i?  (same-bin rule
// npair_half_bin_newton.cpp:84-91, upper-half stencil nstencil_half_bin_3d_newton.cpp:26-38)
__host__ __device__ __forceinline__ bool le_pair_stored_on_i(const Params &P, const unsigned ui[3], const unsigned uj[3],
                                                             int tagi, int tagj, int *ghost) {
  double xi[3], xj[3]; int sh[3];
  le_pair_coords(P, ui, uj, xi, xj, sh);
  const int dbx = le_ref_bin(P, xj[0], 0) - le_ref_bin(P, xi[0], 0);
  const int dby = le_ref_bin(P, xj[1], 1) - le_ref_bin(P, xi[1], 1);
  const int dbz = le_ref_bin(P, xj[2], 2) - le_ref_bin(P, xi[2], 2);
  const int gh = sh[0] | sh[1] | sh[2];
  *ghost = gh != 0;
  if ((dbx | dby | dbz) == 0) {
    if (gh == 0) return tagj > tagi;       // owned j later in the bin's list (ascending local index == tag)
    return !(xj[2] < xi[2] || (xj[2] == xi[2] && (xj[1] < xi[1] || (xj[1] == xi[1] && xj[0] < xi[0]))));  // ghost j
  }
  return dbz > 0 || (dbz == 0 && (dby > 0 || (dby == 0 && dbx > 0)));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
