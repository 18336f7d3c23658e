/* le_b200.h -- C ABI of the B200-native chromatin / loop-extrusion MD engine.
 *
 * This is the drop-in boundary for ONE hot path of polly-code/lammps_le: the
 * Verlet timestep (src/verlet.cpp:223-354) of a bead-spring system with
 *   pair_style lj/cut (WCA)          src/pair_lj_cut.cpp:68-140
 *   bond_style fene | harmonic       src/MOLECULE/bond_fene.cpp:52-128, bond_harmonic.cpp:48-100
 *   fix nve + fix langevin           src/fix_nve.cpp:64-140, src/fix_langevin.cpp:587-777
 *   fix extrusion|ex_load|ex_unload  src/USER-LE/fix_extrusion.cpp:256-872,
 *                                    fix_ex_load.cpp:329-655, fix_ex_unload.cpp:172-372
 *   binned half neighbor lists       src/npair_half_bin_newton.cpp:35-160
 * All state lives on one GPU; every entry point takes plain pointers and
 * sizes (host memory unless stated) and returns 0 on success or a negative
 * LE_E* code, with the message available from le_last_error().  This mirrors
 * the only C-ABI precedent in the reference, the GPU package
 * (src/GPU/pair_lj_cut_gpu.cpp:42-66: ljl_gpu_init/compute/clear).
 *
 * Conventions (those of the reference's Atom class, src/atom.h):
 *   - atom ids ("tags") are 1..N, contiguous; every per-atom array passed in
 *     or out is in TAG ORDER (entry t-1 belongs to tag t), the order the
 *     1-rank reference keeps under `atom_modify sort 0 0`;
 *   - bonds are stored on BOTH atoms (`newton on off`, the only mode in which
 *     USER-LE works): num_bond[N],
 *     bond_type[N*bond_per_atom], bond_atom[N*bond_per_atom];
 *   - special lists as in src/special.cpp: nspecial[N*3] cumulative counts,
 *     special[N*maxspecial];
 *   - units lj (boltz = mvv2e = ftm2v = nktv2p = 1).
 * There is no CPU fallback: every call fails with LE_ENOGPU without a device.
 */
#ifndef LE_B200_H
#define LE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct le_ctx le_ctx;

enum {
  LE_OK = 0,
  LE_EINVAL = -1,   /* bad argument (the reference's error->all "Illegal ... command") */
  LE_ENOGPU = -2,   /* no CUDA device / CUDA runtime failure */
  LE_ESTATE = -3,   /* call out of order (e.g. run before atoms were uploaded) */
  LE_ERUN = -4,     /* run-time abort raised on device: "Bad FENE bond", neighbor overflow,
                       "more than one bond type 2", special list overflow, created != broken */
  LE_ENOMEM = -5
};

enum { LE_BOND_NONE = 0, LE_BOND_FENE = 1, LE_BOND_HARMONIC = 2 };
enum { LE_FIX_EXTRUSION = 1, LE_FIX_EX_UNLOAD = 2, LE_FIX_EX_LOAD = 3 };

/* thermo record, one per thermo output step (thermo_style custom step temp epair emol etotal press
 * + the f_ID[1] counters of the three USER-LE fixes); energies are per atom (thermo_modify norm yes,
 * the lj default, src/thermo.cpp). */
typedef struct le_thermo {
  int64_t step;
  double temp, epair, emol, etotal, press;
  double ke;                /* total kinetic energy (not normalised) */
  double virial[6];         /* pair + bond virial, xx yy zz xy xz yz (not normalised) */
  int64_t nbonds;           /* atom->nbonds */
  int64_t fene_warnings;    /* "FENE bond too long" occurrences in this force evaluation */
  /* the USER-LE fixes' global vectors as `thermo_style custom ... f_ID[1] f_ID[2]` prints them on this step
   * (compute_vector: src/USER-LE/fix_extrusion.cpp:1496-1501, fix_ex_unload.cpp:839-843, fix_ex_load.cpp:1451-1455):
   * index 0 fix extrusion, 1 fix ex_unload, 2 fix ex_load; f1 = bonds of the fix's last event, f2 = cumulative -- except
   * le_f2[0], which is 0 as in the reference (FixExtrusion never accumulates its total; le_stats.extrusion_shifts does) */
  int64_t le_f1[3], le_f2[3];
  double eangle;            /* angle energy per atom (part of emol) */
} le_thermo;

/* run statistics, the numbers the reference prints in Finish::end (src/finish.cpp) */
typedef struct le_stats {
  int64_t steps;            /* timesteps run so far */
  int64_t neigh_builds;     /* "Neighbor list builds" */
  int64_t dangerous_builds; /* "Dangerous builds" */
  int64_t half_pairs;       /* stored half neighbor pairs at the last build ("Total # of neighbors") */
  int64_t full_entries;     /* stored full-list entries at the last build */
  int64_t kernel_launches;  /* kernels this library launched (all calls) */
  int64_t extrusion_shifts, loads, unloads;         /* cumulative event counters */
  int64_t last_extrusion_shifts, last_loads, last_unloads; /* f_ID[1] of the three fixes */
  double last_run_gpu_ms;   /* CUDA-event time of the last le_run's step loop */
  double last_run_le_ms;    /* ... of which inside the USER-LE fixes (post_integrate) */
} le_stats;

/* ---- lifetime ------------------------------------------------------------------------- */
int  le_create(le_ctx **out, int device, const double boxlo[3], const double boxhi[3],
               const int periodic[3]);                      /* Domain::set_initial_box */
void le_destroy(le_ctx *c);
const char *le_last_error(const le_ctx *c);                 /* library.h:236-237 precedent */
const char *le_version(void);

/* ---- force field and run settings (input-script commands) ------------------------------ */
int le_set_types(le_ctx *c, int ntypes, const double *mass /*[ntypes]*/, int nbondtypes);
/* pair_style lj/cut + pair_coeff + pair_modify shift; matrices are [ntypes*ntypes], row-major, 0-based
 * (entry (i-1)*ntypes+(j-1)), already mixed (PairLJCut::init_one, src/pair_lj_cut.cpp:512-535). */
int le_set_pair_lj(le_ctx *c, int ntypes, const double *epsilon, const double *sigma,
                   const double *cut, int shift_flag);
/* bond_coeff: fene params = {K, R0, epsilon, sigma}; harmonic params = {K, r0, -, -} */
int le_set_bond(le_ctx *c, int btype, int style, const double params[4]);
/* angle_style cosine (src/MOLECULE/angle_cosine.cpp:47-170; chain stiffness): le_set_angle_types = the data file's
 * `N angle types`; angle_coeff: cosine params = {K, -, -, -}; le_upload_angles = the "Angles" section (type a1 a2 a3, a2 centre) */
enum { LE_ANGLE_NONE_STYLE = 0, LE_ANGLE_COSINE_STYLE = 1 };
int le_set_angle_types(le_ctx *c, int nangletypes);
int le_set_angle(le_ctx *c, int atype, int style, const double params[4]);
int le_upload_angles(le_ctx *c, int nangles, const int *atype, const int *a1, const int *a2, const int *a3);
int le_set_special(le_ctx *c, const double lj[3]);          /* special_bonds lj a b c (fene = 0 1 1) */
int le_set_neighbor(le_ctx *c, double skin, int every, int delay, int check);
int le_set_neighbor_capacity(le_ctx *c, int max_neighbors_per_atom);  /* neigh_modify one (full rows) */
int le_set_newton(le_ctx *c, int newton_pair, int newton_bond); /* only (1,0) and (1,1) accepted */
int le_set_capacity(le_ctx *c, int bond_per_atom, int maxspecial); /* data-file "extra ... per atom" */
int le_set_timestep(le_ctx *c, double dt);                  /* timestep */
int le_reset_timestep(le_ctx *c, int64_t step);             /* reset_timestep */
int le_thermo_every(le_ctx *c, int nevery);                 /* thermo N (0 = first/last step only) */

/* ---- fixes ----------------------------------------------------------------------------- */
int le_fix_nve(le_ctx *c, int enable);
/* fix nve/limit xmax (src/fix_nve_limit.cpp:70-140): cap the per-step displacement; xmax <= 0 = plain nve */
int le_fix_nve_limit(le_ctx *c, double xmax);
/* fix langevin Tstart Tstop damp seed: uniform noise as FixLangevin::post_force_templated<0,...>,
 * but from a counter-based generator keyed by (seed, step, tag) instead of a sequential RanMars. */
int le_fix_langevin(le_ctx *c, double t_start, double t_stop, double damp, int seed);
/* fix ID all extrusion N neutral left right p_through btype [roadblock]; roadblock = -1 if absent.
 * The reference hard-codes seed 12345 (fix_extrusion.cpp:98); seed <= 0 selects that. */
int le_fix_extrusion(le_ctx *c, int nevery, int neutral, int left, int right, double p_through,
                     int btype, int roadblock, int seed);
/* fix ID all ex_load N itype jtype Rmin btype prob f seed iparam imax inew jparam jmax jnew */
int le_fix_ex_load(le_ctx *c, int nevery, int itype, int jtype, double rc, int btype, double prob,
                   int seed, int imaxbond, int inewtype, int jmaxbond, int jnewtype);
/* fix ID all ex_unload N btype Rmax prob f seed */
int le_fix_ex_unload(le_ctx *c, int nevery, int btype, double rc, double prob, int seed);
/* fix ID all bond/break N bondtype Rmax prob f seed (src/MC/fix_bond_break.cpp, the ancestor of ex_unload: same body, step gate
 * `ntimestep % N` instead of `ntimestep % N - 2`); takes the ex_unload slot (unfix: LE_FIX_EX_UNLOAD) */
int le_fix_bond_break(le_ctx *c, int nevery, int btype, double rmax, double prob, int seed);
/* fix ID all bond/create N itype jtype Rmin bondtype iparam imax inew jparam jmax jnew prob f seed (src/MC/fix_bond_create.cpp:41-200,
 * 349-629, the ancestor of ex_load: events on `ntimestep % N == 0`, partner = the closest eligible listed neighbor within Rmin
 * without the loop-extrusion rules, bond counts taken once in the first run's setup); takes the ex_load slot (unfix:
 * LE_FIX_EX_LOAD; thermo f_ID[k] = slot 2).  inew / jnew < 1 = keep the type.  newton_bond must be off. */
int le_fix_bond_create(le_ctx *c, int nevery, int itype, int jtype, double rmin, int btype, double prob,
                       int seed, int imaxbond, int inewtype, int jmaxbond, int jnewtype);
int le_unfix(le_ctx *c, int which);                         /* unfix */

/* ---- atoms and topology ---------------------------------------------------------------- */
/* x[N*3], v[N*3] (may be NULL = zero), image[N] LAMMPS-packed (may be NULL = 0), all in tag order */
int le_upload_atoms(le_ctx *c, int n, const int *tag, const int *type, const double *x,
                    const double *v, const int *image);
/* read_data "Bonds" section: each bond once; stored on both atoms in file order, then
 * the 1-2/1-3/1-4 special lists are built as Special::build does (src/special.cpp:55-154). */
int le_upload_bonds(le_ctx *c, int nbonds, const int *btype, const int *atom1, const int *atom2);
/* raw per-atom tables exactly as the reference holds them (state replay, read_restart); nspecial == special == NULL:
 * the special lists are built from the bond tables as read_restart does (Special::build) */
int le_upload_topology(le_ctx *c, const int *num_bond, const int *bond_type, const int *bond_atom,
                       const int *nspecial, const int *special);
/* overwrite positions only (x[N*3], image may be NULL = keep); neighbor/bond lists are NOT rebuilt,
 * which is the state a USER-LE fix sees between reneighborings */
int le_set_positions(le_ctx *c, const double *x, const int *image);
int le_set_velocities(le_ctx *c, const double *v);

/* ---- run ------------------------------------------------------------------------------- */
int le_run(le_ctx *c, int64_t nsteps);                      /* run N */
int le_force_rebuild(le_ctx *c);                            /* Neighbor::build(1) now */
/* run N start S stop E (src/run.cpp:90-120): the next le_run calls are segments of one run S..E -- fix langevin's temperature
 * ramp spans S..E instead of restarting per segment (a front end that cuts a run at dump steps); stop <= start: off */
int le_set_run_span(le_ctx *c, int64_t start, int64_t stop);
/* le_run with direct launches and an event before every launch; *kstep_us = average duration of the plain step
 * kernel (launch to next launch on the stream), for live roofline measurements */
int le_run_timed(le_ctx *c, int64_t nsteps, double *kstep_us);
/* name of the plain step kernel le_run launches in the current configuration (for reports; e.g. "k_step3<0,0,1>") */
const char *le_step_kernel_name(le_ctx *c);
/* run one USER-LE fix's post_integrate on the current state, regardless of the step gate */
int le_run_le_event(le_ctx *c, int which);
/* skip n draws of that fix's Marsaglia stream / re-seed it (state replay) */
int le_fix_rng_reset(le_ctx *c, int which, int seed, int64_t ndraws_consumed);
int le_fix_rng_consumed(le_ctx *c, int which, int64_t *ndraws);
/* hand a fix's Marsaglia generator over as the reference holds it, and take it back: state[103] in the layout of
 * RanMars::get_state / set_state (src/random_mars.cpp:297-319).  A host that owns the RanMars object of `fix extrusion` /
 * `ex_load` / `ex_unload` (the run_style binding, lammps_le_b200/lammps_style) continues the SAME stream on the device. */
int le_fix_rng_set_state(le_ctx *c, int which, const double *state103);
int le_fix_rng_get_state(le_ctx *c, int which, double *state103);
/* compute forces/energies at the current positions without integrating (run 0 without fixes):
 * f[N*3] conservative pair+bond force in tag order (may be NULL) */
int le_compute_forces(le_ctx *c, double *f, le_thermo *out);
/* the same forces from the plain (no energy / virial tally) instantiation of the step kernel, the one production
 * timesteps run; f[N*3] */
int le_compute_forces_plain(le_ctx *c, double *f);

/* `minimize etol ftol maxiter maxeval` (src/minimize.cpp:31-60) with min_style cg and the quadratic line search, the
 * reference's defaults (MinCG::iterate src/min_cg.cpp:35-200, MinLineSearch::linemin_quadratic src/min_linesearch.cpp:325-505);
 * the numbers Finish::end prints under "Minimization stats" (src/finish.cpp:192-218) come back in *out (may be NULL).
 * Energies per atom (thermo_modify norm yes).  Velocities are not touched; the timestep advances by the iterations. */
typedef struct le_min_result {
  int stop;                 /* index into Min::stopstrings (src/min.cpp:1058-1073): 0 max iterations, 1 max force evaluations,
                               2 energy tolerance, 3 force tolerance, 4 not downhill, 5 alpha is zero, 6 forces are zero, 7 quadratic factors */
  int niter, neval;
  double einitial, eprevious, efinal;
  double fnorm2_init, fnorm2_final, fnorminf_init, fnorminf_final, alpha_final;
} le_min_result;
int le_minimize(le_ctx *c, double etol, double ftol, int maxiter, int maxeval, le_min_result *out);
const char *le_min_stop_string(int stop);

/* ---- results --------------------------------------------------------------------------- */
int le_natoms(const le_ctx *c);
int64_t le_timestep(const le_ctx *c);
int le_download_x(le_ctx *c, double *x, int *image);        /* wrapped x[N*3] + image[N] */
int le_download_v(le_ctx *c, double *v);
int le_download_types(le_ctx *c, int *type);
int le_download_topology(le_ctx *c, int *num_bond, int *bond_type, int *bond_atom, int *nspecial,
                         int *special);
/* bulk exchange of the atoms THIS GPU owns, in device order, with the double <-> fixed-point conversion done on
 * the GPU (the host only moves flat buffers; use pinned memory for full PCIe speed).  Arrays hold at least
 * le_local_capacity() entries; any output pointer may be NULL.  le_upload_owned addresses atoms by tag (they must
 * be owned by this GPU) and, like le_set_positions, does not rebuild lists. */
int le_local_capacity(const le_ctx *c);
int le_download_owned(le_ctx *c, int *n_out, int *tag, double *x, int *image, double *v);
int le_upload_owned(le_ctx *c, int n, const int *tag, const double *x, const int *image, const double *v);
/* half neighbor list of the last build, CSR over tags: offsets[N+1]; entries = partner tag | which<<30
 * (the reference's j ^ (which << SBBITS), src/npair_half_bin_newton.cpp:113, with j replaced by tag).
 * Call with entries == NULL to get the total count in *nentries. */
int le_download_neighlist(le_ctx *c, int half, int64_t *offsets, int *entries, int64_t *nentries);
/* neighbor->bondlist of the last build: rows (tag_i, tag_j, type), in the reference's order */
int le_download_bondlist(le_ctx *c, int *rows, int64_t *nrows);
int le_thermo_count(const le_ctx *c);
int le_get_thermo(const le_ctx *c, int index, le_thermo *out); /* index < 0 counts from the end */
int le_get_stats(le_ctx *c, le_stats *out);
/* radius of gyration of the whole system from unwrapped coordinates (compute gyration) */
int le_compute_rg(le_ctx *c, double *rg);
/* polymer observables tallied on the device over the atoms this GPU owns (multi-GPU: sum the outputs over ranks):
 *   rg_sums[5]   = sum m, sum m x, sum m y, sum m z, sum m |x|^2 of the UNWRAPPED coordinates
 *                  (compute gyration, src/compute_gyration.cpp:60-100: Rg^2 = S4/S0 - |S1..3/S0|^2)
 *   contacts[ns] = number of bead pairs (t, t + s_list[k]) closer than rc (minimum image; tags are chain positions)
 *   loop_hist[nbins] = histogram of b - a over the bonds (a < b) of type btype, bins of bin_width, last bin = overflow
 *                  (the loops; compute property/local batom1 batom2 btype, src/compute_property_local.cpp:104-117) */
int le_observables(le_ctx *c, int ns, const int *s_list, double rc, int btype, int nbins, int bin_width,
                   double *rg_sums, int64_t *contacts, int64_t *loop_hist);

/* ---- several GPUs of one box: spatial domain decomposition (Comm/CommBrick, src/comm_brick.cpp:452-876) -----------
 * One process (and one le_ctx) per GPU.  The box is cut into x-slabs of neighbor cells; each GPU owns one slab plus
 * `halo_distance` (>= the neighbor cutoff; pass the longest bond + 2 backbone bonds when USER-LE fixes are used) of
 * ghost cells on either side.  Ghost positions are stored by the owning GPU's integrator straight into the
 * neighbor's memory over NVLink (CUDA IPC peer mappings), atoms migrate at every reneighboring, bond / special
 * tables are replicated by tag so an extruder's bonds arrive with its beads.  Every rank makes the SAME sequence of
 * calls with the SAME (global) arguments; uploads take the whole system, downloads fill only the entries of the atoms
 * the GPU owns (the caller zero-fills and sums over ranks).
 *   le_dd_init        after le_create, before le_upload_atoms
 *   le_dd_get_handle  after le_upload_atoms: 64-byte CUDA IPC handle of this GPU's peer-visible arena
 *   le_dd_connect     handles of all ranks (all-gathered by the caller, e.g. torch.distributed), rank-major */
int le_dd_init(le_ctx *c, int rank, int nranks, double halo_distance);
/* slab cuts at upload: 1 (default) = where the cumulative atom count crosses r N / P (`balance 1.0 shift x`, src/balance.cpp),
 * 0 = equal-width slabs (the reference's bricks without a balance command).  `fix balance` (src/fix_balance.cpp:191-270) is
 * the host sequence DDEngine.rebalance(): imbalance factor from the owned counts, then a fresh context cut from the CURRENT
 * configuration with the whole state carried over. */
int le_dd_balance(le_ctx *c, int mode);
int le_dd_get_handle(le_ctx *c, void *handle64);
int le_dd_connect(le_ctx *c, const void *handles);
/* raw per-GPU tallies behind a thermo record / the last le_compute_forces: [16] = sum m v^2, evdwl, ebond,
 * virial[6], FENE warnings, ... (multi-GPU callers sum over ranks before normalising) */
int le_get_thermo_sums(const le_ctx *c, int index, double *out16);
int le_get_force_sums(const le_ctx *c, double *out16);

/* ---- synthetic inputs (host only; bench and tests) -------------------------------------------- */
/* self-avoiding walk(s) of n beads in a periodic cube [0,L)^3: bond length `step`, no two beads closer
 * than rmin; x[n*3] wrapped coordinates, image[n] LAMMPS-packed image flags (may be NULL) */
int le_gen_saw_chains(int n, int nchains, double L, double step, double rmin, uint64_t seed,
                      double *x, int *image);
/* FENE-melt start: nchains*len beads on a snake path through a simple-cubic lattice at density rho */
int le_gen_lattice_melt(int nchains, int len, double rho, double *L, double *x, int *image);

/* ---- `velocity all create T seed ...` (host only) ------------------------------------------------ */
/* ---- restart files (host only): the reference's binary format for this path -- write_restart / read_restart with units lj,
 * atom_style bond, pair lj/cut, bond fene | harmonic | hybrid (src/write_restart.cpp:205-600, src/read_restart.cpp, lmprestart.h).
 * The file carries the per-atom bond tables (extruder bonds included) but no special lists: after reading, hand the tables to
 * le_upload_topology with NULL specials (Special::build, as read_restart does).  Pair matrices are indexed (i-1)*ntypes + (j-1),
 * i <= j, as PairLJCut::write_restart stores them; bond_style hybrid stores its sub-style names only (BondHybrid::write_restart). */
#define LE_RESTART_MAXT 8
typedef struct le_restart_header {
  int64_t ntimestep, natoms, nbonds;
  int ntypes, nbondtypes, bond_per_atom, extra_bond_per_atom, maxspecial;
  int newton_pair, newton_bond, periodic[3], nprocs_file, atom_sortfreq;
  double boxlo[3], boxhi[3], special_lj[3], dt, comm_cutoff;
  double mass[LE_RESTART_MAXT];
  char version[32], units[16], atom_style[16], pair_style[32], bond_style[32];
  double cut_global; int offset_flag, mix_flag, tail_flag;
  int pair_setflag[LE_RESTART_MAXT * LE_RESTART_MAXT];
  double pair_eps[LE_RESTART_MAXT * LE_RESTART_MAXT], pair_sigma[LE_RESTART_MAXT * LE_RESTART_MAXT], pair_cut[LE_RESTART_MAXT * LE_RESTART_MAXT];
  int bond_coeffs_stored;
  double bond_k[LE_RESTART_MAXT], bond_r0[LE_RESTART_MAXT], bond_eps[LE_RESTART_MAXT], bond_sigma[LE_RESTART_MAXT];
  int nhybrid; char hybrid_styles[4][16];
} le_restart_header;
int le_host_restart_read_header(const char *path, le_restart_header *h, char *err, int errlen);
/* atoms in file order; bond_type / bond_atom are [natoms][bond_per_atom]; any output may be NULL */
int le_host_restart_read_atoms(const char *path, int *tag, int *type, int *image, int *molecule, double *x, double *v,
                               int *num_bond, int *bond_type, int *bond_atom, char *err, int errlen);
int le_host_restart_write(const char *path, const le_restart_header *h, const int *tag, const int *type, const int *image,
                          const int *molecule, const double *x, const double *v, const int *num_bond, const int *bond_type,
                          const int *bond_atom, char *err, int errlen);

/* Velocity::create (src/velocity.cpp:162-401) for all atoms of a 3-d system in tag order: Park-Miller draws
 * (src/random_park.cpp) in the reference's order, 1/sqrt(mass) scaling, `mom yes` momentum zeroing, rescale to t_desired
 * with 3N-3 degrees of freedom (compute temp).  loop: 0 = all, 1 = local (one rank), 2 = geom (x[n*3] needed: the generator
 * is re-seeded from each atom's coordinates).  dist: 0 = uniform, 1 = gaussian.  `rot yes`, `sum yes`, `bias yes` and other
 * temperature computes are not provided.  v[n*3] receives the result; feed it to le_set_velocities. */
int le_host_velocity_create(int n, const int *type, const double *mass_per_type, const double *x, double t_desired, int seed,
                            int dist, int mom, int loop, double *v);

/* ---- `compute property/local batom1 batom2 btype` (host only) ------------------------------------ */
/* ComputePropertyLocal::count_bonds / pack (src/compute_property_local.cpp:463-493): the bonds in the order the
 * reference lists them -- atoms in tag order, their bond slots in slot order, with newton_bond off only from the end with
 * the smaller tag, deleted bonds (type 0) skipped.  rows[3*k..] = batom1, batom2, btype; returns the number of rows
 * (rows may be NULL to count). */
int64_t le_host_property_local_bonds(int n, int bpa, const int *num_bond, const int *bond_type, const int *bond_atom,
                                     int newton_bond, int *rows);

#ifdef __cplusplus
}
#endif
#endif /* LE_B200_H */
