/* run_style le/b200 (see verlet_le_b200.h).  reference counterparts: Verlet::init/setup/run (src/verlet.cpp:52-354),
   Run::command (src/run.cpp:176-237), Thermo's computes (src/compute_pe.cpp, compute_pressure.cpp, compute_temp.cpp). */

// The settings of a few styles are private members with no accessor (BondFENE epsilon/sigma, FixLangevin t_start/t_period/
// seed, the USER-LE fixes' arguments and counters).  A maintainer would add `friend class VerletLEB200;` (or extract()
// keys) to those headers; this translation unit, which must build against the UNMODIFIED reference tree, opens them instead.
#define private public
#define protected public
#include "pair.h"
#include "bond_fene.h"
#include "bond_harmonic.h"
#include "bond_hybrid.h"
#include "angle_cosine.h"
#include "fix_langevin.h"
#include "fix_nve_limit.h"
#include "fix_extrusion.h"
#include "fix_ex_load.h"
#include "fix_ex_unload.h"
#ifdef LE_B200_WITH_MC          // the MC package is installed: the ancestors of ex_load / ex_unload run on the engine too
#include "fix_bond_create.h"
#include "fix_bond_break.h"
#endif
#include "random_mars.h"
#undef private
#undef protected

#include "verlet_le_b200.h"
#include "angle.h"
#include "atom.h"
#include "comm.h"
#include "compute.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "modify.h"
#include "neighbor.h"
#include "output.h"
#include "pair.h"
#include "timer.h"
#include "update.h"
#include "utils.h"
#include <cmath>
#include <cstring>
#include <vector>

extern "C" {
#include "le_b200.h"
}

using namespace LAMMPS_NS;

VerletLEB200::VerletLEB200(LAMMPS *lmp, int narg, char **arg) : Integrate(lmp, narg, arg), ctx(nullptr), device(0), has_le(0),
  fix_ext(nullptr), fix_load(nullptr), fix_unload(nullptr)
{
  // run_style le/b200 [device N]
  for (int i = 0; i < narg; i++)
    if (strcmp(arg[i], "device") == 0 && i + 1 < narg) device = utils::inumeric(FLERR, arg[++i], false, lmp);
}

VerletLEB200::~VerletLEB200() { if (ctx) le_destroy(ctx); }

void VerletLEB200::check(int rc)
{
  if (rc) error->all(FLERR, ctx ? le_last_error(ctx) : "run_style le/b200: no CUDA device");
}

void VerletLEB200::init()
{
  Integrate::init();
  if (comm->nprocs != 1) error->all(FLERR, "run_style le/b200: one MPI rank per engine context (several GPUs: le_dd_init, INTEGRATION.md section 5)");
  if (domain->triclinic) error->all(FLERR, "run_style le/b200 requires an orthogonal box");
  if (strcmp(atom->atom_style, "bond") != 0 && strcmp(atom->atom_style, "angle") != 0)
    error->all(FLERR, "run_style le/b200 requires atom_style bond or angle");
  if (force->angle && strcmp(force->angle_style, "cosine") != 0) error->all(FLERR, "run_style le/b200 supports angle_style cosine only");
  if (force->dihedral || force->improper || force->kspace) error->all(FLERR, "run_style le/b200 supports pair lj/cut + bond fene/harmonic + angle cosine only");
}

/* translate the script's settings; called at every setup() because fixes and coefficients may change between runs */
void VerletLEB200::create_context()
{
  if (ctx) { le_destroy(ctx); ctx = nullptr; }
  int per[3] = {domain->xperiodic, domain->yperiodic, domain->zperiodic};
  check(le_create(&ctx, device, domain->boxlo, domain->boxhi, per));
  const int nt = atom->ntypes;
  check(le_set_types(ctx, nt, atom->mass + 1, atom->nbondtypes));

  // pair_style lj/cut: mixed coefficients through Pair::extract, cutoffs through Pair::init_one's table
  Pair *p = force->pair_match("lj/cut", 1);
  if (!p) error->all(FLERR, "run_style le/b200 requires pair_style lj/cut");
  int dim;
  double **eps = (double **) p->extract("epsilon", dim), **sig = (double **) p->extract("sigma", dim);
  std::vector<double> e(nt * nt), s(nt * nt), cu(nt * nt);
  for (int i = 1; i <= nt; i++)
    for (int j = 1; j <= nt; j++) {
      const int a = i <= j ? i : j, b = i <= j ? j : i;                  // the tables are filled for i <= j
      e[(i - 1) * nt + j - 1] = eps[a][b]; s[(i - 1) * nt + j - 1] = sig[a][b]; cu[(i - 1) * nt + j - 1] = sqrt(p->cutsq[a][b]);
    }
  check(le_set_pair_lj(ctx, nt, e.data(), s.data(), cu.data(), p->offset_flag));

  // bond_style fene | harmonic | hybrid of the two
  if (force->bond) {
    auto one = [&](Bond *b, const char *style, int t) {
      if (strcmp(style, "fene") == 0) {
        BondFENE *f = (BondFENE *) b;
        const double prm[4] = {f->k[t], f->r0[t], f->epsilon[t], f->sigma[t]};
        check(le_set_bond(ctx, t, LE_BOND_FENE, prm));
      } else if (strcmp(style, "harmonic") == 0) {
        BondHarmonic *h = (BondHarmonic *) b;
        const double prm[4] = {h->k[t], h->r0[t], 0.0, 0.0};
        check(le_set_bond(ctx, t, LE_BOND_HARMONIC, prm));
      } else error->all(FLERR, "run_style le/b200 supports bond_style fene and harmonic");
    };
    if (strcmp(force->bond_style, "hybrid") == 0) {
      BondHybrid *h = (BondHybrid *) force->bond;
      for (int t = 1; t <= atom->nbondtypes; t++) { const int m = h->map[t]; if (m >= 0) one(h->styles[m], h->keywords[m], t); }
    } else
      for (int t = 1; t <= atom->nbondtypes; t++) one(force->bond, force->bond_style, t);
  }
  // angle_style cosine
  if (force->angle) {
    AngleCosine *ac = (AngleCosine *) force->angle;
    check(le_set_angle_types(ctx, atom->nangletypes));
    for (int t = 1; t <= atom->nangletypes; t++) { const double prm[4] = {ac->k[t], 0.0, 0.0, 0.0}; check(le_set_angle(ctx, t, LE_ANGLE_COSINE_STYLE, prm)); }
  }
  check(le_set_special(ctx, force->special_lj + 1));
  check(le_set_newton(ctx, force->newton_pair, force->newton_bond));
  check(le_set_neighbor(ctx, neighbor->skin, neighbor->every, neighbor->delay, neighbor->dist_check));
  if (neighbor->oneatom != 2000) check(le_set_neighbor_capacity(ctx, neighbor->oneatom));   // neigh_modify one (2000 = the unset default)
  check(le_set_timestep(ctx, update->dt));
  check(le_thermo_every(ctx, 0));

  // fixes, in the order Modify calls them (the USER-LE fixes do nothing under newton_bond on, SURVEY.md section 0: the
  // engine refuses that combination with its own message)
  has_le = 0; fix_ext = fix_load = fix_unload = nullptr; load_is_mc = unload_is_mc = 0;
  for (int i = 0; i < modify->nfix; i++) {
    Fix *f = modify->fix[i];
    if (strcmp(f->style, "nve") == 0) check(le_fix_nve(ctx, 1));
    else if (strcmp(f->style, "nve/limit") == 0) {
      FixNVELimit *l = (FixNVELimit *) f;
      check(le_fix_nve_limit(ctx, l->xlimit));
    } else if (strcmp(f->style, "langevin") == 0) {
      FixLangevin *l = (FixLangevin *) f;
      check(le_fix_langevin(ctx, l->t_start, l->t_stop, l->t_period, l->seed));
    } else if (strcmp(f->style, "extrusion") == 0) {
      FixExtrusion *x = (FixExtrusion *) f;
      check(le_fix_extrusion(ctx, x->nevery, x->neutral_type, x->ctcf_left, x->ctcf_right, x->through_prob, x->btype, x->ctcf_left_right, 0));
      has_le = 1; fix_ext = f;
    } else if (strcmp(f->style, "ex_load") == 0) {
      FixExLoad *x = (FixExLoad *) f;
      check(le_fix_ex_load(ctx, x->nevery, x->iatomtype, x->jatomtype, sqrt(x->cutsq), x->btype, x->fraction, 12345,
                           x->imaxbond, x->inewtype, x->jmaxbond, x->jnewtype));
      has_le = 1; fix_load = f;
    } else if (strcmp(f->style, "ex_unload") == 0) {
      FixExUnload *x = (FixExUnload *) f;
      check(le_fix_ex_unload(ctx, x->nevery, x->btype, sqrt(x->cutsq), x->fraction, 12345));
      has_le = 1; fix_unload = f;
#ifdef LE_B200_WITH_MC
    } else if (strcmp(f->style, "bond/create") == 0) {     // takes the engine's ex_load slot
      FixBondCreate *x = (FixBondCreate *) f;
      if (fix_load) error->all(FLERR, "run_style le/b200: one of fix ex_load / fix bond/create per run");
      if (x->atype || x->dtype || x->itype || x->constrainflag) error->all(FLERR, "run_style le/b200: fix bond/create without atype / dtype / itype / aconstrain");
      check(le_fix_bond_create(ctx, x->nevery, x->iatomtype, x->jatomtype, sqrt(x->cutsq), x->btype, x->fraction, 12345,
                               x->imaxbond, x->inewtype, x->jmaxbond, x->jnewtype));
      has_le = 1; fix_load = f; load_is_mc = 1;
    } else if (strcmp(f->style, "bond/break") == 0) {      // takes the engine's ex_unload slot
      FixBondBreak *x = (FixBondBreak *) f;
      if (fix_unload) error->all(FLERR, "run_style le/b200: one of fix ex_unload / fix bond/break per run");
      check(le_fix_bond_break(ctx, x->nevery, x->btype, sqrt(x->cutsq), x->fraction, 12345));
      has_le = 1; fix_unload = f; unload_is_mc = 1;
#endif
    } else error->all(FLERR, "run_style le/b200 does not support this fix style");
  }
}

// the Marsaglia generator / the counters of whichever fix sits in the ex_load and ex_unload slots
RanMars *VerletLEB200::load_rng()
{
#ifdef LE_B200_WITH_MC
  if (load_is_mc) return ((FixBondCreate *) fix_load)->random;
#endif
  return ((FixExLoad *) fix_load)->random;
}
RanMars *VerletLEB200::unload_rng()
{
#ifdef LE_B200_WITH_MC
  if (unload_is_mc) return ((FixBondBreak *) fix_unload)->random;
#endif
  return ((FixExUnload *) fix_unload)->random;
}

void VerletLEB200::push_state()
{
  const int n = atom->nlocal;
  if ((bigint) n != atom->natoms) error->all(FLERR, "run_style le/b200: all atoms must be on this rank");
  const int bpa = atom->bond_per_atom, ms = atom->maxspecial;
  std::vector<int> tag(n), type(n), img(n), nb(n), bt((size_t) n * bpa, 0), ba((size_t) n * bpa, 0), ns((size_t) n * 3), sp((size_t) n * ms, 0);
  std::vector<double> x((size_t) n * 3), v((size_t) n * 3);
  for (int i = 0; i < n; i++) {
    const int t = atom->tag[i] - 1;                                       // the engine's tables are in tag order
    if (t < 0 || t >= n) error->all(FLERR, "run_style le/b200 requires consecutive atom IDs");
    tag[t] = t + 1; type[t] = atom->type[i]; img[t] = atom->image[i];
    for (int q = 0; q < 3; q++) { x[3 * t + q] = atom->x[i][q]; v[3 * t + q] = atom->v[i][q]; }
    nb[t] = atom->num_bond[i];
    for (int m = 0; m < nb[t]; m++) { bt[(size_t) t * bpa + m] = atom->bond_type[i][m]; ba[(size_t) t * bpa + m] = atom->bond_atom[i][m]; }
    for (int q = 0; q < 3; q++) ns[3 * t + q] = atom->nspecial[i][q];
    for (int m = 0; m < ns[3 * t + 2]; m++) sp[(size_t) t * ms + m] = atom->special[i][m];
  }
  // data-file capacities; with newton_bond on Atom counts a bond on one atom only, the engine keeps it on both
  int cap_b = bpa;
  if (force->newton_bond) {
    std::vector<int> deg(n, 0);
    for (int t = 0; t < n; t++)
      for (int m = 0; m < nb[t]; m++) { deg[t]++; deg[ba[(size_t) t * bpa + m] - 1]++; }
    for (int t = 0; t < n; t++) if (deg[t] > cap_b) cap_b = deg[t];
  }
  check(le_set_capacity(ctx, cap_b, ms));
  check(le_upload_atoms(ctx, n, tag.data(), type.data(), x.data(), v.data(), img.data()));
  if (force->newton_bond) {
    // newton_bond on: Atom holds every bond once, on one of its two atoms (Atom::data_bonds, src/atom.cpp:1261-1278); the
    // engine takes the bond list and builds the special lists as Special::build does
    std::vector<int> b_t, b_1, b_2;
    for (int t = 0; t < n; t++)
      for (int m = 0; m < nb[t]; m++) { b_t.push_back(bt[(size_t) t * bpa + m]); b_1.push_back(t + 1); b_2.push_back(ba[(size_t) t * bpa + m]); }
    check(le_upload_bonds(ctx, (int) b_t.size(), b_t.data(), b_1.data(), b_2.data()));
  } else
    check(le_upload_topology(ctx, nb.data(), bt.data(), ba.data(), ns.data(), sp.data()));
  if (force->angle && atom->nangles > 0) {
    // every angle once: Atom holds it on all three atoms under newton_bond off, on the centre atom only under newton_bond on
    // (Atom::data_angles, src/atom.cpp:1297-1340) -- the centre's copy is there in both cases
    std::vector<int> at, a1, a2, a3;
    for (int i = 0; i < n; i++)
      for (int m = 0; m < atom->num_angle[i]; m++)
        if (atom->angle_atom2[i][m] == atom->tag[i]) {
          at.push_back(atom->angle_type[i][m]); a1.push_back(atom->angle_atom1[i][m]); a2.push_back(atom->angle_atom2[i][m]); a3.push_back(atom->angle_atom3[i][m]);
        }
    check(le_upload_angles(ctx, (int) at.size(), at.data(), a1.data(), a2.data(), a3.data()));
  }
  check(le_reset_timestep(ctx, update->ntimestep));
  // the fixes' Marsaglia generators move to the device in their current state (the constructors of ex_load / ex_unload keep
  // the seed in a local variable, fix_ex_unload.cpp:66, so the state is the only complete record) and come back in pull_state
  double st[103];
  if (fix_ext) { ((FixExtrusion *) fix_ext)->random->get_state(st); check(le_fix_rng_set_state(ctx, LE_FIX_EXTRUSION, st)); }
  if (fix_unload) { unload_rng()->get_state(st); check(le_fix_rng_set_state(ctx, LE_FIX_EX_UNLOAD, st)); }
  if (fix_load) { load_rng()->get_state(st); check(le_fix_rng_set_state(ctx, LE_FIX_EX_LOAD, st)); }
}

void VerletLEB200::pull_state(int forces)
{
  const int n = atom->nlocal;
  std::vector<double> x((size_t) n * 3), v((size_t) n * 3);
  std::vector<int> img(n);
  check(le_download_x(ctx, x.data(), img.data()));
  check(le_download_v(ctx, v.data()));
  for (int i = 0; i < n; i++) {
    const int t = atom->tag[i] - 1;
    for (int q = 0; q < 3; q++) { atom->x[i][q] = x[3 * t + q]; atom->v[i][q] = v[3 * t + q]; }
    atom->image[i] = img[t];
  }
  if (has_le) {
    const int bpa = atom->bond_per_atom, ms = atom->maxspecial;
    std::vector<int> type(n), nb(n), bt((size_t) n * bpa), ba((size_t) n * bpa), ns((size_t) n * 3), sp((size_t) n * ms);
    check(le_download_types(ctx, type.data()));
    check(le_download_topology(ctx, nb.data(), bt.data(), ba.data(), ns.data(), sp.data()));
    for (int i = 0; i < n; i++) {
      const int t = atom->tag[i] - 1;
      atom->type[i] = type[t];
      atom->num_bond[i] = nb[t];
      for (int m = 0; m < nb[t]; m++) { atom->bond_type[i][m] = bt[(size_t) t * bpa + m]; atom->bond_atom[i][m] = ba[(size_t) t * bpa + m]; }
      for (int q = 0; q < 3; q++) atom->nspecial[i][q] = ns[3 * t + q];
      for (int m = 0; m < ns[3 * t + 2]; m++) atom->special[i][m] = sp[(size_t) t * ms + m];
    }
  }
  double st[103];
  if (fix_ext) { check(le_fix_rng_get_state(ctx, LE_FIX_EXTRUSION, st)); ((FixExtrusion *) fix_ext)->random->set_state(st); }
  if (fix_unload) { check(le_fix_rng_get_state(ctx, LE_FIX_EX_UNLOAD, st)); unload_rng()->set_state(st); }
  if (fix_load) { check(le_fix_rng_get_state(ctx, LE_FIX_EX_LOAD, st)); load_rng()->set_state(st); }
  (void) forces;
}

/* Thermo's computes read the global tallies of the force styles (compute pe: Pair::eng_vdwl + Bond::energy,
   src/compute_pe.cpp:87-100; compute pressure: Pair::virial + Bond::virial, src/compute_pressure.cpp:322-343) and the
   fixes' compute_vector(); they are valid on the timestep recorded in update->eflag_global / vflag_global */
void VerletLEB200::publish_thermo(const le_thermo &t)
{
  const double n = (double) atom->natoms;
  if (force->pair) { force->pair->eng_vdwl = t.epair * n; force->pair->eng_coul = 0.0; for (int q = 0; q < 6; q++) force->pair->virial[q] = t.virial[q]; }
  if (force->bond) { force->bond->energy = (t.emol - t.eangle) * n; for (int q = 0; q < 6; q++) force->bond->virial[q] = 0.0; }
  if (force->angle) { force->angle->energy = t.eangle * n; for (int q = 0; q < 6; q++) force->angle->virial[q] = 0.0; }
  atom->nbonds = t.nbonds;
  update->eflag_global = update->vflag_global = update->ntimestep;
  if (fix_ext) { ((FixExtrusion *) fix_ext)->breakcount = (int) t.le_f1[0]; ((FixExtrusion *) fix_ext)->breakcounttotal = (int) t.le_f2[0]; }
#ifdef LE_B200_WITH_MC
  if (fix_unload && unload_is_mc) { ((FixBondBreak *) fix_unload)->breakcount = (int) t.le_f1[1]; ((FixBondBreak *) fix_unload)->breakcounttotal = (int) t.le_f2[1]; }
  if (fix_load && load_is_mc) { ((FixBondCreate *) fix_load)->createcount = (int) t.le_f1[2]; ((FixBondCreate *) fix_load)->createcounttotal = (int) t.le_f2[2]; }
#endif
  if (fix_unload && !unload_is_mc) { ((FixExUnload *) fix_unload)->breakcount = (int) t.le_f1[1]; ((FixExUnload *) fix_unload)->breakcounttotal = (int) t.le_f2[1]; }
  if (fix_load && !load_is_mc) { ((FixExLoad *) fix_load)->createcount = (int) t.le_f1[2]; ((FixExLoad *) fix_load)->createcounttotal = (int) t.le_f2[2]; }
}

void VerletLEB200::setup(int flag)
{
  if (comm->me == 0 && screen) {
    fputs("Setting up le/b200 run ...\n", screen);
    if (flag) fmt::print(screen, "  Unit style    : {}\n  Current step  : {}\n  Time step     : {}\n  Engine        : {}\n",
                         update->unit_style, update->ntimestep, update->dt, le_version());
  }
  update->setupflag = 1;
  atom->setup();
  domain->pbc();
  domain->reset_box();
  create_context();
  push_state();
  // forces, energies and virial of the current state (Verlet::setup computes them for the step-0 thermo line)
  std::vector<double> f((size_t) atom->nlocal * 3);
  le_thermo t;
  check(le_compute_forces(ctx, f.data(), &t));
  for (int i = 0; i < atom->nlocal; i++) { const int k = atom->tag[i] - 1; for (int q = 0; q < 3; q++) atom->f[i][q] = f[3 * k + q]; }
  publish_thermo(t);
  for (int i = 0; i < modify->ncompute; i++) modify->compute[i]->setup();      // (Modify::setup without the fixes' setup(): degrees of freedom of compute temp)
  output->setup(flag);
  update->setupflag = 0;
}

void VerletLEB200::setup_minimal(int)
{
  update->setupflag = 1;
  create_context();
  push_state();
  update->setupflag = 0;
}

void VerletLEB200::run(int n)
{
  // the engine runs whole segments on the device; LAMMPS sees the state again wherever it has output to write
  bigint left = n;
  while (left > 0) {
    bigint seg = output->next - update->ntimestep;
    if (seg <= 0 || seg > left) seg = left;
    check(le_run(ctx, (int64_t) seg));
    update->ntimestep += seg;
    left -= seg;
    le_thermo t;
    check(le_get_thermo(ctx, -1, &t));
    pull_state(0);
    publish_thermo(t);
    if (update->ntimestep == output->next) {
      timer->stamp();
      output->write(update->ntimestep);
      timer->stamp(Timer::OUTPUT);
    }
  }
}

void VerletLEB200::cleanup()
{
  if (!ctx) return;
  le_stats st;
  if (le_get_stats(ctx, &st) == LE_OK) { neighbor->ncalls = (int) st.neigh_builds; neighbor->ndanger = (int) st.dangerous_builds; }
}
