// ref_harness.cpp -- state-capture driver around the UNMODIFIED reference library (oracle/_ref).
//
// TEST INFRASTRUCTURE ONLY: built by oracle/build_ref.py into oracle/_ref/ref_harness, run by tests/,
// by oracle/make_golden.py and by bench.py's reference arm.  The product never links or runs it.
//
// It is the reference's own main() (src/main.cpp:36-68) plus two additions made from OUTSIDE the library:
//   * an extra fix style `le/snap`, registered in Modify::fix_map at run time, whose post_integrate hook
//     (a) optionally snaps every coordinate to the engine's 32-bit fixed-point grid, so that both codes
//         see bit-identical positions, and
//     (b) writes the complete per-atom topology state (bond tables, special lists, types), the bond list,
//         the pair neighbor list, Neighbor::xhold and the Marsaglia counters of the three USER-LE fixes
//         immediately BEFORE (instance defined ahead of the USER-LE fixes) or AFTER (instance defined
//         behind them) every USER-LE event;
//   * `-final FILE`: after the input script ends, one record with forces, energies and virials.
// Private members are reached with a test-only `#define private public`; nothing in the reference is edited.
//
// usage:  ref_harness -in deck [-final out.bin] [other lmp options]
//         ref_harness -ranmars SEED N -log none        (prints N draws of RanMars(SEED))
//         deck line:  fix ID all le/snap <file> <pre|post> [grid]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <mpi.h>

#define private public
#define protected public
#include "atom.h"
#include "bond.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "fix.h"
#include "fix_ex_load.h"
#include "fix_ex_unload.h"
#include "fix_extrusion.h"
#include "fix_bond_create.h"      // src/MC: the ancestors of fix ex_load / fix ex_unload (compiled in by oracle/build_ref.py, EXTRA_STYLES)
#include "fix_bond_break.h"
#include "force.h"
#include "input.h"
#include "lammps.h"
#include "modify.h"
#include "neigh_list.h"
#include "neighbor.h"
#include "pair.h"
#include "random_mars.h"
#include "update.h"
#undef private
#undef protected

using namespace LAMMPS_NS;

namespace {

struct Header {
  int32_t magic, kind, which, n, bpa, maxspecial, nbondlist, has_force;
  int64_t step, nneigh;
  double boxlo[3], boxhi[3];
  int32_t rngc[3];       // RanMars::c * 2^24 of fix extrusion / ex_unload / ex_load (-1: fix absent)
  int32_t counters[4];   // f_extrusion[1], f_unload[1], f_load[1], atom->nbonds
  double energy[2];      // eng_vdwl, bond energy
  double virial[12];     // pair virial[6], bond virial[6]
  double temp_ke;        // sum m v^2
};

template <class T> void put(FILE *f, const std::vector<T> &v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }

Fix *find_fix(LAMMPS *lmp, const char *style) {
  for (int i = 0; i < lmp->modify->nfix; i++)
    if (strcmp(lmp->modify->fix[i]->style, style) == 0) return lmp->modify->fix[i];
  return nullptr;
}

int rng_c(RanMars *r) { return r ? (int)llround(r->c * 16777216.0) : -1; }

// which USER-LE fix fires on this timestep (0 none, 1 extrusion, 2 unload / fix bond/break, 3 load / fix bond/create)
int firing(LAMMPS *lmp) {
  const bigint n = lmp->update->ntimestep;
  int which = 0, count = 0;
  if (Fix *f = find_fix(lmp, "extrusion")) if (n % f->nevery - 1 == 0) { which = 1; count++; }
  if (Fix *f = find_fix(lmp, "ex_unload")) if (n % f->nevery - 2 == 0) { which = 2; count++; }
  if (Fix *f = find_fix(lmp, "ex_load")) if (n % f->nevery - 3 == 0) { which = 3; count++; }
  if (Fix *f = find_fix(lmp, "bond/create")) if (n % f->nevery == 0) { which = 3; count++; }   // (records as fix 3: it is ex_load's ancestor)
  if (Fix *f = find_fix(lmp, "bond/break")) if (n % f->nevery == 0) { which = 2; count++; }    // (records as fix 2: ex_unload's ancestor)
  if (count > 1) lmp->error->all(FLERR, "le/snap: two USER-LE fixes fire on the same step; choose other periods");
  return which;
}

void write_record(LAMMPS *lmp, FILE *fp, int kind, int which, bool lists, bool forces) {
  Atom *atom = lmp->atom;
  const int n = atom->nlocal;
  const int bpa = atom->bond_per_atom, ms = atom->maxspecial;
  Header h;
  memset(&h, 0, sizeof h);
  h.magic = 0x4C455331; h.kind = kind; h.which = which; h.n = n; h.bpa = bpa; h.maxspecial = ms;
  h.step = lmp->update->ntimestep;
  for (int k = 0; k < 3; k++) { h.boxlo[k] = lmp->domain->boxlo[k]; h.boxhi[k] = lmp->domain->boxhi[k]; }
  FixExtrusion *fe = (FixExtrusion *)find_fix(lmp, "extrusion");
  FixExUnload *fu = (FixExUnload *)find_fix(lmp, "ex_unload");
  FixExLoad *fl = (FixExLoad *)find_fix(lmp, "ex_load");
  h.rngc[0] = fe ? rng_c(fe->random) : -1;
  FixBondBreak *fb = (FixBondBreak *)find_fix(lmp, "bond/break");
  h.rngc[1] = fu ? rng_c(fu->random) : fb ? rng_c(fb->random) : -1;
  FixBondCreate *fc = (FixBondCreate *)find_fix(lmp, "bond/create");
  h.rngc[2] = fl ? rng_c(fl->random) : fc ? rng_c(fc->random) : -1;
  h.counters[0] = fe ? fe->breakcount : 0;
  h.counters[1] = fu ? fu->breakcount : fb ? fb->breakcount : 0;
  h.counters[2] = fl ? fl->createcount : fc ? fc->createcount : 0;
  h.counters[3] = (int)atom->nbonds;
  h.has_force = forces ? 1 : 0;
  if (forces) {
    if (lmp->force->pair) { h.energy[0] = lmp->force->pair->eng_vdwl; for (int k = 0; k < 6; k++) h.virial[k] = lmp->force->pair->virial[k]; }
    if (lmp->force->bond) { h.energy[1] = lmp->force->bond->energy; for (int k = 0; k < 6; k++) h.virial[6 + k] = lmp->force->bond->virial[k]; }
  }
  std::vector<double> x((size_t)n * 3), xh((size_t)n * 3, 0.0), f((size_t)n * 3, 0.0), v((size_t)n * 3, 0.0);
  std::vector<int32_t> img(n), type(n), nb(n), bt((size_t)n * bpa, 0), ba((size_t)n * bpa, 0), ns((size_t)n * 3), sp((size_t)n * ms, 0);
  tagint *tag = atom->tag;
  double ke = 0.0;
  for (int i = 0; i < n; i++) {
    const size_t t = (size_t)tag[i] - 1;
    for (int k = 0; k < 3; k++) {
      x[3 * t + k] = atom->x[i][k];
      f[3 * t + k] = atom->f[i][k];
      v[3 * t + k] = atom->v[i][k];
      if (lmp->neighbor->xhold && i < lmp->neighbor->maxhold) xh[3 * t + k] = lmp->neighbor->xhold[i][k];
    }
    const double m = atom->mass[atom->type[i]];
    ke += m * (atom->v[i][0] * atom->v[i][0] + atom->v[i][1] * atom->v[i][1] + atom->v[i][2] * atom->v[i][2]);
    img[t] = (int32_t)atom->image[i];
    type[t] = atom->type[i];
    nb[t] = atom->num_bond[i];
    for (int m2 = 0; m2 < atom->num_bond[i]; m2++) { bt[t * bpa + m2] = atom->bond_type[i][m2]; ba[t * bpa + m2] = atom->bond_atom[i][m2]; }
    for (int k = 0; k < 3; k++) ns[3 * t + k] = atom->nspecial[i][k];
    for (int k = 0; k < atom->nspecial[i][2]; k++) sp[t * ms + k] = atom->special[i][k];
  }
  h.temp_ke = ke;
  std::vector<int32_t> bl;
  std::vector<int64_t> off;
  std::vector<int32_t> ent;
  if (lists) {
    Neighbor *nbr = lmp->neighbor;
    for (int k = 0; k < nbr->nbondlist; k++) {
      bl.push_back(tag[nbr->bondlist[k][0]]); bl.push_back(tag[nbr->bondlist[k][1]]); bl.push_back(nbr->bondlist[k][2]);
    }
    h.nbondlist = nbr->nbondlist;
    NeighList *list = lmp->force->pair ? lmp->force->pair->list : nullptr;
    if (list) {
      std::vector<std::vector<int32_t>> rows(n);
      for (int ii = 0; ii < list->inum; ii++) {
        const int i = list->ilist[ii];
        std::vector<int32_t> &r = rows[tag[i] - 1];
        for (int jj = 0; jj < list->numneigh[i]; jj++) {
          const int j = list->firstneigh[i][jj];
          r.push_back((int32_t)(tag[j & NEIGHMASK] | ((j >> SBBITS) << SBBITS)));
        }
      }
      off.resize(n + 1);
      int64_t o = 0;
      for (int t = 0; t < n; t++) { off[t] = o; o += (int64_t)rows[t].size(); ent.insert(ent.end(), rows[t].begin(), rows[t].end()); }
      off[n] = o;
      h.nneigh = o;
    }
  }
  fwrite(&h, sizeof h, 1, fp);
  put(fp, x); put(fp, xh); put(fp, v); put(fp, f);
  put(fp, img); put(fp, type); put(fp, nb); put(fp, bt); put(fp, ba); put(fp, ns); put(fp, sp);
  put(fp, bl);
  if (lists && !off.empty()) { put(fp, off); put(fp, ent); }
  fflush(fp);
}

// snap every owned coordinate to the engine's grid: x = lo + u*scale (+ w*L outside the box), u = rint(frac*2^32)
void snap_to_grid(LAMMPS *lmp) {
  Atom *atom = lmp->atom;
  const double two32 = 4294967296.0;
  for (int k = 0; k < 3; k++) {
    const double lo = lmp->domain->boxlo[k], L = lmp->domain->boxhi[k] - lmp->domain->boxlo[k];
    const double scale = L / two32;
    for (int i = 0; i < atom->nlocal; i++) {
      const double fr = (atom->x[i][k] - lo) / L;
      double w = floor(fr);
      double u = rint((fr - w) * two32);
      if (u >= two32) { u -= two32; w += 1.0; }
      double xq = lo + u * scale;
      if (w != 0.0) xq = xq + w * L;
      atom->x[i][k] = xq;
    }
  }
}

class FixLeSnap : public Fix {
 public:
  FILE *fp;
  int post, grid;
  FixLeSnap(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg), fp(nullptr), post(0), grid(0) {
    if (narg < 5) error->all(FLERR, "Illegal fix le/snap command");
    fp = fopen(arg[3], "wb");
    if (!fp) error->all(FLERR, "fix le/snap: cannot open output file");
    post = strcmp(arg[4], "post") == 0;
    for (int k = 5; k < narg; k++) if (strcmp(arg[k], "grid") == 0) grid = 1;
  }
  ~FixLeSnap() { if (fp) fclose(fp); }
  int setmask() { return FixConst::POST_INTEGRATE | FixConst::PRE_EXCHANGE; }
  // Verlet::setup calls this before pbc/exchange/borders and the first neighbor build (src/verlet.cpp:103-117),
  // so Neighbor::xhold and the step-0 forces are taken on grid positions too
  void setup_pre_exchange() { if (grid && !post) snap_to_grid(lmp); }
  void pre_exchange() {}
  void post_integrate() {
    if (grid && !post) snap_to_grid(lmp);
    const int which = firing(lmp);
    if (!which) return;
    write_record(lmp, fp, post ? 1 : 0, which, !post, false);
  }
};

Fix *make_snap(LAMMPS *lmp, int narg, char **arg) { return new FixLeSnap(lmp, narg, arg); }

}  // namespace

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  const char *final_file = nullptr;
  int grid_final = 0;
  int mars_seed = 0, mars_n = 0;
  std::vector<char *> args;
  for (int k = 0; k < argc; k++) {
    if (strcmp(argv[k], "-ranmars") == 0 && k + 2 < argc) { mars_seed = atoi(argv[++k]); mars_n = atoi(argv[++k]); continue; }
    if (strcmp(argv[k], "-final") == 0 && k + 1 < argc) { final_file = argv[++k]; continue; }
    if (strcmp(argv[k], "-gridfinal") == 0) { grid_final = 1; continue; }
    args.push_back(argv[k]);
  }
  int rc = 0;
  try {
    LAMMPS *lammps = new LAMMPS((int)args.size(), args.data(), MPI_COMM_WORLD);
    (*lammps->modify->fix_map)["le/snap"] = &make_snap;
    if (mars_seed > 0) {
      // known-answer vectors of the reference's own RanMars (src/random_mars.cpp): exact doubles as hex floats
      RanMars rm(lammps, mars_seed);
      for (int k = 0; k < mars_n; k++) printf("RANMARS %a\n", rm.uniform());
      delete lammps;
      MPI_Finalize();
      return 0;
    }
    lammps->input->file();
    if (final_file) {
      (void)grid_final;
      FILE *fp = fopen(final_file, "wb");
      if (!fp) { fprintf(stderr, "cannot open %s\n", final_file); rc = 2; }
      else { write_record(lammps, fp, 2, 0, true, true); fclose(fp); }
    }
    delete lammps;
  } catch (LAMMPSAbortException &ae) {
    fprintf(stderr, "LAMMPS abort: %s\n", ae.message.c_str());
    rc = 1;
  } catch (LAMMPSException &e) {
    fprintf(stderr, "LAMMPS error: %s\n", e.message.c_str());
    rc = 1;
  }
  MPI_Finalize();
  return rc;
}
