/* run_style le/b200 -- the LAMMPS-side binding of libleb200.so (include/le_b200.h): an integrator style that hands the
   whole timestep loop of a chromatin + loop-extrusion deck to the B200 engine, the way src/KOKKOS/verlet_kokkos.cpp
   replaces Verlet.  A maintainer drops this pair of files into src/USER-LE/ (the repository's tests build it against the
   reference sources).  Everything the engine needs is read from the objects the input
   script has already built (Force, Modify, Neighbor, Atom); the deck itself only gains the line `run_style le/b200`. */
#ifdef INTEGRATE_CLASS

IntegrateStyle(le/b200,VerletLEB200)

#else

#ifndef LMP_VERLET_LE_B200_H
#define LMP_VERLET_LE_B200_H

#include "integrate.h"

struct le_ctx;
struct le_thermo;

namespace LAMMPS_NS {

class VerletLEB200 : public Integrate {
 public:
  VerletLEB200(class LAMMPS *, int, char **);
  virtual ~VerletLEB200();
  virtual void init();
  virtual void setup(int flag);
  virtual void setup_minimal(int);
  virtual void run(int);
  virtual void cleanup();

 private:
  ::le_ctx *ctx;
  int device;
  int has_le;                      // a USER-LE fix is defined: bonds, specials and types come back after every segment
  class Fix *fix_ext, *fix_load, *fix_unload;
  int load_is_mc, unload_is_mc;      // the slot holds fix bond/create / fix bond/break (MC package) instead of the USER-LE fix
  class RanMars *load_rng();
  class RanMars *unload_rng();

  void check(int rc);
  void create_context();           // Force / Modify / Neighbor settings -> le_set_* / le_fix_*
  void push_state();               // Atom -> le_upload_atoms / le_upload_topology
  void pull_state(int forces);     // engine -> Atom (x, v, image, type, bond and special tables)
  void publish_thermo(const ::le_thermo &t);   // energies / virial / fix counters where Thermo's computes look for them
};

}

#endif
#endif
