// le_restart.cpp -- the reference's binary restart file for this path (host only, part of libleb200_host.so and libleb200.so).
//
//   write: WriteRestart::write / header / type_arrays / force_fields / file_layout (src/write_restart.cpp:205-420, 425-600),
//          Group::write_restart (src/group.cpp:693-712), Modify::write_restart (src/modify.cpp:1374-1417),
//          PairLJCut::write_restart (src/pair_lj_cut.cpp:575-630), BondFENE / BondHarmonic / BondHybrid::write_restart,
//          AtomVec::pack_restart with atom_style bond's "molecule num_bond bond_type bond_atom" (src/atom_vec.cpp:1468-1560,
//          src/MOLECULE/atom_vec_bond.cpp:45); integers travel as the bit pattern of a 64-bit integer inside a double (ubuf).
//   read:  ReadRestart::command / header / type_arrays / force_fields / file_layout (src/read_restart.cpp).
// The file carries the per-atom BOND tables but no special lists: the reader's caller rebuilds them (le_upload_topology with
// NULL specials = Special::build, as src/read_restart.cpp:520-530 does).  One file, one "proc" section (a file the reference
// wrote on several ranks has several PERPROC sections: they are read one after the other).
#include "../../include/le_b200.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {
const char MAGIC[] = "LammpS RestartT";                  // lmprestart.h
enum { ENDIAN_OK = 0x0001, ENDIAN_SWAP = 0x1000, FORMAT_REVISION = 2 };
enum { VERSION, SMALLINT, TAGINT, BIGINT, UNITS, NTIMESTEP, DIMENSION, NPROCS, PROCGRID, NEWTON_PAIR, NEWTON_BOND, XPERIODIC, YPERIODIC,
       ZPERIODIC, BOUNDARY, ATOM_STYLE, NATOMS, NTYPES, NBONDS, NBONDTYPES, BOND_PER_ATOM, NANGLES, NANGLETYPES, ANGLE_PER_ATOM,
       NDIHEDRALS, NDIHEDRALTYPES, DIHEDRAL_PER_ATOM, NIMPROPERS, NIMPROPERTYPES, IMPROPER_PER_ATOM, TRICLINIC, BOXLO, BOXHI, XY, XZ, YZ,
       SPECIAL_LJ, SPECIAL_COUL, MASS, PAIR, BOND, ANGLE, DIHEDRAL, IMPROPER, MULTIPROC, MPIIO, PROCSPERFILE, PERPROC, IMAGEINT, BOUNDMIN,
       TIMESTEP, ATOM_ID, ATOM_MAP_STYLE, ATOM_MAP_USER, ATOM_SORTFREQ, ATOM_SORTBIN, COMM_MODE, COMM_CUTOFF, COMM_VEL, NO_PAIR,
       EXTRA_BOND_PER_ATOM, EXTRA_ANGLE_PER_ATOM, EXTRA_DIHEDRAL_PER_ATOM, EXTRA_IMPROPER_PER_ATOM, EXTRA_SPECIAL_PER_ATOM,
       ATOM_MAXSPECIAL, NELLIPSOIDS, NLINES, NTRIS, NBODIES };

int seterr(char *err, int errlen, const char *fmt, ...) {
  if (err && errlen > 0) { va_list ap; va_start(ap, fmt); vsnprintf(err, errlen, fmt, ap); va_end(ap); }
  return LE_EINVAL;
}
inline double as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }     // ubuf(int).d
inline long long as_int(double d) { long long v; memcpy(&v, &d, 8); return v; }      // ubuf(double).i

struct Writer {
  FILE *fp;
  void i(int v) { fwrite(&v, 4, 1, fp); }
  void fi(int flag, int v) { i(flag); i(v); }
  void fb(int flag, long long v) { i(flag); fwrite(&v, 8, 1, fp); }
  void fd(int flag, double v) { i(flag); fwrite(&v, 8, 1, fp); }
  void fs(int flag, const char *s) { const int n = (int)strlen(s) + 1; i(flag); i(n); fwrite(s, 1, n, fp); }
  void fiv(int flag, int n, const int *v) { i(flag); i(n); fwrite(v, 4, n, fp); }
  void fdv(int flag, int n, const double *v) { i(flag); i(n); fwrite(v, 8, n, fp); }
};

struct Reader {
  FILE *fp; bool bad = false;
  int i() { int v = 0; if (fread(&v, 4, 1, fp) != 1) bad = true; return v; }
  long long b() { long long v = 0; if (fread(&v, 8, 1, fp) != 1) bad = true; return v; }
  double d() { double v = 0; if (fread(&v, 8, 1, fp) != 1) bad = true; return v; }
  std::string s() { const int n = i(); if (bad || n < 0 || n > 4096) { bad = true; return ""; } std::string t(n, '\0'); if (n && fread(&t[0], 1, n, fp) != (size_t)n) bad = true; if (n) t.resize(n - 1); return t; }
  void iv(int *out, int want) { const int n = i(); for (int k = 0; k < n && !bad; k++) { const int v = i(); if (k < want) out[k] = v; } }
  void dv(double *out, int want) { const int n = i(); for (int k = 0; k < n && !bad; k++) { const double v = d(); if (k < want) out[k] = v; } }
};

void copy_str(char *dst, size_t cap, const std::string &s) { snprintf(dst, cap, "%s", s.c_str()); }

// everything before the per-proc atom sections; leaves the file positioned at the first PERPROC flag
int read_front(Reader &R, le_restart_header *h, char *err, int errlen) {
  char magic[16] = {0};
  if (fread(magic, 1, 16, R.fp) != 16 || strcmp(magic, MAGIC) != 0) return seterr(err, errlen, "Invalid LAMMPS restart file");
  const int endian = R.i();
  if (endian == ENDIAN_SWAP) return seterr(err, errlen, "Restart file byte ordering is swapped");
  if (endian != ENDIAN_OK) return seterr(err, errlen, "Restart file byte ordering is not recognized");
  const int rev = R.i();
  if (rev > FORMAT_REVISION) return seterr(err, errlen, "Restart file format revision incompatible with current LAMMPS version");
  memset(h, 0, sizeof *h);
  // header (flag, value) pairs up to -1
  for (int flag = R.i(); flag >= 0 && !R.bad; flag = R.i()) {
    switch (flag) {
      case VERSION: copy_str(h->version, sizeof h->version, R.s()); break;
      case SMALLINT: if (R.i() != 4) return seterr(err, errlen, "Smallint setting in lmptype.h is not compatible"); break;
      case IMAGEINT: if (R.i() != 4) return seterr(err, errlen, "Imageint setting in lmptype.h is not compatible"); break;
      case TAGINT: if (R.i() != 4) return seterr(err, errlen, "Tagint setting in lmptype.h is not compatible"); break;
      case BIGINT: if (R.i() != 8) return seterr(err, errlen, "Bigint setting in lmptype.h is not compatible"); break;
      case UNITS: copy_str(h->units, sizeof h->units, R.s()); break;
      case NTIMESTEP: h->ntimestep = R.b(); break;
      case DIMENSION: if (R.i() != 3) return seterr(err, errlen, "restart file is not three-dimensional"); break;
      case NPROCS: h->nprocs_file = R.i(); break;
      case PROCGRID: { int g[3]; R.iv(g, 3); break; }
      case NEWTON_PAIR: h->newton_pair = R.i(); break;
      case NEWTON_BOND: h->newton_bond = R.i(); break;
      case XPERIODIC: h->periodic[0] = R.i(); break;
      case YPERIODIC: h->periodic[1] = R.i(); break;
      case ZPERIODIC: h->periodic[2] = R.i(); break;
      case BOUNDARY: { int bnd[6]; R.iv(bnd, 6); break; }
      case BOUNDMIN: { double m[6]; R.dv(m, 6); break; }
      case ATOM_STYLE: {
        copy_str(h->atom_style, sizeof h->atom_style, R.s());
        const int nargs = R.i();
        for (int k = 0; k < nargs && !R.bad; k++) R.s();
        break;
      }
      case NATOMS: h->natoms = R.b(); break;
      case NTYPES: h->ntypes = R.i(); break;
      case NBONDS: h->nbonds = R.b(); break;
      case NBONDTYPES: h->nbondtypes = R.i(); break;
      case BOND_PER_ATOM: h->bond_per_atom = R.i(); break;
      case NANGLES: case NDIHEDRALS: case NIMPROPERS: case NELLIPSOIDS: case NLINES: case NTRIS: case NBODIES: R.b(); break;
      case NANGLETYPES: case ANGLE_PER_ATOM: case NDIHEDRALTYPES: case DIHEDRAL_PER_ATOM: case NIMPROPERTYPES: case IMPROPER_PER_ATOM:
      case ATOM_ID: case ATOM_MAP_STYLE: case ATOM_MAP_USER: case COMM_MODE: case COMM_VEL:
      case EXTRA_ANGLE_PER_ATOM: case EXTRA_DIHEDRAL_PER_ATOM: case EXTRA_IMPROPER_PER_ATOM: case EXTRA_SPECIAL_PER_ATOM: R.i(); break;
      case ATOM_SORTFREQ: h->atom_sortfreq = R.i(); break;
      case TRICLINIC: if (R.i() != 0) return seterr(err, errlen, "restart file holds a triclinic box"); break;
      case BOXLO: R.dv(h->boxlo, 3); break;
      case BOXHI: R.dv(h->boxhi, 3); break;
      case XY: case XZ: case YZ: case ATOM_SORTBIN: R.d(); break;
      case SPECIAL_LJ: R.dv(h->special_lj, 3); break;
      case SPECIAL_COUL: { double c[3]; R.dv(c, 3); break; }
      case TIMESTEP: h->dt = R.d(); break;
      case COMM_CUTOFF: h->comm_cutoff = R.d(); break;
      case EXTRA_BOND_PER_ATOM: h->extra_bond_per_atom = R.i(); break;
      case ATOM_MAXSPECIAL: h->maxspecial = R.i(); break;
      default: return seterr(err, errlen, "Invalid flag in header section of restart file");
    }
  }
  if (R.bad) return seterr(err, errlen, "restart file ends inside the header");
  if (strcmp(h->atom_style, "bond") != 0) return seterr(err, errlen, "restart file has atom_style %s: this path reads atom_style bond", h->atom_style);
  if (h->ntypes < 1 || h->ntypes > LE_RESTART_MAXT || h->nbondtypes < 0 || h->nbondtypes > LE_RESTART_MAXT)
    return seterr(err, errlen, "restart file has %d atom types / %d bond types (at most %d)", h->ntypes, h->nbondtypes, LE_RESTART_MAXT);
  // groups
  {
    const int ngroup = R.i();
    int count = 0;
    for (int k = 0; k < 32 && count < ngroup && !R.bad; k++) {
      const int n = R.i();
      if (n) { std::string t(n, '\0'); if (fread(&t[0], 1, n, R.fp) != (size_t)n) R.bad = true; count++; }
    }
  }
  // type arrays
  for (int flag = R.i(); flag >= 0 && !R.bad; flag = R.i()) {
    if (flag == MASS) R.dv(h->mass, LE_RESTART_MAXT);
    else return seterr(err, errlen, "Invalid flag in type arrays section of restart file");
  }
  // force fields
  const int nt = h->ntypes, nbt = h->nbondtypes;
  for (int flag = R.i(); flag >= 0 && !R.bad; flag = R.i()) {
    if (flag == PAIR) {
      copy_str(h->pair_style, sizeof h->pair_style, R.s());
      if (strcmp(h->pair_style, "lj/cut") != 0) return seterr(err, errlen, "restart file has pair_style %s: this path reads lj/cut", h->pair_style);
      h->cut_global = R.d(); h->offset_flag = R.i(); h->mix_flag = R.i(); h->tail_flag = R.i();      // write_restart_settings
      for (int a = 0; a < nt; a++)
        for (int b = a; b < nt; b++) {
          const int k = a * nt + b;
          h->pair_setflag[k] = R.i();
          if (h->pair_setflag[k]) { h->pair_eps[k] = R.d(); h->pair_sigma[k] = R.d(); h->pair_cut[k] = R.d(); }
        }
    } else if (flag == NO_PAIR) {
      copy_str(h->pair_style, sizeof h->pair_style, R.s());
    } else if (flag == BOND) {
      copy_str(h->bond_style, sizeof h->bond_style, R.s());
      if (strcmp(h->bond_style, "fene") == 0) {
        for (int t = 0; t < nbt; t++) h->bond_k[t] = R.d();
        for (int t = 0; t < nbt; t++) h->bond_r0[t] = R.d();
        for (int t = 0; t < nbt; t++) h->bond_eps[t] = R.d();
        for (int t = 0; t < nbt; t++) h->bond_sigma[t] = R.d();
        h->bond_coeffs_stored = 1;
      } else if (strcmp(h->bond_style, "harmonic") == 0) {
        for (int t = 0; t < nbt; t++) h->bond_k[t] = R.d();
        for (int t = 0; t < nbt; t++) h->bond_r0[t] = R.d();
        h->bond_coeffs_stored = 1;
      } else if (strcmp(h->bond_style, "hybrid") == 0) {
        // BondHybrid::write_restart stores the sub-style names only: the script must give bond_coeff again
        h->nhybrid = R.i();
        for (int m = 0; m < h->nhybrid && !R.bad; m++) { const std::string nm = R.s(); if (m < 4) copy_str(h->hybrid_styles[m], sizeof h->hybrid_styles[m], nm); }
        if (h->nhybrid > 4) return seterr(err, errlen, "bond_style hybrid with %d sub-styles", h->nhybrid);
      } else return seterr(err, errlen, "restart file has bond_style %s: this path reads fene, harmonic and hybrid of the two", h->bond_style);
    } else return seterr(err, errlen, "restart file holds angle / dihedral / improper styles: outside this path");
  }
  // fixes with restart info (Modify::read_restart): this path's fixes store none; skip what a reference run may have stored
  {
    const int nglobal = R.i();
    for (int k = 0; k < nglobal && !R.bad; k++) { R.s(); R.s(); const int nbytes = R.i(); if (nbytes < 0 || fseek(R.fp, nbytes, SEEK_CUR)) R.bad = true; }
    const int nper = R.i();
    if (nper > 0) return seterr(err, errlen, "restart file holds per-atom fix data (%d fixes): outside this path", nper);
  }
  // file layout
  for (int flag = R.i(); flag >= 0 && !R.bad; flag = R.i()) {
    if (flag == MULTIPROC) { if (R.i() != 0) return seterr(err, errlen, "multi-file restart (%%): read the base file's pieces with the reference"); }
    else if (flag == MPIIO) { if (R.i() != 0) return seterr(err, errlen, "MPI-IO restart files are not read"); }
    else return seterr(err, errlen, "Invalid flag in peratom section of restart file");
  }
  if (R.bad) return seterr(err, errlen, "restart file ends before the atom sections");
  return LE_OK;
}
}  // namespace

extern "C" int le_host_restart_read_header(const char *path, le_restart_header *h, char *err, int errlen) {
  if (!path || !h) return LE_EINVAL;
  FILE *fp = fopen(path, "rb");
  if (!fp) return seterr(err, errlen, "Cannot open restart file %s", path);
  Reader R{fp};
  const int rc = read_front(R, h, err, errlen);
  fclose(fp);
  return rc;
}

/* atoms in FILE order (the caller sorts by tag); bond_type / bond_atom are [natoms][bond_per_atom] */
extern "C" int le_host_restart_read_atoms(const char *path, int *tag, int *type, int *image, int *molecule, double *x, double *v,
                                          int *num_bond, int *bond_type, int *bond_atom, char *err, int errlen) {
  if (!path) return LE_EINVAL;
  FILE *fp = fopen(path, "rb");
  if (!fp) return seterr(err, errlen, "Cannot open restart file %s", path);
  Reader R{fp};
  le_restart_header h;
  int rc = read_front(R, &h, err, errlen);
  if (rc) { fclose(fp); return rc; }
  const int bpa = h.bond_per_atom;
  long long k = 0;
  std::vector<double> buf;
  for (;;) {
    const int flag = R.i();
    if (R.bad || flag != PERPROC) break;                      // the closing magic string follows the last section
    const int n = R.i();
    if (R.bad || n < 0) { rc = seterr(err, errlen, "Invalid flag in peratom section of restart file"); break; }
    buf.resize(n);
    if (n && fread(buf.data(), 8, n, fp) != (size_t)n) { rc = seterr(err, errlen, "restart file ends inside an atom section"); break; }
    int m = 0;
    while (m < n) {
      const int len = (int)buf[m];
      if (len < 12 || m + len > n || k >= h.natoms) { rc = seterr(err, errlen, "corrupt atom record in restart file"); break; }
      const double *a = &buf[m];
      if (x) { x[3 * k] = a[1]; x[3 * k + 1] = a[2]; x[3 * k + 2] = a[3]; }
      if (tag) tag[k] = (int)as_int(a[4]);
      if (type) type[k] = (int)as_int(a[5]);
      if (image) image[k] = (int)as_int(a[7]);
      if (v) { v[3 * k] = a[8]; v[3 * k + 1] = a[9]; v[3 * k + 2] = a[10]; }
      if (molecule) molecule[k] = (int)as_int(a[11]);
      const int nb = (int)as_int(a[12]);
      if (nb < 0 || nb > bpa || 13 + 2 * nb > len) { rc = seterr(err, errlen, "corrupt bond table in restart file"); break; }
      if (num_bond) num_bond[k] = nb;
      for (int q = 0; q < nb; q++) {
        if (bond_type) bond_type[k * bpa + q] = (int)as_int(a[13 + q]);
        if (bond_atom) bond_atom[k * bpa + q] = (int)as_int(a[13 + nb + q]);
      }
      m += len; k++;
    }
    if (rc) break;
  }
  fclose(fp);
  if (!rc && k != h.natoms) rc = seterr(err, errlen, "Did not assign all restart atoms correctly");
  return rc;
}

/* one file, one proc section; arrays in the order they shall be stored (the reference writes local order) */
extern "C" int le_host_restart_write(const char *path, const le_restart_header *h, const int *tag, const int *type, const int *image,
                                     const int *molecule, const double *x, const double *v, const int *num_bond, const int *bond_type,
                                     const int *bond_atom, char *err, int errlen) {
  if (!path || !h || !tag || !type || !x) return LE_EINVAL;
  FILE *fp = fopen(path, "wb");
  if (!fp) return seterr(err, errlen, "Cannot open restart file %s", path);
  Writer W{fp};
  fwrite(MAGIC, 1, 16, fp); W.i(ENDIAN_OK); W.i(FORMAT_REVISION);
  // header
  W.fs(VERSION, h->version[0] ? h->version : "29 Oct 2020");
  W.fi(SMALLINT, 4); W.fi(IMAGEINT, 4); W.fi(TAGINT, 4); W.fi(BIGINT, 8);
  W.fs(UNITS, h->units[0] ? h->units : "lj");
  W.fb(NTIMESTEP, h->ntimestep);
  W.fi(DIMENSION, 3); W.fi(NPROCS, 1);
  { const int g[3] = {1, 1, 1}; W.fiv(PROCGRID, 3, g); }
  W.fi(NEWTON_PAIR, h->newton_pair); W.fi(NEWTON_BOND, h->newton_bond);
  W.fi(XPERIODIC, h->periodic[0]); W.fi(YPERIODIC, h->periodic[1]); W.fi(ZPERIODIC, h->periodic[2]);
  { int bnd[6]; for (int q = 0; q < 6; q++) bnd[q] = h->periodic[q / 2] ? 0 : 1; W.fiv(BOUNDARY, 6, bnd); }
  { const double mn[6] = {0, 0, 0, 0, 0, 0}; W.fdv(BOUNDMIN, 6, mn); }
  W.fs(ATOM_STYLE, "bond"); W.i(0);
  W.fb(NATOMS, h->natoms); W.fi(NTYPES, h->ntypes);
  W.fb(NBONDS, h->nbonds); W.fi(NBONDTYPES, h->nbondtypes); W.fi(BOND_PER_ATOM, h->bond_per_atom);
  W.fb(NANGLES, 0); W.fi(NANGLETYPES, 0); W.fi(ANGLE_PER_ATOM, 0);
  W.fb(NDIHEDRALS, 0); W.fi(NDIHEDRALTYPES, 0); W.fi(DIHEDRAL_PER_ATOM, 0);
  W.fb(NIMPROPERS, 0); W.fi(NIMPROPERTYPES, 0); W.fi(IMPROPER_PER_ATOM, 0);
  W.fi(TRICLINIC, 0); W.fdv(BOXLO, 3, h->boxlo); W.fdv(BOXHI, 3, h->boxhi);
  W.fd(XY, 0.0); W.fd(XZ, 0.0); W.fd(YZ, 0.0);
  W.fdv(SPECIAL_LJ, 3, h->special_lj); W.fdv(SPECIAL_COUL, 3, h->special_lj);
  W.fd(TIMESTEP, h->dt);
  W.fi(ATOM_ID, 1); W.fi(ATOM_MAP_STYLE, 1); W.fi(ATOM_MAP_USER, 0); W.fi(ATOM_SORTFREQ, h->atom_sortfreq); W.fd(ATOM_SORTBIN, 0.0);
  W.fi(COMM_MODE, 0); W.fd(COMM_CUTOFF, h->comm_cutoff); W.fi(COMM_VEL, 0);
  W.fi(EXTRA_BOND_PER_ATOM, h->extra_bond_per_atom); W.fi(EXTRA_ANGLE_PER_ATOM, 0); W.fi(EXTRA_DIHEDRAL_PER_ATOM, 0);
  W.fi(EXTRA_IMPROPER_PER_ATOM, 0); W.fi(ATOM_MAXSPECIAL, h->maxspecial);
  W.fb(NELLIPSOIDS, 0); W.fb(NLINES, 0); W.fb(NTRIS, 0); W.fb(NBODIES, 0);
  W.i(-1);
  // groups: "all"
  W.i(1); W.i(4); fwrite("all", 1, 4, fp);
  // type arrays
  W.fdv(MASS, h->ntypes, h->mass); W.i(-1);
  // force fields
  const int nt = h->ntypes, nbt = h->nbondtypes;
  if (h->pair_style[0]) {
    W.fs(PAIR, "lj/cut");
    fwrite(&h->cut_global, 8, 1, fp); W.i(h->offset_flag); W.i(h->mix_flag); W.i(h->tail_flag);
    for (int a = 0; a < nt; a++)
      for (int b = a; b < nt; b++) {
        const int k = a * nt + b;
        W.i(h->pair_setflag[k]);
        if (h->pair_setflag[k]) { fwrite(&h->pair_eps[k], 8, 1, fp); fwrite(&h->pair_sigma[k], 8, 1, fp); fwrite(&h->pair_cut[k], 8, 1, fp); }
      }
  }
  if (h->bond_style[0]) {
    W.fs(BOND, h->bond_style);
    if (strcmp(h->bond_style, "fene") == 0) { fwrite(h->bond_k, 8, nbt, fp); fwrite(h->bond_r0, 8, nbt, fp); fwrite(h->bond_eps, 8, nbt, fp); fwrite(h->bond_sigma, 8, nbt, fp); }
    else if (strcmp(h->bond_style, "harmonic") == 0) { fwrite(h->bond_k, 8, nbt, fp); fwrite(h->bond_r0, 8, nbt, fp); }
    else if (strcmp(h->bond_style, "hybrid") == 0) {
      W.i(h->nhybrid);
      for (int m = 0; m < h->nhybrid; m++) { const int n = (int)strlen(h->hybrid_styles[m]) + 1; W.i(n); fwrite(h->hybrid_styles[m], 1, n, fp); }
    } else { fclose(fp); return seterr(err, errlen, "write_restart: bond_style %s", h->bond_style); }
  }
  W.i(-1);
  // fixes with restart info: none
  W.i(0); W.i(0);
  // file layout
  W.fi(MULTIPROC, 0); W.fi(MPIIO, 0); W.i(-1);
  // the atoms: one section
  const int bpa = h->bond_per_atom;
  std::vector<double> buf;
  buf.reserve((size_t)h->natoms * 18);
  for (long long k = 0; k < h->natoms; k++) {
    const int nb = num_bond ? num_bond[k] : 0;
    const size_t m0 = buf.size();
    buf.push_back(0.0);
    buf.push_back(x[3 * k]); buf.push_back(x[3 * k + 1]); buf.push_back(x[3 * k + 2]);
    buf.push_back(as_double(tag[k])); buf.push_back(as_double(type[k])); buf.push_back(as_double(1)); buf.push_back(as_double(image ? image[k] : ((512 << 20) | (512 << 10) | 512)));
    buf.push_back(v ? v[3 * k] : 0.0); buf.push_back(v ? v[3 * k + 1] : 0.0); buf.push_back(v ? v[3 * k + 2] : 0.0);
    buf.push_back(as_double(molecule ? molecule[k] : 0));
    buf.push_back(as_double(nb));
    for (int q = 0; q < nb; q++) buf.push_back(as_double(bond_type[k * bpa + q]));
    for (int q = 0; q < nb; q++) buf.push_back(as_double(bond_atom[k * bpa + q]));
    buf[m0] = (double)(buf.size() - m0);
  }
  if (buf.size() > 0x7fffffffULL) { fclose(fp); return seterr(err, errlen, "write_restart: more than 2^31 values in one proc section"); }
  W.fdv(PERPROC, (int)buf.size(), buf.data());
  fwrite(MAGIC, 1, 16, fp);
  const bool bad = ferror(fp) != 0;
  fclose(fp);
  return bad ? seterr(err, errlen, "I/O error while writing restart") : LE_OK;
}
