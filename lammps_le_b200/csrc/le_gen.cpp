// le_gen.cpp -- synthetic input generators (host only; used by bench.py and the tests).
//
// The reference ships no chromatin input at all, and its melt input
// bench/data.chain came out of tools/chain.f.  These generators produce the systems SURVEY.md
// section 8(d) defines: a self-avoiding bead-spring walk wrapped into a periodic cube, and a
// lattice-started FENE melt.  Nothing here is on the hot path.
#include "../../include/le_b200.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>

namespace {
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) {}
  uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

inline int pack_image(int ix, int iy, int iz) {
  return ((ix + 512) & 1023) | (((iy + 512) & 1023) << 10) | (((iz + 512) & 1023) << 20);
}
}  // namespace

// Self-avoiding random walk of n beads, bond length `step`, no two beads closer than rmin (checked against
// every bead already placed through a periodic hash grid), wrapped into the cube [0,L)^3.
// x[n*3] receives wrapped coordinates, image[n] the LAMMPS-packed image flags.  nchains > 1 starts a new
// walk at a random point every n/nchains beads.
extern "C" int le_gen_saw_chains(int n, int nchains, double L, double step, double rmin, uint64_t seed,
                                 double *x, int *image) {
  if (n < 1 || nchains < 1 || n % nchains || !(L > 2.0 * step) || !x) return LE_EINVAL;
  Rng rng(seed);
  const int ng = (int)floor(L / (rmin > 0.5 ? rmin : 0.5));
  const int g = ng < 1 ? 1 : (ng > 1024 ? 1024 : ng);
  const double inv = g / L;
  std::vector<int> head((size_t)g * g * g, -1), next(n, -1);
  std::vector<double> ux((size_t)n * 3);  // unwrapped
  const double rmin2 = rmin * rmin;
  const int len = n / nchains;
  auto wrap = [&](double v) { v = fmod(v, L); if (v < 0) v += L; if (v >= L) v = 0.0; return v; };
  auto cell = [&](double v) { int c = (int)(v * inv); return c >= g ? g - 1 : c; };
  auto min_d2 = [&](const double w[3], int skip_from) {
    // smallest squared distance to any placed bead with index < skip_from's protected neighbours
    double best = 1e300;
    const int cx = cell(w[0]), cy = cell(w[1]), cz = cell(w[2]);
    const int r = (g >= 3) ? 1 : 0;
    for (int dz = -r; dz <= r; dz++)
      for (int dy = -r; dy <= r; dy++)
        for (int dx = -r; dx <= r; dx++) {
          const int ax = (cx + dx + g) % g, ay = (cy + dy + g) % g, az = (cz + dz + g) % g;
          for (int j = head[((size_t)az * g + ay) * g + ax]; j >= 0; j = next[j]) {
            if (j >= skip_from) continue;
            double d2 = 0;
            for (int q = 0; q < 3; q++) {
              double d = w[q] - wrap(ux[3 * (size_t)j + q]);
              d -= L * rint(d / L);
              d2 += d * d;
            }
            if (d2 < best) best = d2;
          }
        }
    return best;
  };
  for (int k = 0; k < n; k++) {
    double cand[3], bestc[3] = {0, 0, 0}, bestd = -1;
    const bool start = (k % len) == 0;
    for (int attempt = 0; attempt < 200; attempt++) {
      if (start) {
        for (int q = 0; q < 3; q++) cand[q] = rng.uni() * L;
      } else {
        const double z = 2.0 * rng.uni() - 1.0, ph = 6.283185307179586 * rng.uni(), s = sqrt(1.0 - z * z);
        cand[0] = ux[3 * (size_t)(k - 1)] + step * s * cos(ph);
        cand[1] = ux[3 * (size_t)(k - 1) + 1] + step * s * sin(ph);
        cand[2] = ux[3 * (size_t)(k - 1) + 2] + step * z;
      }
      const double w[3] = {wrap(cand[0]), wrap(cand[1]), wrap(cand[2])};
      // the bonded predecessor sits at distance `step` by construction: exclude only it
      const double d2 = min_d2(w, start ? k : k - 1);
      if (d2 > bestd) { bestd = d2; for (int q = 0; q < 3; q++) bestc[q] = cand[q]; }
      if (d2 >= rmin2) break;
    }
    for (int q = 0; q < 3; q++) ux[3 * (size_t)k + q] = bestc[q];
    const double w[3] = {wrap(bestc[0]), wrap(bestc[1]), wrap(bestc[2])};
    const size_t c = ((size_t)cell(w[2]) * g + cell(w[1])) * g + cell(w[0]);
    next[k] = head[c];
    head[c] = k;
    int im[3];
    for (int q = 0; q < 3; q++) {
      x[3 * (size_t)k + q] = w[q];
      im[q] = (int)floor(bestc[q] / L);
    }
    if (image) image[k] = pack_image(im[0], im[1], im[2]);
  }
  return LE_OK;
}

// FENE melt start: a boustrophedon path through a simple-cubic lattice of m^3 >= nchains*len sites with
// spacing a = rho^(-1/3), cut into chains of `len` beads.  Returns the box length in *L.
extern "C" int le_gen_lattice_melt(int nchains, int len, double rho, double *L, double *x, int *image) {
  if (nchains < 1 || len < 1 || !(rho > 0) || !L || !x) return LE_EINVAL;
  const long long n = (long long)nchains * len;
  int m = (int)ceil(cbrt((double)n) - 1e-9);
  while ((long long)m * m * m < n) m++;
  const double boxl = cbrt((double)n / rho);
  const double a = boxl / m;
  *L = boxl;
  long long k = 0;
  for (int iz = 0; iz < m && k < n; iz++)
    for (int jy = 0; jy < m && k < n; jy++) {
      const int iy = (iz & 1) ? m - 1 - jy : jy;
      for (int jx = 0; jx < m && k < n; jx++) {
        const int ix = ((jy + iz * m) & 1) ? m - 1 - jx : jx;   // consecutive rows run in opposite directions
        x[3 * k] = (ix + 0.25) * a; x[3 * k + 1] = (iy + 0.25) * a; x[3 * k + 2] = (iz + 0.25) * a;
        if (image) image[k] = pack_image(0, 0, 0);
        k++;
      }
    }
  return LE_OK;
}

// ------------------------------------------------------------------------------------------------
// `velocity all create T seed [dist uniform|gaussian] [mom yes|no] [loop all|local|geom]`
//   reference: Velocity::create src/velocity.cpp:162-401, RanPark src/random_park.cpp:25-128,
//   Velocity::zero_momentum src/velocity.cpp:741-767, ComputeTemp::compute_scalar src/compute_temp.cpp:83-110
//   (dof = 3N - 3), Velocity::rescale src/velocity.cpp:716-735.  Host only: the deck front end and the Python
//   mirror call it and upload the result.
// ------------------------------------------------------------------------------------------------
namespace {
struct RanPark {                      // Park-Miller minimal standard generator, src/random_park.cpp:25-74
  int seed, save;
  double second;
  explicit RanPark(int s) : seed(s), save(0), second(0.0) {}
  double uniform() {
    const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
    const int k = seed / IQ;
    seed = IA * (seed - k * IQ) - IR * k;
    if (seed < 0) seed += IM;
    return (1.0 / IM) * seed;
  }
  double gaussian() {
    double first;
    if (!save) {
      double v1, v2, rsq;
      do {
        v1 = 2.0 * uniform() - 1.0;
        v2 = 2.0 * uniform() - 1.0;
        rsq = v1 * v1 + v2 * v2;
      } while (rsq >= 1.0 || rsq == 0.0);
      const double fac = sqrt(-2.0 * log(rsq) / rsq);
      second = v1 * fac;
      first = v2 * fac;
      save = 1;
    } else {
      first = second;
      save = 0;
    }
    return first;
  }
  void reset(int ibase, const double *coord) {      // src/random_park.cpp:92-128: hash of the seed and the coordinates
    unsigned int hash = 0;
    const char *str = (const char *)&ibase;
    for (size_t i = 0; i < sizeof(int); i++) { hash += str[i]; hash += (hash << 10); hash ^= (hash >> 6); }
    str = (const char *)coord;
    for (size_t i = 0; i < 3 * sizeof(double); i++) { hash += str[i]; hash += (hash << 10); hash ^= (hash >> 6); }
    hash += (hash << 3); hash ^= (hash >> 11); hash += (hash << 15);
    seed = hash & 0x7ffffff;
    if (!seed) seed = 1;
    for (int i = 0; i < 5; i++) uniform();
    save = 0;
  }
};
}  // namespace

extern "C" int le_host_velocity_create(int n, const int *type, const double *mass_per_type, const double *x, double t_desired, int seed,
                                       int dist, int mom, int loop, double *v) {
  if (n < 1 || !type || !mass_per_type || !v || seed <= 0 || t_desired < 0.0 || dist < 0 || dist > 1 || loop < 0 || loop > 2) return LE_EINVAL;
  if (loop == 2 && !x) return LE_EINVAL;
  RanPark random(loop == 2 ? 1 : seed);
  if (loop == 1) for (int i = 0; i < 100; i++) random.uniform();     // WARMUP of loop local (seed + me, me = 0)
  for (int i = 0; i < n; i++) {
    if (loop == 2) random.reset(seed, x + 3 * (size_t)i);
    double vx, vy, vz;
    if (dist == 0) { vx = random.uniform() - 0.5; vy = random.uniform() - 0.5; vz = random.uniform() - 0.5; }
    else { vx = random.gaussian(); vy = random.gaussian(); vz = random.gaussian(); }
    const double factor = 1.0 / sqrt(mass_per_type[type[i] - 1]);
    v[3 * (size_t)i] = vx * factor; v[3 * (size_t)i + 1] = vy * factor; v[3 * (size_t)i + 2] = vz * factor;
  }
  if (mom) {   // Group::vcm: p[] summed in atom order, divided by the total mass; then subtracted from every atom
    double p[3] = {0.0, 0.0, 0.0}, masstotal = 0.0;
    for (int i = 0; i < n; i++) masstotal += mass_per_type[type[i] - 1];
    for (int i = 0; i < n; i++) {
      const double m = mass_per_type[type[i] - 1];
      p[0] += v[3 * (size_t)i] * m; p[1] += v[3 * (size_t)i + 1] * m; p[2] += v[3 * (size_t)i + 2] * m;
    }
    double vcm[3] = {0.0, 0.0, 0.0};
    if (masstotal > 0.0) for (int k = 0; k < 3; k++) vcm[k] = p[k] / masstotal;
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) v[3 * (size_t)i + k] -= vcm[k];
  }
  // compute temp: t = sum m v^2 / dof, dof = 3N - 3 (extra_dof = dimension)
  double t = 0.0;
  for (int i = 0; i < n; i++)
    t += (v[3 * (size_t)i] * v[3 * (size_t)i] + v[3 * (size_t)i + 1] * v[3 * (size_t)i + 1] + v[3 * (size_t)i + 2] * v[3 * (size_t)i + 2]) *
         mass_per_type[type[i] - 1];
  const double dof = 3.0 * n - 3.0;
  if (dof <= 0.0) return LE_EINVAL;
  t *= 1.0 / dof;
  if (t == 0.0) return LE_EINVAL;                                  // "Attempting to rescale a 0.0 temperature"
  const double factor = sqrt(t_desired / t);
  for (size_t k = 0; k < 3 * (size_t)n; k++) v[k] *= factor;
  return LE_OK;
}

// compute property/local batom1 batom2 btype (src/compute_property_local.cpp:463-493, pack_batom1/2, pack_btype)
extern "C" int64_t le_host_property_local_bonds(int n, int bpa, const int *num_bond, const int *bond_type, const int *bond_atom,
                                                int newton_bond, int *rows) {
  if (n < 0 || bpa < 1 || !num_bond || !bond_type || !bond_atom) return -1;
  int64_t m = 0;
  for (int a1 = 0; a1 < n; a1++)
    for (int i = 0; i < num_bond[a1]; i++) {
      const int t2 = bond_atom[(size_t)a1 * bpa + i];
      if (t2 < 1 || t2 > n) continue;                       // partner unknown to this proc
      if (newton_bond == 0 && a1 + 1 > t2) continue;
      if (bond_type[(size_t)a1 * bpa + i] == 0) continue;
      if (rows) { rows[3 * m] = a1 + 1; rows[3 * m + 1] = t2; rows[3 * m + 2] = bond_type[(size_t)a1 * bpa + i]; }
      m++;
    }
  return m;
}
