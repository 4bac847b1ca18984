// ref_shim.cc -- extern "C" doorway into the UNMODIFIED reference objects.
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Compiled by oracle/Makefile together with the
// reference's own LJPot.cc / LJClusterPot.cc / PotHelpers.cc and the vendored
// vesin-single-build.cpp, read in place from /root/reference, into oracle/_ref/libref.so.
// No reference source is copied into this repository; this file only calls their API.
#include <array>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "rgpot/LennardJones/LJClusterPot.hpp"
#include "rgpot/LennardJones/LJPot.hpp"
#include "rgpot/Morse/MorsePot.hpp"
#include "rgpot/ZBL/ZBLPot.hpp"
#include "vesin.h"

#include "oracle.h"

extern "C" {

// rgpot::LJPot / LJClusterPot forceImpl through the reference classes themselves.
int ref_lj_force(long n, const double *pos, const int *Z, const double *box, double u0,
                 double cutoff, double psi, int cluster, double *F, double *energy) {
  rgpot::ForceInput in{static_cast<size_t>(n), pos, Z, box};
  rgpot::ForceOut out{F, 0.0, 0.0};
  try {
    if (cluster) {
      rgpot::LJClusterPot pot{rgpot::LJClusterConfig{u0, cutoff, psi}};
      pot.forceImpl(in, &out);
    } else {
      rgpot::LJPot pot{rgpot::LJConfig{u0, cutoff, psi}};
      pot.forceImpl(in, &out);
    }
  } catch (...) {
    return 1;
  }
  *energy = out.energy;
  return 0;
}

// rgpot::MorsePot::forceImpl through the reference class itself.
int ref_morse_force(long n, const double *pos, const int *Z, const double *box, double De, double a,
                    double re, double cutoff, double *F, double *energy) {
  rgpot::ForceInput in{static_cast<size_t>(n), pos, Z, box};
  rgpot::ForceOut out{F, 0.0, 0.0};
  try {
    rgpot::MorsePot pot{rgpot::MorseConfig{De, a, re, cutoff}};
    pot.forceImpl(in, &out);
  } catch (...) {
    return 1;
  }
  *energy = out.energy;
  return 0;
}

// rgpot::ZBLPot::forceImpl through the reference class itself.
int ref_zbl_force(long n, const double *pos, const int *Z, const double *box, double cut_inner,
                  double cut_global, double *F, double *energy) {
  rgpot::ForceInput in{static_cast<size_t>(n), pos, Z, box};
  rgpot::ForceOut out{F, 0.0, 0.0};
  try {
    rgpot::ZBLPot pot{rgpot::ZBLConfig{cut_inner, cut_global}};
    pot.forceImpl(in, &out);
  } catch (...) {
    return 1;
  }
  *energy = out.energy;
  return 0;
}

uint64_t ref_lj_params_key(double u0, double cutoff, double psi, int cluster) {
  if (cluster) return rgpot::LJClusterPot{rgpot::LJClusterConfig{u0, cutoff, psi}}.paramsKey();
  return rgpot::LJPot{rgpot::LJConfig{u0, cutoff, psi}}.paramsKey();
}

// The vesin call rgpot_neighbors.f90:102-110 makes, turned into the same CSR table
// (rgpot_neighbors.f90:126-160).  `handle` keeps the VesinNeighborList alive across calls
// like the module-saved table does (so Verlet reuse can be exercised); pass NULL for a
// one-shot build.
struct RefTable {
  VesinNeighborList list;
};

void *ref_table_new() { return new RefTable(); }
void ref_table_delete(void *h) {
  if (!h) return;
  auto *t = static_cast<RefTable *>(h);
  vesin_free(&t->list);
  delete t;
}

int ref_vesin_table(void *handle, long n, const double *pos, const double *box, double cutoff,
                    double skin, int n_threads, oracle_nlist_t *out, char *err, size_t errlen) {
  std::memset(out, 0, sizeof(*out));
  RefTable local;
  RefTable *t = handle ? static_cast<RefTable *>(handle) : &local;
  VesinOptions opt{};
  opt.cutoff = cutoff;
  opt.full = true;
  opt.sorted = true;
  opt.algorithm = VesinAutoAlgorithm;
  opt.skin = skin;
  opt.n_threads = n_threads;
  opt.return_shifts = true;
  opt.return_distances = true;
  opt.return_vectors = true;
  const bool periodic[3] = {true, true, true};
  const char *msg = nullptr;
  double cell[3][3];
  std::memcpy(cell, box, sizeof cell);
  int st = vesin_neighbors(reinterpret_cast<const double(*)[3]>(pos), static_cast<size_t>(n), cell,
                           periodic, VesinDevice{VesinCPU, 0}, opt, &t->list, &msg);
  if (st != 0) {
    std::snprintf(err, errlen, "%s", msg ? msg : "vesin failed");
    if (!handle) vesin_free(&local.list);
    return st;
  }
  const auto &L = t->list;
  out->row = static_cast<int32_t *>(std::calloc(static_cast<size_t>(n) + 1, sizeof(int32_t)));
  long total = 0, selfimg = 0;
  for (size_t p = 0; p < L.length; ++p) {
    if (L.pairs[p][0] == L.pairs[p][1]) {
      ++selfimg;
      continue;
    }
    out->row[L.pairs[p][0] + 1]++;
    ++total;
  }
  for (long i = 0; i < n; ++i) out->row[i + 1] += out->row[i];
  const size_t cap = static_cast<size_t>(total > 0 ? total : 1);
  out->n_pairs = total;
  out->n_self_images = selfimg;
  out->idx = static_cast<int32_t *>(std::malloc(sizeof(int32_t) * cap));
  out->shift = static_cast<int32_t *>(std::malloc(sizeof(int32_t) * 3 * cap));
  out->vec = static_cast<double *>(std::malloc(sizeof(double) * 3 * cap));
  out->dist = static_cast<double *>(std::malloc(sizeof(double) * cap));
  std::vector<int32_t> fill(static_cast<size_t>(n > 0 ? n : 1), 0);
  for (size_t p = 0; p < L.length; ++p) {
    const size_t i = L.pairs[p][0], j = L.pairs[p][1];
    if (i == j) continue;
    const long slot = out->row[i] + fill[i]++;
    out->idx[slot] = static_cast<int32_t>(j);
    for (int k = 0; k < 3; ++k) {
      out->shift[3 * slot + k] = L.shifts[p][k];
      out->vec[3 * slot + k] = L.vectors[p][k];
    }
    out->dist[slot] = L.distances[p];
  }
  if (!handle) vesin_free(&local.list);
  return 0;
}

// The vendored vesin itself, raw: flat copies of VesinNeighborList for the neighbor-list parity
// tests (full / half lists, self images kept, periodic or free).  Caller frees with ref_vesin_raw_free.
typedef struct {
  long length;
  long long *pairs;  // 2 per entry
  int32_t *shifts;   // 3 per entry
  double *dists;
  double *vecs;      // 3 per entry
} ref_raw_list_t;

void ref_vesin_raw_free(ref_raw_list_t *r) {
  if (!r) return;
  std::free(r->pairs);
  std::free(r->shifts);
  std::free(r->dists);
  std::free(r->vecs);
  std::memset(r, 0, sizeof *r);
}

int ref_vesin_raw(long n, const double *pos, const double *box, int periodic_all, double cutoff, int full,
                  ref_raw_list_t *out, char *err, size_t errlen) {
  std::memset(out, 0, sizeof *out);
  VesinNeighborList list;
  VesinOptions opt{};
  opt.cutoff = cutoff;
  opt.full = full != 0;
  opt.sorted = true;
  opt.algorithm = VesinAutoAlgorithm;
  opt.skin = 0.0;
  opt.n_threads = 1;
  opt.return_shifts = true;
  opt.return_distances = true;
  opt.return_vectors = true;
  const bool periodic[3] = {periodic_all != 0, periodic_all != 0, periodic_all != 0};
  const char *msg = nullptr;
  double cell[3][3];
  std::memcpy(cell, box, sizeof cell);
  int st = vesin_neighbors(reinterpret_cast<const double(*)[3]>(pos), static_cast<size_t>(n), cell, periodic,
                           VesinDevice{VesinCPU, 0}, opt, &list, &msg);
  if (st != 0) {
    std::snprintf(err, errlen, "%s", msg ? msg : "vesin failed");
    vesin_free(&list);
    return st;
  }
  const size_t m = list.length, cap = m ? m : 1;
  out->length = static_cast<long>(m);
  out->pairs = static_cast<long long *>(std::malloc(sizeof(long long) * 2 * cap));
  out->shifts = static_cast<int32_t *>(std::malloc(sizeof(int32_t) * 3 * cap));
  out->dists = static_cast<double *>(std::malloc(sizeof(double) * cap));
  out->vecs = static_cast<double *>(std::malloc(sizeof(double) * 3 * cap));
  for (size_t p = 0; p < m; ++p) {
    out->pairs[2 * p] = static_cast<long long>(list.pairs[p][0]);
    out->pairs[2 * p + 1] = static_cast<long long>(list.pairs[p][1]);
    for (int k = 0; k < 3; ++k) {
      out->shifts[3 * p + k] = list.shifts[p][k];
      out->vecs[3 * p + k] = list.vectors[p][k];
    }
    out->dists[p] = list.distances[p];
  }
  vesin_free(&list);
  return 0;
}

// Linear-scaling LJ comparator ("reference components"): the vendored vesin cell list
// (half list, all host threads for the filter) feeding the pair lambda of LJPot.cc:61-81.  NOT the
// reference algorithm -- LJPot scans all N^2/2 pairs -- but what the reference's own parts give when
// put together; same pair set and pair terms for boxes wider than two cutoffs.
int ref_lj_vesin_force(long n, const double *pos, const double *box, double u0, double cutoff, double psi,
                       int n_threads, double *F, double *energy, long *n_pairs, char *err, size_t errlen) {
  *energy = 0.0;
  if (n_pairs) *n_pairs = 0;
  for (long i = 0; i < 3 * n; ++i) F[i] = 0.0;
  if (n < 2) return 0;
  VesinNeighborList list;
  VesinOptions opt{};
  opt.cutoff = cutoff;
  opt.full = false;
  opt.sorted = false;
  opt.algorithm = VesinAutoAlgorithm;
  opt.skin = 0.0;
  opt.n_threads = n_threads;
  opt.return_shifts = false;
  opt.return_distances = true;
  opt.return_vectors = true;
  const bool periodic[3] = {true, true, true};
  const char *msg = nullptr;
  double cell[3][3] = {{box[0], 0, 0}, {0, box[4], 0}, {0, 0, box[8]}};  // LJPot.cc:52-56: the diagonal only
  int st = vesin_neighbors(reinterpret_cast<const double(*)[3]>(pos), static_cast<size_t>(n), cell, periodic,
                           VesinDevice{VesinCPU, 0}, opt, &list, &msg);
  if (st != 0) {
    std::snprintf(err, errlen, "%s", msg ? msg : "vesin failed");
    vesin_free(&list);
    return st;
  }
  const double a_c = std::pow(psi / cutoff, 6.0), cut_u = 4.0 * u0 * a_c * (a_c - 1.0);  // LJPot.hpp:45-60
  const double psi2 = psi * psi;
  double e = 0.0;
  for (size_t p = 0; p < list.length; ++p) {
    const size_t i = list.pairs[p][0], j = list.pairs[p][1];
    const double dx = -list.vectors[p][0], dy = -list.vectors[p][1], dz = -list.vectors[p][2];  // r_i - r_j
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double inv = 1.0 / r2, sr2 = psi2 * inv, a = sr2 * sr2 * sr2, b = 4.0 * u0 * a;
    e += b * (a - 1.0) - cut_u;
    const double fs = 6.0 * b * inv * (2.0 * a - 1.0);
    F[3 * i] += fs * dx;
    F[3 * i + 1] += fs * dy;
    F[3 * i + 2] += fs * dz;
    F[3 * j] -= fs * dx;
    F[3 * j + 1] -= fs * dy;
    F[3 * j + 2] -= fs * dz;
  }
  *energy = e;
  if (n_pairs) *n_pairs = static_cast<long>(list.length);
  vesin_free(&list);
  return 0;
}

// CuH2 as the reference evaluates it, as far as this image allows: the reference's own
// vesin (unmodified) feeding the restated Fortran passes (oracle/cuh2_oracle.c).
int ref_cuh2_force(void *handle, int32_t natoms, const double *pos, const int32_t *Z,
                   const double *cell, double *F, double *energy, char *err, size_t errlen) {
  if (err && errlen) err[0] = 0;
  *energy = 0.0;
  for (long i = 0; i < 3L * (natoms > 0 ? natoms : 0); ++i) F[i] = 0.0;
  if (natoms < 1) return 0;
  for (int32_t i = 0; i < natoms; ++i)
    if (Z[i] != 29 && Z[i] != 1) {
      oracle_nlist_t empty{};
      return oracle_cuh2_force_table(natoms, Z, &empty, F, energy, err, errlen);
    }
  oracle_nlist_t t;
  char verr[256] = {0};
  // n_threads = 0 -> all cores for the filter; the Verlet rebuild stays single-threaded
  // inside vesin exactly as in the reference (vesin-single-build.cpp:1921-1929).
  int st = ref_vesin_table(handle, natoms, pos, cell, 6.1, 0.1 * 6.1, 0, &t, verr, sizeof verr);
  if (st != 0) {
    std::snprintf(err, errlen, "rgpot_neighbors: vesin build failed: %s", verr);
    return st;
  }
  st = oracle_cuh2_force_table(natoms, Z, &t, F, energy, err, errlen);
  oracle_nlist_free(&t);
  return st;
}

} // extern "C"
