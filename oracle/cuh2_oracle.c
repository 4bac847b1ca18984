/*
 * cuh2_oracle.c -- CPU restatement of rgpot's CuH2 copper-hydrogen EAM (a Fortran 2018
 * kernel upstream; no Fortran compiler exists in this image, so it is restated in C).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity: PINNED against
 * CppCore/tests/CuH2PotTest.cc:11-31 (17-digit energy and 12 force components),
 * tests/rpc_integ.py:52-78 (dimer) and the invariance suite of
 * CppCore/rgpot/fortran/test_cuh2.f90:71-137.
 *
 * All line numbers: CppCore/rgpot/fortran/rgpot_cuh2.f90 unless noted.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* cuh2_params_t defaults, :58-105 */
typedef struct {
  double da[3], aa[3], db[3], ab[3]; /* 0 CuCu, 1 HCu, 2 HH */
  double scale[2], ba[2], bb[2], gamma[2];
  int neta[2]; /* 0 Cu, 1 H */
  double f_cu[8], f_h[5];
  double rcut;
  double phicut[3], rhocut[2];
} params_t;

static const params_t kDefaults = {
    {2862.3, 86.1495, 79.5013},
    {3.51236, 4.21154, 2.47961},
    {-109.107, 15355.5, -107.555},
    {1.75618, 6.07604, 2.99918},
    {0.273072, 2.14389},
    {3.69051, 3.77701},
    {7.38101, 0.0},
    {512.0, 0.0},
    {6, 0},
    {-112.945, 8510.04, -261734.0, 4780090.0, -52341900.0, 339124000.0, -1201500000.0,
     1796190000.0},
    {-81.7537, 838.668, -3952.35, 8767.87, -6599.37},
    6.1,
    {0.0, 0.0, 0.0},
    {0.0, 0.0}};

/* real ** integer: libgfortran's pow_r8_i4 square-and-multiply. */
static double powi(double x, int n) {
  double p = 1.0;
  if (n == 0) return p;
  unsigned u = (unsigned)(n < 0 ? -n : n);
  for (;;) {
    if (u & 1u) p *= x;
    u >>= 1;
    if (u)
      x *= x;
    else
      break;
  }
  return n < 0 ? 1.0 / p : p;
}

/* :143-218 */
static double phi(const params_t *p, int k, double r) {
  const double t1 = p->da[k] * exp(-p->aa[k] * r);
  const double t2 = p->db[k] * exp(-p->ab[k] * r);
  return t1 + t2 - p->phicut[k];
}
static double dphi(const params_t *p, int k, double r) {
  const double t1 = p->da[k] * exp(-p->aa[k] * r);
  const double t2 = p->db[k] * exp(-p->ab[k] * r);
  return -(p->aa[k] * t1 + p->ab[k] * t2) / r;
}
/* :228-291 */
static double rho(const params_t *p, int k, double r) {
  const double t1 = exp(-p->ba[k] * r);
  const double t2 = p->gamma[k] * exp(-p->bb[k] * r);
  const double t3 = p->scale[k] * powi(r, p->neta[k]);
  return t3 * (t1 + t2) - p->rhocut[k];
}
static double drho(const params_t *p, int k, double r) {
  const double rinv = 1.0 / r;
  const double t1 = exp(-p->ba[k] * r);
  const double t2 = p->gamma[k] * exp(-p->bb[k] * r);
  const double t3 = p->scale[k] * powi(r, p->neta[k]);
  const double raw = t3 * (t1 + t2);
  return ((double)p->neta[k] * raw * rinv - t3 * (p->ba[k] * t1 + p->bb[k] * t2)) * rinv;
}
/* :298-370 -- explicit powers, left-to-right sums */
static double f_cu(const params_t *p, double x) {
  const double x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x, x6 = x5 * x, x7 = x6 * x,
               x8 = x7 * x;
  const double *f = p->f_cu;
  return f[0] * x + f[1] * x2 + f[2] * x3 + f[3] * x4 + f[4] * x5 + f[5] * x6 + f[6] * x7 +
         f[7] * x8;
}
static double df_cu(const params_t *p, double x) {
  const double x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x, x6 = x5 * x, x7 = x6 * x;
  const double *f = p->f_cu;
  return f[0] + 2.0 * f[1] * x + 3.0 * f[2] * x2 + 4.0 * f[3] * x3 + 5.0 * f[4] * x4 +
         6.0 * f[5] * x5 + 7.0 * f[6] * x6 + 8.0 * f[7] * x7;
}
static double f_h(const params_t *p, double x) {
  const double x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x;
  const double *f = p->f_h;
  return f[0] * x + f[1] * x2 + f[2] * x3 + f[3] * x4 + f[4] * x5;
}
static double df_h(const params_t *p, double x) {
  const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
  const double *f = p->f_h;
  return f[0] + 2.0 * f[1] * x + 3.0 * f[2] * x2 + 4.0 * f[3] * x3 + 5.0 * f[4] * x4;
}

/* cuh2_cut_shifted :118-135 */
static params_t cut_shifted(const params_t *self) {
  params_t p = *self;
  for (int k = 0; k < 3; ++k) p.phicut[k] = 0.0;
  for (int k = 0; k < 2; ++k) p.rhocut[k] = 0.0;
  p.phicut[0] = phi(&p, 0, p.rcut);
  p.phicut[1] = phi(&p, 1, p.rcut);
  p.phicut[2] = phi(&p, 2, p.rcut);
  p.rhocut[0] = rho(&p, 0, p.rcut);
  p.rhocut[1] = rho(&p, 1, p.rcut);
  return p;
}

static const params_t *shifted_defaults(void) {
  static params_t p;
  static int ready = 0;
  if (!ready) {
    p = cut_shifted(&kDefaults);
    ready = 1;
  }
  return &p;
}

double oracle_cuh2_phi(int kind, double r) { return phi(shifted_defaults(), kind, r); }
double oracle_cuh2_dphi(int kind, double r) { return dphi(shifted_defaults(), kind, r); }
double oracle_cuh2_rho(int kind, double r) { return rho(shifted_defaults(), kind, r); }
double oracle_cuh2_drho(int kind, double r) { return drho(shifted_defaults(), kind, r); }
double oracle_cuh2_embed(int kind, double x) {
  return kind == 0 ? f_cu(shifted_defaults(), x) : f_h(shifted_defaults(), x);
}
double oracle_cuh2_dembed(int kind, double x) {
  return kind == 0 ? df_cu(shifted_defaults(), x) : df_h(shifted_defaults(), x);
}

enum { SP_CU = 0, SP_H = 1 };

/* pair kind from the two species, :559-568 */
static int pair_kind(int zi, int zj) {
  if (zi == SP_CU && zj == SP_CU) return 0;
  if (zi == SP_H && zj == SP_H) return 2;
  return 1;
}

int oracle_cuh2_force_table(int32_t natoms, const int32_t *Z, const oracle_nlist_t *tp, double *F,
                            double *energy, char *err, size_t errlen) {
  const oracle_nlist_t t = *tp;
  if (err && errlen) err[0] = 0;
  *energy = 0.0;
  for (long i = 0; i < 3L * (natoms > 0 ? natoms : 0); ++i) F[i] = 0.0;
  if (natoms < 1) return 0;

  /* classify :463-489 (one-based atom number in the message) */
  int *species = malloc(sizeof(int) * (size_t)natoms);
  for (int32_t i = 0; i < natoms; ++i) {
    if (Z[i] == 29)
      species[i] = SP_CU;
    else if (Z[i] == 1)
      species[i] = SP_H;
    else {
      snprintf(err, errlen,
               "rgpot_cuh2: atom %d has atomic number %d; this potential covers Cu (29) and H "
               "(1) only",
               i + 1, Z[i]);
      free(species);
      return 2;
    }
  }
  const params_t par = cut_shifted(&kDefaults); /* :415 */

  /* coincident-pair guard :420-429 */
  for (long s = 0; s < t.n_pairs; ++s)
    if (t.dist[s] <= 0.0) {
      snprintf(err, errlen, "rgpot_cuh2: two atoms occupy the same position");
      free(species);
      return 3;
    }

  double *rho_i = malloc(sizeof(double) * (size_t)natoms);
  double *dfd = malloc(sizeof(double) * (size_t)natoms);
  double *e_pair = malloc(sizeof(double) * (size_t)natoms);
  double *e_emb = malloc(sizeof(double) * (size_t)natoms);

  /* pass one, atom_density :506-530 */
  for (int32_t i = 0; i < natoms; ++i) {
    double acc = 0.0;
    for (long s = t.row[i]; s < t.row[i + 1]; ++s) {
      const double r = t.dist[s];
      if (r >= par.rcut) continue;
      acc = acc + rho(&par, species[t.idx[s]], r);
    }
    rho_i[i] = acc;
  }
  /* pass two :443-451 */
  for (int32_t i = 0; i < natoms; ++i) {
    if (species[i] == SP_CU) {
      e_emb[i] = f_cu(&par, rho_i[i]);
      dfd[i] = df_cu(&par, rho_i[i]);
    } else {
      e_emb[i] = f_h(&par, rho_i[i]);
      dfd[i] = df_h(&par, rho_i[i]);
    }
  }
  /* pass three, atom_force :536-587 */
  for (int32_t i = 0; i < natoms; ++i) {
    double e = 0.0, f[3] = {0.0, 0.0, 0.0};
    const int zi = species[i];
    for (long s = t.row[i]; s < t.row[i + 1]; ++s) {
      const double r = t.dist[s];
      if (r >= par.rcut) continue;
      const int32_t j = t.idx[s];
      const int zj = species[j];
      const int k = pair_kind(zi, zj);
      e = e + 0.5 * phi(&par, k, r);
      double w = dphi(&par, k, r);
      const double drho_ij = drho(&par, zj, r);
      const double drho_ji = drho(&par, zi, r);
      w = w + drho_ij * dfd[i] + drho_ji * dfd[j];
      for (int c = 0; c < 3; ++c) f[c] = f[c] + w * t.vec[3 * s + c];
    }
    e_pair[i] = e;
    for (int c = 0; c < 3; ++c) F[3 * i + c] = f[c];
  }
  /* :459 -- two sequential sums */
  double sp = 0.0, se = 0.0;
  for (int32_t i = 0; i < natoms; ++i) sp += e_pair[i];
  for (int32_t i = 0; i < natoms; ++i) se += e_emb[i];
  *energy = sp + se;

  free(rho_i);
  free(dfd);
  free(e_pair);
  free(e_emb);
  free(species);
  return 0;
}

int oracle_cuh2_force(int32_t natoms, const double *pos, const int32_t *Z, const double *cell,
                      double *F, double *energy, char *err, size_t errlen,
                      oracle_nlist_t *table_out) {
  /* rgpot_cuh2_capi.f90:45-50 */
  if (err && errlen) err[0] = 0;
  *energy = 0.0;
  for (long i = 0; i < 3L * (natoms > 0 ? natoms : 0); ++i) F[i] = 0.0;
  if (table_out) memset(table_out, 0, sizeof(*table_out));
  if (natoms < 1) return 0;
  /* classify runs before the table build (:412-417): report a foreign species first */
  for (int32_t i = 0; i < natoms; ++i)
    if (Z[i] != 29 && Z[i] != 1) {
      oracle_nlist_t empty;
      memset(&empty, 0, sizeof empty);
      return oracle_cuh2_force_table(natoms, Z, &empty, F, energy, err, errlen);
    }
  /* table%build :417 -> rgpot_neighbors.f90:57-161, skin = 0.1 * cutoff (:90) */
  oracle_nlist_t t;
  char verr[256];
  int st = oracle_nlist_build(natoms, pos, cell, kDefaults.rcut, 0.1 * kDefaults.rcut, &t, verr,
                              sizeof verr);
  if (st != 0) {
    snprintf(err, errlen, "rgpot_neighbors: vesin build failed: %s", verr);
    return st;
  }
  st = oracle_cuh2_force_table(natoms, Z, &t, F, energy, err, errlen);
  if (table_out && st == 0)
    *table_out = t;
  else
    oracle_nlist_free(&t);
  return st;
}
