/* fehe_oracle.c -- CPU restatement of rgpot's Fe-He embedded-atom potential.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path (tests/, __graft_entry__.smoke(),
 * bench.py's CPU legs).  Nothing under rgpot_b200/ may include, link or call it.
 *
 * Follows CppCore/rgpot/fortran/rgpot_fehe.f90 function by function (the Fortran kernel cannot be
 * compiled in this image: no gfortran / flang):
 *   parameters :66-156        cutoff :163-169
 *   pair_fe_fe / dpair_fe_fe :182-227     pair_fe_he / dpair :230-249     pair_he_he / dpair :253-304
 *   density_dband / d :311-332            density_sband / d :335-355
 *   embed_dband / d :363-388              embed_sband / d :391-413
 *   pair_terms :421-438   pair_densities :442-456   pair_density_derivs :459-473
 *   atom_densities :480-500   atom_contribution :509-538   map_species :541-569   driver :573-620
 * and rgpot_fehe_capi.f90:36-63 for the C entry.  The neighbor table is rgpot_neighbors.f90:57-161
 * over the vendored vesin (oracle_nlist_build), cutoff 5.4 (the largest of the five), skin 0.54.
 * Integer powers are libgfortran's square-and-multiply.
 *
 * Pins: CppCore/tests/FortranPotsTest.cc:146-155 (eOn's FeHeTest: E = -43.959774 to 1e-4, max |F| =
 * 1.245094 to 1e-3, no net force) and the physics checks of test_fehe.f90 (translation invariance,
 * net force, central differences with He in the cell, foreign species) re-run in
 * tests/test_oracle_golden.py.  Parity status: pinned only to those tolerances.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

enum { FE = 1, HE = 2 };

static const double bohr_radius = 0.529177210818181818;
static const double rho_floor = 1.0e-35;

static const double rcut_fe_pair = 5.3, rcut_fe_rho = 4.2, rcut_fehe_pair = 3.9028, rcut_fehe_rho = 4.10,
                    rcut_he_pair = 5.4;
static const double fe_zbl_scale = 9.7342365892908e+03;
static const double fe_zbl_weight[4] = {0.18180, 0.50990, 0.28020, 0.02817};
static const double fe_zbl_decay[4] = {2.8616724320005e+01, 8.4267310396064e+00, 3.6030244464156e+00,
                                       1.8028536321603e+00};
static const double fe_bridge[4] = {7.4122709384068e+00, -6.4180690713367e-01, -2.6043547961722e+00,
                                    6.2625393931230e-01};
static const double fe_bridge_lo = 1.0, fe_bridge_hi = 2.05;
static const double fe_pair_knot[13] = {2.2, 2.3, 2.4, 2.5, 2.6, 2.7, 2.8, 3.0, 3.3, 3.7, 4.2, 4.7, 5.3};
static const double fe_pair_coeff[13] = {
    -2.7444805994228e+01, 1.5738054058489e+01,  2.2077118733936e+00,  -2.4989799053251e+00, 4.2099676494795e+00,
    -7.7361294129713e-01, 8.0656414937789e-01,  -2.3194358924605e+00, 2.6577406128280e+00,  -1.0260416933564e+00,
    3.5018615891957e-01,  -5.8531821042271e-02, -3.0458824556234e-03};
static const double fe_rho_knot[3] = {2.4, 3.2, 4.2};
static const double fe_rho_coeff[3] = {1.1686859407970e+01, -1.4710740098830e-02, 4.7193527075943e-01};
static const double fe_emb_quadratic = -6.7314115586063e-04, fe_emb_quartic = 7.6514905604792e-08;
static const double fehe_knot[8] = {1.5440, 1.6155, 1.6896, 1.8017, 2.0482, 2.3816, 3.5067, 3.9028};
static const double fehe_coeff[8] = {559.804426025391, -45.916354995728, 35.550312671661, 164.319865173340,
                                     -1.727464405060,  0.106771826237,   0.073715269849,  0.038235287677};
static const double sband_amp = 20.0;
#define SBAND_DECAY (1.5312426703208 / bohr_radius)
static const double sband_emb_sqrt = 0.220813420485, sband_emb_quadratic = 1.367508764130,
                    sband_emb_quartic = 3.382256025271;
static const double he_rm = 2.9683, he_c6 = 1.35186623, he_c8 = 0.41495143, he_c10 = 0.17151143,
                    he_aa = 186924.404, he_alpha = 10.5717543, he_beta = -2.07758779, he_damp = 1.438;
#define HE_EPS (10.956 * 1.380658e-23 / 1.602177e-19)

static double powi(double x, int n) {
  double y = (n & 1) ? x : 1.0;
  while (n >>= 1) {
    x = x * x;
    if (n & 1) y = y * x;
  }
  return y;
}
static double pos0(double v) { return v > 0.0 ? v : 0.0; } /* max(v, 0) */

static double pair_fe_fe(double r) {
  if (r > rcut_fe_pair) return 0.0;
  if (r < fe_bridge_lo) {
    double s = 0.0;
    for (int k = 0; k < 4; ++k) s += fe_zbl_weight[k] * exp(-fe_zbl_decay[k] * r);
    return fe_zbl_scale / r * s;
  } else if (r < fe_bridge_hi) {
    return exp(fe_bridge[0] + fe_bridge[1] * r + fe_bridge[2] * r * r + fe_bridge[3] * r * r * r);
  }
  double s = 0.0;
  for (int k = 0; k < 13; ++k) s += fe_pair_coeff[k] * powi(pos0(fe_pair_knot[k] - r), 3);
  return s;
}
static double dpair_fe_fe(double r) {
  if (r > rcut_fe_pair) return 0.0;
  if (r < fe_bridge_lo) {
    double screen[4], s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < 4; ++k) screen[k] = fe_zbl_weight[k] * exp(-fe_zbl_decay[k] * r);
    for (int k = 0; k < 4; ++k) s1 += screen[k];
    for (int k = 0; k < 4; ++k) s2 += fe_zbl_decay[k] * screen[k];
    return -fe_zbl_scale / (r * r) * s1 - fe_zbl_scale / r * s2;
  } else if (r < fe_bridge_hi) {
    const double expo = exp(fe_bridge[0] + fe_bridge[1] * r + fe_bridge[2] * r * r + fe_bridge[3] * r * r * r);
    return expo * (fe_bridge[1] + 2.0 * fe_bridge[2] * r + 3.0 * fe_bridge[3] * r * r);
  }
  double s = 0.0;
  for (int k = 0; k < 13; ++k) s += fe_pair_coeff[k] * powi(pos0(fe_pair_knot[k] - r), 2);
  return -3.0 * s;
}
static double pair_fe_he(double r) {
  if (r > rcut_fehe_pair) return 0.0;
  double s = 0.0;
  for (int k = 0; k < 8; ++k) s += fehe_coeff[k] * powi(pos0(fehe_knot[k] - r), 3);
  return s;
}
static double dpair_fe_he(double r) {
  if (r > rcut_fehe_pair) return 0.0;
  double s = 0.0;
  for (int k = 0; k < 8; ++k) s += fehe_coeff[k] * powi(pos0(fehe_knot[k] - r), 2);
  return -3.0 * s;
}
static double pair_he_he(double r) {
  if (r > rcut_he_pair) return 0.0;
  const double x = r / he_rm;
  const double dispersion = he_c6 / powi(x, 6) + he_c8 / powi(x, 8) + he_c10 / powi(x, 10);
  const double core = he_aa * exp(-he_alpha * x + he_beta * x * x);
  double damp = 1.0;
  if (x < he_damp) damp = exp(-powi(he_damp / x - 1.0, 2));
  return HE_EPS * (core - damp * dispersion);
}
static double dpair_he_he(double r) {
  if (r > rcut_he_pair) return 0.0;
  const double x = r / he_rm;
  const double dispersion = he_c6 / powi(x, 6) + he_c8 / powi(x, 8) + he_c10 / powi(x, 10);
  const double ddispersion = -(6.0 * he_c6 / powi(x, 7) + 8.0 * he_c8 / powi(x, 9) + 10.0 * he_c10 / powi(x, 11));
  const double core = he_aa * exp(-he_alpha * x + he_beta * x * x);
  double damp = 1.0, ddamp = 0.0;
  if (x < he_damp) {
    damp = exp(-powi(he_damp / x - 1.0, 2));
    ddamp = damp * 2.0 * he_damp * (he_damp / x - 1.0) / (x * x);
  }
  return HE_EPS * (core * (-he_alpha + 2.0 * he_beta * x) - ddamp * dispersion - damp * ddispersion) / he_rm;
}
static double density_dband(double r) {
  if (r > rcut_fe_rho) return 0.0;
  double s = 0.0;
  for (int k = 0; k < 3; ++k) s += fe_rho_coeff[k] * powi(pos0(fe_rho_knot[k] - r), 3);
  return s;
}
static double ddensity_dband(double r) {
  if (r > rcut_fe_rho) return 0.0;
  double s = 0.0;
  for (int k = 0; k < 3; ++k) s += fe_rho_coeff[k] * powi(pos0(fe_rho_knot[k] - r), 2);
  return -3.0 * s;
}
static double density_sband(double r) {
  if (r > rcut_fehe_rho) return 0.0;
  return sband_amp * powi(r, 3) * exp(-2.0 * SBAND_DECAY * r);
}
static double ddensity_sband(double r) {
  if (r > rcut_fehe_rho) return 0.0;
  return sband_amp * powi(r, 2) * exp(-2.0 * SBAND_DECAY * r) * (3.0 - 2.0 * SBAND_DECAY * r);
}
static double embed_dband(int species, double rho) {
  if (species != FE || rho < rho_floor) return 0.0;
  return -sqrt(rho) + fe_emb_quadratic * powi(rho, 2) + fe_emb_quartic * powi(rho, 4);
}
static double dembed_dband(int species, double rho) {
  if (species != FE || rho < rho_floor) return 0.0;
  return -0.5 / sqrt(rho) + 2.0 * fe_emb_quadratic * rho + 4.0 * fe_emb_quartic * powi(rho, 3);
}
static double embed_sband(double rho) {
  if (rho < rho_floor) return 0.0;
  return sband_emb_sqrt * sqrt(rho) + sband_emb_quadratic * powi(rho, 2) + sband_emb_quartic * powi(rho, 4);
}
static double dembed_sband(double rho) {
  if (rho < rho_floor) return 0.0;
  return 0.5 * sband_emb_sqrt / sqrt(rho) + 2.0 * sband_emb_quadratic * rho + 4.0 * sband_emb_quartic * powi(rho, 3);
}

/* which: 0/1 pair_fe_fe / d, 2/3 pair_fe_he / d, 4/5 pair_he_he / d, 6/7 density_dband / d,
 * 8/9 density_sband / d (of r); 10/11 embed_dband / d (Fe), 12/13 embed_sband / d (of rho) */
double oracle_fehe_scalar(int which, double x) {
  switch (which) {
    case 0: return pair_fe_fe(x);
    case 1: return dpair_fe_fe(x);
    case 2: return pair_fe_he(x);
    case 3: return dpair_fe_he(x);
    case 4: return pair_he_he(x);
    case 5: return dpair_he_he(x);
    case 6: return density_dband(x);
    case 7: return ddensity_dband(x);
    case 8: return density_sband(x);
    case 9: return ddensity_sband(x);
    case 10: return embed_dband(FE, x);
    case 11: return dembed_dband(FE, x);
    case 12: return embed_sband(x);
    default: return dembed_sband(x);
  }
}

static void pair_terms(int si, int sj, double r, double *v, double *dvdr) {
  if (si == FE && sj == FE) {
    *v = pair_fe_fe(r);
    *dvdr = dpair_fe_fe(r);
  } else if (si == HE && sj == HE) {
    *v = pair_he_he(r);
    *dvdr = dpair_he_he(r);
  } else {
    *v = pair_fe_he(r);
    *dvdr = dpair_fe_he(r);
  }
}

int oracle_fehe_force(int32_t natoms, const double *pos, const int32_t *Z, const double *cell, double *F,
                      double *energy, char *err, size_t errlen) {
  /* rgpot_fehe_capi.f90:50-56 */
  if (err && errlen) err[0] = 0;
  *energy = 0.0;
  for (long i = 0; i < 3L * (natoms > 0 ? natoms : 0); ++i) F[i] = 0.0;
  if (natoms < 1) return 0;
  /* map_species :541-569 */
  int *species = malloc(sizeof(int) * (size_t)natoms);
  for (int32_t i = 0; i < natoms; ++i) {
    if (Z[i] == 26)
      species[i] = FE;
    else if (Z[i] == 2)
      species[i] = HE;
    else {
      snprintf(err, errlen,
               "rgpot_fehe: unsupported atomic number %d on atom %d; this potential covers iron (26) and helium "
               "(2) only",
               Z[i], i + 1);
      free(species);
      return 2;
    }
  }
  oracle_nlist_t t;
  char verr[256];
  int st = oracle_nlist_build(natoms, pos, cell, rcut_he_pair, 0.1 * rcut_he_pair, &t, verr, sizeof verr);
  if (st != 0) {
    snprintf(err, errlen, "rgpot_neighbors: vesin build failed: %s", verr);
    free(species);
    return st;
  }
  double *rho_d = malloc(sizeof(double) * (size_t)natoms), *rho_s = malloc(sizeof(double) * (size_t)natoms);
  double *dfd = malloc(sizeof(double) * (size_t)natoms), *dfs = malloc(sizeof(double) * (size_t)natoms);
  double *e_embed = malloc(sizeof(double) * (size_t)natoms), *e_pair = malloc(sizeof(double) * (size_t)natoms);
  /* atom_densities :480-500 */
  for (int32_t i = 0; i < natoms; ++i) {
    double d = 0.0, s = 0.0;
    for (long k = t.row[i]; k < t.row[i + 1]; ++k) {
      const int sj = species[t.idx[k]];
      double pd = 0.0, ps = 0.0;
      if (species[i] == FE && sj == FE)
        pd = density_dband(t.dist[k]);
      else if (species[i] != sj)
        ps = density_sband(t.dist[k]);
      d = d + pd;
      s = s + ps;
    }
    rho_d[i] = d;
    rho_s[i] = s;
  }
  for (int32_t i = 0; i < natoms; ++i) {
    e_embed[i] = embed_dband(species[i], rho_d[i]) + embed_sband(rho_s[i]);
    dfd[i] = dembed_dband(species[i], rho_d[i]);
    dfs[i] = dembed_sband(rho_s[i]);
  }
  /* atom_contribution :509-538 */
  for (int32_t i = 0; i < natoms; ++i) {
    double e = 0.0, f[3] = {0.0, 0.0, 0.0};
    for (long k = t.row[i]; k < t.row[i + 1]; ++k) {
      const int32_t j = t.idx[k];
      const double r = t.dist[k];
      double v, dvdr, drho_d = 0.0, drho_s = 0.0;
      pair_terms(species[i], species[j], r, &v, &dvdr);
      if (species[i] == FE && species[j] == FE)
        drho_d = ddensity_dband(r);
      else if (species[i] != species[j])
        drho_s = ddensity_sband(r);
      e = e + 0.5 * v;
      const double radial = dvdr + (dfd[i] + dfd[j]) * drho_d + (dfs[i] + dfs[j]) * drho_s;
      for (int c = 0; c < 3; ++c) f[c] = f[c] + radial * (t.vec[3 * k + c] / r);
    }
    e_pair[i] = e;
    for (int c = 0; c < 3; ++c) F[3 * i + c] = f[c];
  }
  double sp = 0.0, se = 0.0;
  for (int32_t i = 0; i < natoms; ++i) sp += e_pair[i];
  for (int32_t i = 0; i < natoms; ++i) se += e_embed[i];
  *energy = sp + se;
  free(rho_d);
  free(rho_s);
  free(dfd);
  free(dfs);
  free(e_embed);
  free(e_pair);
  free(species);
  oracle_nlist_free(&t);
  return 0;
}
