/*
 * lj_oracle.c -- CPU restatement of rgpot's shifted 12-6 Lennard-Jones path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity: PINNED against
 * CppCore/tests/LJClusterPotTest.cc:55-112 and the compiled reference (oracle/_ref).
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

/* vesin_visit.hpp:23-25 -- round to nearest through a truncating cast, halves away
 * from zero. */
static double fold_round(double t) { return (double)(long long)(t + copysign(0.5, t)); }

typedef struct {
  double w[3];
  double inv[3];
  double rc2;
} scan_t;

/* PairListCache.hpp:176-184 (fold_params) with LJPot.cc:54-55 / LJClusterPot.cc:46-55
 * (options and the degenerate-cell substitution). */
static void scan_setup(scan_t *s, const double *box, double cutoff, int cluster) {
  static const double free_box[9] = {1e6, 0, 0, 0, 1e6, 0, 0, 0, 1e6};
  const double *b = box;
  double rc = cutoff;
  if (cluster) {
    rc = (cutoff > 0.0) ? cutoff : 1.0e6;
    int ok = box != NULL && box[0] > 0.0 && box[4] > 0.0 && box[8] > 0.0;
    if (!ok) b = free_box;
  }
  for (int k = 0; k < 3; ++k) {
    s->w[k] = b[4 * k];
    s->inv[k] = (!cluster && s->w[k] != 0.0) ? 1.0 / s->w[k] : 0.0;
  }
  s->rc2 = rc * rc;
}

/* vesin_visit.hpp:92-109: one folded separation r_j - r_i, returns r2. */
static double folded(const scan_t *s, const double *p, long i, long j, double d[3]) {
  for (int k = 0; k < 3; ++k) {
    double t = p[3 * j + k] - p[3 * i + k];
    t -= s->w[k] * fold_round(t * s->inv[k]);
    d[k] = t;
  }
  return d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
}

long oracle_lj_force(long n, const double *pos, const double *box, double u0, double cutoff,
                     double psi, int cluster, double *F, double *energy) {
  *energy = 0.0;
  for (long i = 0; i < 3 * n; ++i) F[i] = 0.0;
  /* LJPot.cc:50-52 (LJClusterPot.cc:43-45 has no cutoff early-out). */
  if (n < 2 || (!cluster && cutoff <= 0.0)) return 0;

  /* LJPot.hpp:52-53 */
  const double a6 = pow(psi / cutoff, 6.0);
  const double shiftU = 4.0 * u0 * a6 * (a6 - 1.0);
  const double psi2 = psi * psi;

  scan_t s;
  scan_setup(&s, box, cutoff, cluster);
  double acc = 0.0;
  long visited = 0;
  for (long i = 0; i < n; ++i) {
    for (long j = i + 1; j < n; ++j) {
      double d[3];
      const double r2 = folded(&s, pos, i, j, d);
      if (r2 <= s.rc2) { /* vesin_visit.hpp:104 -- inclusive */
        /* PairListCache.hpp:94-95 hands the lambda d = r_i - r_j. LJPot.cc:61-81. */
        const double invR2 = 1.0 / r2;
        const double sr2 = psi2 * invR2;
        const double a = sr2 * sr2 * sr2;
        const double b = 4.0 * u0 * a;
        acc += b * (a - 1.0) - shiftU;
        const double fs = 6.0 * b * invR2 * (2.0 * a - 1.0);
        for (int k = 0; k < 3; ++k) {
          const double f = fs * (-d[k]);
          F[3 * i + k] += f;
          F[3 * j + k] -= f;
        }
        ++visited;
      }
    }
  }
  *energy = acc;
  return visited;
}

/* MorsePot::forceImpl, CppCore/rgpot/Morse/MorsePot.cc:33-78 (same scan and fold as LJPot),
 * shift CppCore/rgpot/Morse/MorsePot.hpp:56-61. */
long oracle_morse_force(long n, const double *pos, const double *box, double De, double a,
                        double re, double cutoff, double *F, double *energy) {
  *energy = 0.0;
  for (long i = 0; i < 3 * n; ++i) F[i] = 0.0;
  if (n < 2 || cutoff <= 0.0) return 0; /* MorsePot.cc:43-45 */
  const double dc = 1.0 - exp(-a * (cutoff - re));
  const double ecut = De * dc * dc - De;
  const double twoDeA = 2.0 * De * a;
  scan_t s;
  scan_setup(&s, box, cutoff, 0);
  double acc = 0.0;
  long visited = 0;
  for (long i = 0; i < n; ++i) {
    for (long j = i + 1; j < n; ++j) {
      double dv[3];
      const double r2 = folded(&s, pos, i, j, dv);
      if (r2 <= s.rc2) {
        const double r = sqrt(r2);
        const double d = 1.0 - exp(-a * (r - re));
        acc += De * d * d - De - ecut;
        const double fs = twoDeA * d * (d - 1.0) / r;
        for (int k = 0; k < 3; ++k) {
          const double f = fs * (-dv[k]);
          F[3 * i + k] += f;
          F[3 * j + k] -= f;
        }
        ++visited;
      }
    }
  }
  *energy = acc;
  return visited;
}

/* ---- ZBL: CppCore/rgpot/ZBL/ZBLPot.cc:20-72 (constants, e_zbl, dzbldr, d2zbldr2),
 * :88-129 (buildTables), :172-248 (forceImpl; free boundaries, r2 <= cut_global^2). */
typedef struct {
  double d1, d2, d3, d4, zze, sw1, sw2, sw3, sw4, sw5;
} zbl_t;
static double zbl_e(const zbl_t *p, double r) {
  const double r_inv = 1.0 / r;
  const double sum = 0.18175 * exp(-p->d1 * r) + 0.50986 * exp(-p->d2 * r) +
                     0.28022 * exp(-p->d3 * r) + 0.02817 * exp(-p->d4 * r);
  return p->zze * sum * r_inv;
}
static double zbl_de(const zbl_t *p, double r) {
  const double r_inv = 1.0 / r;
  const double e1 = exp(-p->d1 * r), e2 = exp(-p->d2 * r), e3 = exp(-p->d3 * r), e4 = exp(-p->d4 * r);
  const double sum = 0.18175 * e1 + 0.50986 * e2 + 0.28022 * e3 + 0.02817 * e4;
  const double sum_p = -0.18175 * p->d1 * e1 - 0.50986 * p->d2 * e2 - 0.28022 * p->d3 * e3 - 0.02817 * p->d4 * e4;
  return p->zze * (sum_p - sum * r_inv) * r_inv;
}
static double zbl_d2e(const zbl_t *p, double r) {
  const double r_inv = 1.0 / r;
  const double e1 = exp(-p->d1 * r), e2 = exp(-p->d2 * r), e3 = exp(-p->d3 * r), e4 = exp(-p->d4 * r);
  const double sum = 0.18175 * e1 + 0.50986 * e2 + 0.28022 * e3 + 0.02817 * e4;
  const double sum_p = 0.18175 * e1 * p->d1 + 0.50986 * e2 * p->d2 + 0.28022 * e3 * p->d3 + 0.02817 * e4 * p->d4;
  const double sum_pp = 0.18175 * e1 * p->d1 * p->d1 + 0.50986 * e2 * p->d2 * p->d2 +
                        0.28022 * e3 * p->d3 * p->d3 + 0.02817 * e4 * p->d4 * p->d4;
  return p->zze * (sum_pp + 2.0 * sum_p * r_inv + 2.0 * sum * r_inv * r_inv) * r_inv;
}
static zbl_t zbl_pair(int zi_, int zj_, double cut_inner, double cut_global) {
  const double zi = (double)zi_, zj = (double)zj_, tc = cut_global - cut_inner;
  const double a_inv = (pow(zi, 0.23) + pow(zj, 0.23)) / 0.46850;
  zbl_t p;
  p.d1 = 3.19980 * a_inv;
  p.d2 = 0.94229 * a_inv;
  p.d3 = 0.40290 * a_inv;
  p.d4 = 0.20162 * a_inv;
  p.zze = zi * zj * 14.399645;
  const double fc = zbl_e(&p, cut_global), fcp = zbl_de(&p, cut_global), fcpp = zbl_d2e(&p, cut_global);
  const double swa = (-3.0 * fcp + tc * fcpp) / (tc * tc);
  const double swb = (2.0 * fcp - tc * fcpp) / (tc * tc * tc);
  p.sw1 = swa;
  p.sw2 = swb;
  p.sw3 = swa / 3.0;
  p.sw4 = swb / 4.0;
  p.sw5 = -fc + (tc / 2.0) * fcp - (tc * tc / 12.0) * fcpp;
  return p;
}

long oracle_zbl_force(long n, const double *pos, const int *Z, double cut_inner, double cut_global,
                      double *F, double *energy) {
  *energy = 0.0;
  for (long i = 0; i < 3 * n; ++i) F[i] = 0.0;
  if (n < 2) return 0; /* ZBLPot.cc:180-182 */
  const double rc2 = cut_global * cut_global, rin2 = cut_inner * cut_inner;
  double acc = 0.0;
  long visited = 0;
  for (long i = 0; i < n; ++i) {
    for (long j = i + 1; j < n; ++j) {
      double d[3];
      for (int k = 0; k < 3; ++k) d[k] = pos[3 * j + k] - pos[3 * i + k]; /* no fold: free boundaries */
      const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      if (r2 <= rc2) {
        /* the table entry is symmetric in (i, j): build it for the ordered pair like buildTables
         * does (i <= j in species order) so the same pow() arguments are used */
        const int za = Z[i] < Z[j] ? Z[i] : Z[j], zb = Z[i] < Z[j] ? Z[j] : Z[i];
        const zbl_t p = zbl_pair(za, zb, cut_inner, cut_global);
        const double r = sqrt(r2);
        double e_pair = zbl_e(&p, r) + p.sw5;
        double dEdr = zbl_de(&p, r);
        if (r2 > rin2) {
          const double t = r - cut_inner;
          e_pair += t * t * t * (p.sw3 + p.sw4 * t);
          dEdr += t * t * (p.sw1 + p.sw2 * t);
        }
        acc += e_pair;
        const double fs = -dEdr / r;
        for (int k = 0; k < 3; ++k) {
          const double f = fs * (-d[k]);
          F[3 * i + k] += f;
          F[3 * j + k] -= f;
        }
        ++visited;
      }
    }
  }
  *energy = acc;
  return visited;
}

long oracle_lj_pairs(long n, const double *pos, const double *box, double cutoff, int cluster,
                     int32_t *pairs, long cap) {
  if (n < 2 || (!cluster && cutoff <= 0.0)) return 0;
  scan_t s;
  scan_setup(&s, box, cutoff, cluster);
  long count = 0;
  for (long i = 0; i < n; ++i) {
    for (long j = i + 1; j < n; ++j) {
      double d[3];
      if (folded(&s, pos, i, j, d) <= s.rc2) {
        if (count < cap) {
          pairs[2 * count] = (int32_t)i;
          pairs[2 * count + 1] = (int32_t)j;
        }
        ++count;
      }
    }
  }
  return count;
}

/* Opt-in triclinic LJ: the pair math of LJPot.cc:61-81 applied to every periodic image
 * inside the cutoff (vesin semantics, d2 < rc2).  No reference counterpart exists:
 * parity unpinned.  Gather form, half energy per directed entry. */
int oracle_lj_force_allimages(long n, const double *pos, const double *box, double u0,
                              double cutoff, double psi, double *F, double *energy, char *err,
                              size_t errlen) {
  *energy = 0.0;
  for (long i = 0; i < 3 * n; ++i) F[i] = 0.0;
  if (n < 1 || cutoff <= 0.0) return 0;
  oracle_nlist_t t;
  int st = oracle_nlist_build(n, pos, box, cutoff, 0.1 * cutoff, &t, err, errlen);
  if (st != 0) return st;
  const double a6 = pow(psi / cutoff, 6.0);
  const double shiftU = 4.0 * u0 * a6 * (a6 - 1.0);
  const double psi2 = psi * psi;
  double acc = 0.0;
  for (long i = 0; i < n; ++i) {
    for (long s = t.row[i]; s < t.row[i + 1]; ++s) {
      const double *v = &t.vec[3 * s];
      const double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      const double invR2 = 1.0 / r2;
      const double sr2 = psi2 * invR2;
      const double a = sr2 * sr2 * sr2;
      const double b = 4.0 * u0 * a;
      acc += 0.5 * (b * (a - 1.0) - shiftU);
      const double fs = 6.0 * b * invR2 * (2.0 * a - 1.0);
      for (int k = 0; k < 3; ++k) F[3 * i + k] += fs * (-v[k]);
    }
  }
  /* self-image entries (i == j) are dropped by oracle_nlist_build like the CuH2 table;
   * they would add energy only.  Documented limitation of the opt-in mode. */
  *energy = acc;
  oracle_nlist_free(&t);
  return 0;
}
