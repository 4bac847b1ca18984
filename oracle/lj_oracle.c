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
