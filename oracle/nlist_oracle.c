/*
 * nlist_oracle.c -- CPU restatement of the neighbor search under rgpot's CuH2 path:
 * vesin's cell list + Verlet candidate filter, then rgpot's CSR table.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity: PINNED against the vendored vesin
 * compiled into oracle/_ref (tests/test_oracle_vs_ref.py, bit-exact pair sets).
 *
 * All line numbers: third_party/vesin/vesin-single-build.cpp unless noted.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  double m[3][3];   /* rows = cell vectors */
  double inv[3][3]; /* :108-128 */
  double face[3];   /* :271-275 distances between faces */
} bbox_t;

static void cross(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
/* Vector::normalize :53-55 : v * (1 / norm) */
static void normalize(double v[3]) {
  const double s = 1.0 / sqrt(dot3(v, v));
  v[0] = s * v[0];
  v[1] = s * v[1];
  v[2] = s * v[2];
}

/* BoundingBox ctor for a fully periodic cell, :158-276 ; Matrix::determinant/inverse
 * :100-128. */
static int bbox_init(bbox_t *b, const double *box, char *err, size_t errlen) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) b->m[i][j] = box[3 * i + j];
  double(*m)[3] = b->m;
  const double det = m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) -
                     m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                     m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  if (fabs(det) < 1e-30) {
    snprintf(err, errlen, "the box matrix is not invertible");
    return 1;
  }
  b->inv[0][0] = (m[1][1] * m[2][2] - m[2][1] * m[1][2]) / det;
  b->inv[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / det;
  b->inv[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det;
  b->inv[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / det;
  b->inv[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det;
  b->inv[1][2] = (m[1][0] * m[0][2] - m[0][0] * m[1][2]) / det;
  b->inv[2][0] = (m[1][0] * m[2][1] - m[2][0] * m[1][1]) / det;
  b->inv[2][1] = (m[2][0] * m[0][1] - m[0][0] * m[2][1]) / det;
  b->inv[2][2] = (m[0][0] * m[1][1] - m[1][0] * m[0][1]) / det;
  double na[3], nb[3], nc[3];
  cross(m[1], m[2], na);
  cross(m[2], m[0], nb);
  cross(m[0], m[1], nc);
  normalize(na);
  normalize(nb);
  normalize(nc);
  b->face[0] = fabs(dot3(na, m[0]));
  b->face[1] = fabs(dot3(nb, m[1]));
  b->face[2] = fabs(dot3(nc, m[2]));
  return 0;
}

/* divmod :1322-1332, python convention. */
static void divmod32(int32_t a, int32_t b, int32_t *q, int32_t *r) {
  *q = a / b;
  *r = a % b;
  if (*r < 0) {
    *r += b;
    *q -= 1;
  }
}

typedef struct {
  int32_t idx;
  int32_t s[3];
} cpoint_t;

typedef struct {
  long len, cap;
  int32_t *ij; /* 2 per entry */
  int32_t *sh; /* 3 per entry */
} cand_t;

static int cand_push(cand_t *c, int32_t i, int32_t j, const int32_t s[3]) {
  if (c->len == c->cap) {
    long nc = c->cap ? 2 * c->cap : 4096;
    int32_t *a = realloc(c->ij, sizeof(int32_t) * 2 * (size_t)nc);
    if (!a) return 1;
    c->ij = a;
    int32_t *b = realloc(c->sh, sizeof(int32_t) * 3 * (size_t)nc);
    if (!b) return 1;
    c->sh = b;
    c->cap = nc;
  }
  c->ij[2 * c->len] = i;
  c->ij[2 * c->len + 1] = j;
  memcpy(&c->sh[3 * c->len], s, 3 * sizeof(int32_t));
  c->len++;
  return 0;
}

/* Separation of one candidate, :1199 / :1972 : (p_j - p_i) + S.cartesian(box), with
 * CellShift::cartesian = vector * matrix (:372-383, :138-143), then dot (:40-42). */
static double candidate_vec(const bbox_t *b, const double *pos, int32_t i, int32_t j,
                            const int32_t s[3], double v[3]) {
  const double s0 = (double)s[0], s1 = (double)s[1], s2 = (double)s[2];
  for (int k = 0; k < 3; ++k) {
    const double sc = s0 * b->m[0][k] + s1 * b->m[1][k] + s2 * b->m[2][k];
    v[k] = (pos[3 * j + k] - pos[3 * i + k]) + sc;
  }
  return v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
}

int oracle_nlist_build(long n, const double *pos, const double *box, double cutoff, double skin,
                       oracle_nlist_t *out, char *err, size_t errlen) {
  memset(out, 0, sizeof(*out));
  if (err && errlen) err[0] = 0;
  /* vesin_neighbors argument checks :2312-2320 */
  if (!isfinite(cutoff) || cutoff <= 0) {
    snprintf(err, errlen, "cutoff must be a finite, positive number");
    return 1;
  }
  bbox_t bb;
  if (bbox_init(&bb, box, err, errlen)) return 1;
  /* make_bounding_for :299-311 -- non-finite coordinates are refused */
  for (long i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k)
      if (!isfinite(pos[3 * i + k])) {
        snprintf(err, errlen, "point %ld has non-finite coordinate along axis %d: %f", i, k,
                 pos[3 * i + k]);
        return 1;
      }

  /* VerletList::rebuild :1914-1948 -- cell list at cutoff + skin */
  const double rc_build = cutoff + skin;
  /* CellList ctor :1345-1419 */
  double nc[3];
  for (int k = 0; k < 3; ++k) {
    nc[k] = trunc(bb.face[k] / rc_build);
    if (nc[k] < 1.0) nc[k] = 1.0;
  }
  if (nc[0] * nc[1] * nc[2] > 1e5) {
    const double rxy = nc[0] / nc[1], ryz = nc[1] / nc[2];
    nc[2] = trunc(cbrt(1e5 / (rxy * ryz * ryz)));
    nc[1] = trunc(ryz * nc[2]);
    nc[0] = trunc(rxy * nc[1]);
  }
  int32_t nsearch[3], shape[3];
  for (int k = 0; k < 3; ++k) {
    nsearch[k] = (int32_t)ceil(rc_build * nc[k] / bb.face[k]);
    if (nsearch[k] < 1) nsearch[k] = 1;
    shape[k] = (int32_t)nc[k];
  }
  const long ncells = (long)shape[0] * shape[1] * shape[2];

  /* add_point :1421-1442 (bucket the points; two passes instead of vectors of vectors) */
  int32_t *cell_of = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  cpoint_t *pts = malloc(sizeof(cpoint_t) * (size_t)(n > 0 ? n : 1));
  long *start = calloc((size_t)ncells + 1, sizeof(long));
  cpoint_t *sorted = malloc(sizeof(cpoint_t) * (size_t)(n > 0 ? n : 1));
  if (!cell_of || !pts || !start || !sorted) {
    snprintf(err, errlen, "failed to allocate memory");
    return 1;
  }
  for (long i = 0; i < n; ++i) {
    const double *p = &pos[3 * i];
    int32_t c[3];
    for (int k = 0; k < 3; ++k) {
      /* cartesian * inverse :138-143, :246 */
      const double f = p[0] * bb.inv[0][k] + p[1] * bb.inv[1][k] + p[2] * bb.inv[2][k];
      int32_t ci = (int32_t)floor(f * (double)shape[k]);
      int32_t q, r;
      divmod32(ci, shape[k], &q, &r);
      pts[i].s[k] = q;
      c[k] = r;
    }
    pts[i].idx = (int32_t)i;
    cell_of[i] = (int32_t)(((long)shape[0] * shape[1] * c[2]) + ((long)shape[0] * c[1]) + c[0]);
    start[cell_of[i] + 1]++;
  }
  for (long c = 0; c < ncells; ++c) start[c + 1] += start[c];
  {
    long *fill = calloc((size_t)ncells, sizeof(long));
    for (long i = 0; i < n; ++i) sorted[start[cell_of[i]] + fill[cell_of[i]]++] = pts[i];
    free(fill);
  }

  /* foreach_pair :1444-1519 with the full-list callback of cell_list_neighbors :1166-1227 */
  cand_t cand = {0, 0, NULL, NULL};
  const double rcb2 = rc_build * rc_build;
  for (long lin = 0; lin < ncells; ++lin) {
    const int32_t cx = (int32_t)(lin % shape[0]);
    const int32_t cy = (int32_t)((lin / shape[0]) % shape[1]);
    const int32_t cz = (int32_t)(lin / ((long)shape[0] * shape[1]));
    if (start[lin] == start[lin + 1]) continue;
    for (int32_t dx = -nsearch[0]; dx <= nsearch[0]; ++dx)
      for (int32_t dy = -nsearch[1]; dy <= nsearch[1]; ++dy)
        for (int32_t dz = -nsearch[2]; dz <= nsearch[2]; ++dz) {
          int32_t cs[3], nb[3];
          divmod32(cx + dx, shape[0], &cs[0], &nb[0]);
          divmod32(cy + dy, shape[1], &cs[1], &nb[1]);
          divmod32(cz + dz, shape[2], &cs[2], &nb[2]);
          const long nlin = ((long)shape[0] * shape[1] * nb[2]) + ((long)shape[0] * nb[1]) + nb[0];
          for (long a = start[lin]; a < start[lin + 1]; ++a)
            for (long bq = start[nlin]; bq < start[nlin + 1]; ++bq) {
              const cpoint_t *pi = &sorted[a], *pj = &sorted[bq];
              int32_t s[3];
              for (int k = 0; k < 3; ++k) s[k] = cs[k] + pi->s[k] - pj->s[k];
              if (pi->idx == pj->idx && s[0] == 0 && s[1] == 0 && s[2] == 0) continue;
              double v[3];
              if (candidate_vec(&bb, pos, pi->idx, pj->idx, s, v) < rcb2) {
                if (cand_push(&cand, pi->idx, pj->idx, s)) {
                  snprintf(err, errlen, "failed to allocate memory");
                  return 1;
                }
              }
            }
        }
  }
  free(cell_of);
  free(pts);
  free(start);
  free(sorted);

  /* VerletList::filter :1950-2000 (exact cutoff, strict <) followed by
   * rgpot_neighbors.f90:126-160 (count, prefix sum, scatter; i == j dropped). */
  const double rc2 = cutoff * cutoff;
  out->row = calloc((size_t)n + 1, sizeof(int32_t));
  unsigned char *keep = malloc((size_t)(cand.len > 0 ? cand.len : 1));
  long total = 0, selfimg = 0;
  for (long k = 0; k < cand.len; ++k) {
    double v[3];
    const int32_t i = cand.ij[2 * k], j = cand.ij[2 * k + 1];
    const double d2 = candidate_vec(&bb, pos, i, j, &cand.sh[3 * k], v);
    keep[k] = 0;
    if (d2 < rc2) {
      if (i == j) {
        ++selfimg;
        continue;
      }
      keep[k] = 1;
      out->row[i + 1]++;
      ++total;
    }
  }
  for (long i = 0; i < n; ++i) out->row[i + 1] += out->row[i];
  out->n_pairs = total;
  out->n_self_images = selfimg;
  const size_t cap = (size_t)(total > 0 ? total : 1);
  out->idx = malloc(sizeof(int32_t) * cap);
  out->shift = malloc(sizeof(int32_t) * 3 * cap);
  out->vec = malloc(sizeof(double) * 3 * cap);
  out->dist = malloc(sizeof(double) * cap);
  int32_t *fill = calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  for (long k = 0; k < cand.len; ++k) {
    if (!keep[k]) continue;
    const int32_t i = cand.ij[2 * k], j = cand.ij[2 * k + 1];
    double v[3];
    const double d2 = candidate_vec(&bb, pos, i, j, &cand.sh[3 * k], v);
    const long slot = out->row[i] + fill[i]++;
    out->idx[slot] = j;
    memcpy(&out->shift[3 * slot], &cand.sh[3 * k], 3 * sizeof(int32_t));
    memcpy(&out->vec[3 * slot], v, 3 * sizeof(double));
    out->dist[slot] = sqrt(d2);
  }
  free(fill);
  free(keep);
  free(cand.ij);
  free(cand.sh);
  return 0;
}

void oracle_nlist_free(oracle_nlist_t *t) {
  free(t->row);
  free(t->idx);
  free(t->shift);
  free(t->vec);
  free(t->dist);
  memset(t, 0, sizeof(*t));
}
