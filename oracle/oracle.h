/*
 * oracle.h -- CPU restatement of rgpot's LJ / CuH2 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load liboracle.so.  The
 * product path (rgpot_b200/csrc) never links or calls it and has no CPU fallback.
 *
 * Every function cites the reference file:line (relative to the rgpot checkout) whose
 * arithmetic it restates.  Parity status: PINNED -- see tests/test_oracle_golden.py
 * (reference golden vectors) and tests/test_oracle_vs_ref.py (the reference's own
 * LJPot / vesin objects compiled into oracle/_ref).
 *
 * Build: plain C99, `-O2 -ffp-contract=off` (the reference is built for baseline
 * x86-64, i.e. without FMA contraction; the pair-set decisions depend on that).
 */
#ifndef RGPOT_B200_ORACLE_H
#define RGPOT_B200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- LJ ------ */

/* LJPot::forceImpl (periodic_mask = 7) and LJClusterPot::forceImpl (periodic_mask = 0,
 * cluster = 1).  CppCore/rgpot/LennardJones/LJPot.cc:38-84, LJClusterPot.cc:33-83,
 * pair scan CppCore/rgpot/nlist/vesin_visit.hpp:82-110.
 * Returns the number of visited (accepted) unordered pairs.                        */
long oracle_lj_force(long n, const double *pos, const double *box, double u0, double cutoff,
                     double psi, int cluster, double *F, double *energy);

/* MorsePot::forceImpl: CppCore/rgpot/Morse/MorsePot.cc:33-78.            */
long oracle_morse_force(long n, const double *pos, const double *box, double De, double a,
                        double re, double cutoff, double *F, double *energy);

/* ZBLPot::forceImpl: CppCore/rgpot/ZBL/ZBLPot.cc:172-248.                  */
long oracle_zbl_force(long n, const double *pos, const int *Z, double cut_inner, double cut_global,
                      double *F, double *energy);

/* Same scan, pair set only: writes up to cap (i, j) pairs (i < j) into pairs[2*k..]
 * and returns the total count.  vesin_visit.hpp:82-110.                            */
long oracle_lj_pairs(long n, const double *pos, const double *box, double cutoff, int cluster,
                     int32_t *pairs, long cap);

/* -------------------------------------------------------- neighbor list --- */

typedef struct {
  long n_pairs;       /* directed entries, self-images (i == j) already dropped      */
  int32_t *row;       /* n_atoms + 1, zero-based CSR row starts                       */
  int32_t *idx;       /* neighbor index per entry                                     */
  int32_t *shift;     /* 3 per entry: S such that vec = r_j - r_i + S.H               */
  double *vec;        /* 3 per entry                                                  */
  double *dist;       /* per entry                                                    */
  long n_self_images; /* entries with i == j (dropped by rgpot_neighbors.f90:132)     */
} oracle_nlist_t;

/* vesin full list (cell list at cutoff + skin, filtered at cutoff) turned into the CSR
 * table of rgpot_neighbors.f90:57-161.  third_party/vesin/vesin-single-build.cpp:
 * 152-350 (BoundingBox), 1345-1519 (CellList), 1914-2000 (Verlet rebuild + filter).
 * Fully periodic boxes only (rgpot_neighbors.f90:87).  Returns 0 or a vesin-style
 * non-zero status (singular box, non-finite coordinate) with a message in err.       */
int oracle_nlist_build(long n, const double *pos, const double *box, double cutoff, double skin,
                       oracle_nlist_t *out, char *err, size_t errlen);
void oracle_nlist_free(oracle_nlist_t *t);

/* ------------------------------------------------------------- CuH2 ------- */

/* rgpot_cuh2_force (rgpot_cuh2_capi.f90:31-58) -> cuh2_energy_forces
 * (rgpot_cuh2.f90:381-460).  Status 0 / 1 / 2 / 3 as the reference; message in err.
 * If `table` is non-NULL it receives the neighbor table used (caller frees).         */
int oracle_cuh2_force(int32_t natoms, const double *pos, const int32_t *Z, const double *cell,
                      double *F, double *energy, char *err, size_t errlen,
                      oracle_nlist_t *table);

/* The three passes over a caller-supplied table (lets the reference's own vesin feed
 * the restated physics: oracle/ref_shim.cc).  rgpot_cuh2.f90:412-459.               */
int oracle_cuh2_force_table(int32_t natoms, const int32_t *Z, const oracle_nlist_t *table,
                            double *F, double *energy, char *err, size_t errlen);

/* Scalar pieces, exported so tests can pin them individually (rgpot_cuh2.f90:143-370).
 * kind: 0 CuCu, 1 HCu, 2 HH for phi; 0 Cu, 1 H for rho / embedding.                  */
double oracle_cuh2_phi(int kind, double r);
double oracle_cuh2_dphi(int kind, double r);
double oracle_cuh2_rho(int kind, double r);
double oracle_cuh2_drho(int kind, double r);
double oracle_cuh2_embed(int kind, double rho);
double oracle_cuh2_dembed(int kind, double rho);

/* ------------------------------------------------------------- EAM-Al ----- */

/* rgpot_eam_al_force (rgpot_eam_al_capi.f90:31-58) -> eam_al_energy_forces
 * (rgpot_eam_al.f90:216-258).  Status 0 or a neighbor-build status with a message.    */
int oracle_eam_al_force(int32_t natoms, const double *pos, const double *cell, double *F,
                        double *energy, char *err, size_t errlen);
int oracle_eam_al_force_table(int32_t natoms, const oracle_nlist_t *table, double *F, double *energy);
/* which: 0 pair_energy, 1 pair_deriv, 2 density, 3 density_deriv (of r); 4 embedding,
 * 5 embedding_deriv (of rho).  rgpot_eam_al.f90:125-213.                               */
double oracle_eam_al_scalar(int which, double x);

/* ------------------------------------------------------------- FeHe ------- */

/* rgpot_fehe_force (rgpot_fehe_capi.f90:36-63) -> fehe_energy_forces (rgpot_fehe.f90:573-620).
 * Status 0, 2 (unsupported atomic number) or a neighbor-build status, message in err.      */
int oracle_fehe_force(int32_t natoms, const double *pos, const int32_t *Z, const double *cell, double *F,
                      double *energy, char *err, size_t errlen);
double oracle_fehe_scalar(int which, double x); /* see fehe_oracle.c */

/* LJ pair math over a vesin-semantics (all images, general triclinic) list: the oracle
 * for the opt-in triclinic LJ mode.  There is NO reference implementation of this
 * (LJPot.cc ignores off-diagonal box terms): parity unpinned, self-consistency only.  */
int oracle_lj_force_allimages(long n, const double *pos, const double *box, double u0,
                              double cutoff, double psi, double *F, double *energy, char *err,
                              size_t errlen);

/* --------------------------------------------------- result-cache key --- */

/* XXH3_64bits of xxHash 0.8.3 (seed 0, default secret) -- the hash inside Potential<D>::cacheKey. */
uint64_t oracle_xxh3_64(const void *data, size_t len);
/* Potential<D>::cacheKey, CppCore/rgpot/Potential.hpp:324-336.                                   */
uint64_t oracle_cache_key(size_t n, const double *pos, const int *atmnrs, const double *box, size_t pot_type,
                          uint64_t params_key);

#ifdef __cplusplus
}
#endif
#endif
