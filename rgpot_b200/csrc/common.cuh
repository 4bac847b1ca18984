// common.cuh -- shared types and exact-arithmetic helpers of the B200 engine.
//
// Bit-exact neighbor decisions: the reference is compiled for baseline x86-64 (no FMA), so
// every operation that feeds an accept/reject test is written with __dadd_rn / __dmul_rn,
// which nvcc never contracts.  Everything after the decision (the physics) may use FMA.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rb {

#define RB_HD __host__ __device__ __forceinline__
#define RB_D __device__ __forceinline__

constexpr int kKindLJ = 0;
constexpr int kKindLJCluster = 1;
constexpr int kKindCuH2 = 2;

// Status flag bits raised by kernels (device word), mapped to ABI statuses on the host.
constexpr unsigned kFlagNonFinite = 1u;    // -> status 1 (vesin: non-finite coordinate)
constexpr unsigned kFlagForeignZ = 2u;     // -> status 2
constexpr unsigned kFlagCoincident = 4u;   // -> status 3
constexpr unsigned kFlagShiftRange = 8u;   // atom further than +-511 cells outside the box
constexpr unsigned kFlagListOverflow = 16u;  // per-system, internal: retry on the slow path
constexpr unsigned kFlagAtomShift = 32u;   // informational: some atom sits outside the cell
constexpr unsigned kFlagBigShift = 128u;    // informational: some atom sits more than one cell vector outside
constexpr unsigned kFlagSingularBox = 64u;  // -> status 1 (vesin: box matrix not invertible)

struct Status {
  unsigned flags;
  int first_bad_atom;  // smallest original index that raised NonFinite / ForeignZ (else INT_MAX)
  int bad_value;       // its atomic number (ForeignZ)
  int slow_atoms;      // diagnostics: atoms a brick kernel evaluated on its slow path
  int max_halo;        // feedback: largest halo population a brick kernel met (sizes the next call)
  int ticket;          // k_reduce_slices: blocks that have delivered their slice sum
  unsigned long long pool_used;  // survivor-list pool of the EAM brick kernels: entries handed out
  unsigned long long prof[4];  // -DRB_BRICK_TIMING builds: warp-clock sums of setup / stage / search / physics
};

// Geometry of one periodic (or free) cell, prepared on the host with the reference's own
// operation order (vesin-single-build.cpp:100-143, 158-276).
struct Box {
  double H[9];     // rows = cell vectors (LJ reference mode: diag(box[0], box[4], box[8]))
  double Hinv[9];  // inverse, vesin's cofactor formula
  double face[3];  // distances between faces
  double lo[3];    // bounding-box origin of non-periodic directions
  double ext[3];   // bounding-box extent of non-periodic directions
  int periodic[3];
  int pad;
};

// Cell grid: nc cells per direction, stencil half-widths ns, nearest-image dedupe ranges.
struct Grid {
  int nc[3];
  int ns_lo[3];  // stencil runs over delta in [-ns_lo, +ns_hi]
  int ns_hi[3];
  int ncells;
};

// sorted-atom record: position + packed metadata (original index, species, wrap shift).
// meta: bits 0..27 original index (< 2^28 atoms) | 28..31 species / type index (16) |
//       32..41, 42..51, 52..61 shift + 512.
constexpr int kMaxAtoms = (1 << 28) - 1;
RB_HD unsigned long long pack_meta(int orig, int species, int s0, int s1, int s2) {
  return (unsigned long long)((unsigned)orig & 0x0fffffffu) | ((unsigned long long)(species & 15) << 28) |
         ((unsigned long long)((s0 + 512) & 1023) << 32) |
         ((unsigned long long)((s1 + 512) & 1023) << 42) |
         ((unsigned long long)((s2 + 512) & 1023) << 52);
}
RB_HD int meta_orig(unsigned long long m) { return (int)(m & 0x0fffffffull); }
RB_HD int meta_species(unsigned long long m) { return (int)((m >> 28) & 15); }
RB_HD int meta_shift(unsigned long long m, int k) { return (int)((m >> (32 + 10 * k)) & 1023) - 512; }

// compile-time flag passed to generic lambdas (two specialisations of one loop body)
template <bool B>
struct BoolTag {
  static constexpr bool value = B;
};

// ---- exact (never-contracted) arithmetic -------------------------------------------------
RB_D double xadd(double a, double b) { return __dadd_rn(a, b); }
RB_D double xsub(double a, double b) { return __dadd_rn(a, -b); }
RB_D double xmul(double a, double b) { return __dmul_rn(a, b); }

// dx*dx + dy*dy + dz*dz, left to right, separately rounded (vesin Vector::dot :40-42,
// vesin_visit.hpp:103).
RB_D double xnorm2(double dx, double dy, double dz) {
  return xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
}

// CellShift::cartesian component k: s0*H[0][k] + s1*H[1][k] + s2*H[2][k]
// (vesin-single-build.cpp:372-383 via operator*(Vector, Matrix) :138-143).
RB_D double shift_component(const double* __restrict__ H, int k, double s0, double s1, double s2) {
  return xadd(xadd(xmul(s0, H[k]), xmul(s1, H[3 + k])), xmul(s2, H[6 + k]));
}

// stage stamps of the single-system kernels (tuning builds only: -DRB_STAMPS, tools/build_stamps.sh)
#ifdef RB_STAMPS
static __device__ unsigned long long g_stamps[32];
#define RB_STAMP(k)                                                   \
  do {                                                                \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                        \
      unsigned long long t_;                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)::"memory"); \
      g_stamps[k] = t_;                                               \
    }                                                                 \
  } while (0)
// (any block: the caller restricts the thread)
#define RB_STAMP_ANY(k)                                             \
  do {                                                              \
    unsigned long long t_;                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)::"memory"); \
    g_stamps[k] = t_;                                               \
  } while (0)
#else
#define RB_STAMP(k) \
  do {              \
  } while (0)
#define RB_STAMP_ANY(k) \
  do {                  \
  } while (0)
#endif

// vesin_visit.hpp:23-25 fold: d - w * (double)(long long)(d*inv + copysign(0.5, d*inv)).
RB_D double fold_component(double d, double w, double inv) {
  const double t = xmul(d, inv);
  const double n = (double)(long long)xadd(t, copysign(0.5, t));
  return xsub(d, xmul(w, n));
}

// python-convention divmod (vesin-single-build.cpp:1322-1332)
RB_HD void divmod_i(int a, int b, int& q, int& r) {
  q = a / b;
  r = a % b;
  if (r < 0) {
    r += b;
    q -= 1;
  }
}

// parameters of the pair potentials that share the LJ pair search (LJ, LJCluster, Morse)
struct LJParams {
  double u0, psi2, shiftU, rc2;
  double w[3], inv[3];  // fold widths / inverse widths (inv = 0 disables the fold)
  int morse;            // 0: shifted 12-6 LJ, 1: shifted Morse, 2: ZBL
  int zbl_nt;           // ZBL: number of species types
  double m_de, m_a, m_re, m_ecut, m_two_de_a;  // MorsePot.hpp:56-61, MorsePot.cc:51
  const double* zbl_table;  // ZBL: nt x nt x 10 coefficients (d1..d4, zze, sw1..sw5), ZBLPot.hpp:35-46
  double zbl_rin, zbl_rin2;  // ZBL: switching starts above this (squared) distance
};

}  // namespace rb
