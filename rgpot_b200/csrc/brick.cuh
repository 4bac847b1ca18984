// brick.cuh -- brick kernels for one large system (the >= 100k-atom configurations).
//
// One thread block owns a BRICK of cells of the cell list (cell edge >= rc / ns, ns = 1 or 2, so
// the stencil of a cell is (2 ns + 1)^3 cells; with ns = 2 an atom tests 211 candidates in an LJ
// fluid where 27 cutoff-sized cells offer 364).  Three stages, all in shared memory:
//
//   stage   the atoms of the brick's halo (brick + ns cells on every side, periodic images
//           included) are copied once: FP64 positions, and an FP16 copy of the image-shifted
//           position relative to the brick centre (plus image code and species).  A brick of 108
//           atoms stages 864 candidates (8 per atom); a per-cell walk reads 27+.
//   search  every thread owns an atom (kTPA threads share one and split the stencil rows): FP16
//           prefilter over the atom's own stencil -- the (2 ns + 1)^2 rows of 2 ns + 1 cells that
//           are contiguous in the staged order -- one 8-byte shared load, five half2 operations,
//           a compare and a 16-bit store into the thread's private survivor list (the kernel is
//           bound by shared-memory wavefronts: an 8-byte record costs two per warp load, a 16-byte
//           FP32 record four).  No FP64, no ballots, no queues, no compaction, no block barrier
//           before the next stage; the threshold carries the FP16 error bound.
//   physics the thread walks its list: the separation is redone in FP64 with the REFERENCE's
//           expression ((p_j - p_i) + shift, unfused norm: accept / reject is bit-exact), the pair
//           terms are evaluated and accumulated in registers.  Every lane works on a survivor;
//           the sums have a fixed order (stencil order, then a fixed shuffle tree over the kTPA
//           threads): bit-reproducible, no atomics.
//
// Geometry: atoms up to one cell vector outside the cell are taken (wrap counts folded into the
// image codes); the kernels return at once when the binning flags report an atom further out and
// the gather kernels of gather.cuh take over.  A brick whose halo does not fit the staged capacity is
// cut in two; a single cell that still does not fit is walked straight from global memory by the
// same block (brick_slow_atom: same candidate order, same pair terms, same summation tree, hence
// the same bits); a full private list only means another search / physics round.  The EAM density
// pass hands its lists to the force pass through a global pool, so the search runs once.
#pragma once

#include <cuda_fp16.h>

#include "gather.cuh"

namespace rb {

// TILE_PAIR: Morse / ZBL on the LJ geometry; TILE_AL_*: the EAM passes with the aluminium model;
// TILE_FEHE_*: the Fe-He passes (two density channels, so two F'(rho) per staged atom)
enum {
  TILE_LJ = 0, TILE_EAM_DENSITY = 1, TILE_EAM_FORCE = 2, TILE_PAIR = 3, TILE_AL_DENSITY = 4, TILE_AL_FORCE = 5,
  TILE_FEHE_DENSITY = 6, TILE_FEHE_FORCE = 7
};
constexpr bool tile_is_density(int pot) { return pot == TILE_EAM_DENSITY || pot == TILE_AL_DENSITY || pot == TILE_FEHE_DENSITY; }
constexpr bool tile_is_force(int pot) { return pot == TILE_EAM_FORCE || pot == TILE_AL_FORCE || pot == TILE_FEHE_FORCE; }
constexpr bool tile_is_fehe(int pot) { return pot == TILE_FEHE_DENSITY || pot == TILE_FEHE_FORCE; }
constexpr int tile_model(int pot) { return (pot == TILE_AL_DENSITY || pot == TILE_AL_FORCE) ? EAM_AL : EAM_CUH2; }
// FP64 words staged per halo atom: position, + F'(rho) in a force pass, + the second channel's (Fe-He)
constexpr int tile_words(int pot) { return pot == TILE_FEHE_FORCE ? 5 : (tile_is_force(pot) ? 4 : 3); }

#ifndef RB_PHYS_UNROLL
#define RB_PHYS_UNROLL 1
#endif
constexpr int kPhysUnroll = RB_PHYS_UNROLL;  // survivors per trip of the physics loop (tuning knob)
constexpr int kBrickMaxThreads = 384;  // 2 blocks of 384 or 3 of 256 per SM at <= 85 registers
constexpr int kBrickMaxHalo = 512;  // (4 + 2*2)^3
constexpr int kBrickMaxCells = 64;
constexpr int kShiftCodes = 125;    // image counts q in [-2, 2]^3


struct TileOut {
  double* F;            // forces, original order (LJ, EAM force)
  double* eatom;        // per-atom energies, sorted order (LJ, EAM force): folded in one fixed order
  double* dfd_sorted;   // EAM: F'(rho), sorted order (written by density, read by force)
  double* eemb_sorted;  // EAM: F(rho)
  double* dfs_sorted;   // Fe-He: F'(rho) of the s-band channel
  Status* st;
  // EAM: the density pass hands its survivor lists to the force pass (same plan, same boxes, same
  // thread per atom), which then skips the search.  pool: 16-bit staged indices in warp-interleaved
  // chunks (entry k of lane l at chunk + 32 k + l); lref[atom * TPA + sub] = chunk | count << 40,
  // kNoList where the search has to be repeated (list took several rounds, pool exhausted).
  unsigned short* pool;
  unsigned long long* lref;
  unsigned long long pool_cap;  // entries
  // tile kernels (tile.cu): the pool holds the survivor bit masks of a 16-atom pass, word w of lane l
  // at chunk + 32 w + l; lref[sorted index of the pass's first atom] = chunk, kNoList = search again
  unsigned* poolw;
};
constexpr unsigned long long kNoList = ~0ull;

struct BrickPlan {
  int bd[3];      // cells per brick and direction
  int nb[3];      // bricks per direction
  int ns[3];      // stencil half width per direction (cells)
  int nbricks;
  int cap_cand;   // staged halo atoms a block can hold (multiple of 32)
  int cap_list;   // survivor-list entries per thread
  int tpa;        // threads per atom (1, 2, 4)
  int threads;    // block size
  float thr2f;    // FP32 prefilter threshold (conservative superset of d2 <= rc2)
  unsigned smem;  // dynamic shared memory per block
};

RB_HD unsigned brick_smem_bytes(int cap_cand, int cap_list, int threads, int words) {
  return (unsigned)(cap_cand * (words * sizeof(double) + sizeof(uint2) + sizeof(unsigned short)) +
                    (size_t)cap_list * threads * sizeof(unsigned short) + 16);
}

// 1/a for normal positive a: hardware seed + one cubic Newton step (~1 ulp), no slow path
RB_D double rcp_fast(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a, y, 1.0);  // |e| <= 2^-22 (seed accuracy); y (1 + e + e^2) leaves e^3
  return fma(y, fma(e, e, e), y);
}

// image vector S.H in the reference's arithmetic: w * q per direction for the LJ family (its fold
// subtracts w * round(d / w)), vesin's shift product for the all-image tables
template <bool PAIR_GEO>
RB_D void image_vector(const GatherArgs& a, int q0, int q1, int q2, double sv[3]) {
  if (PAIR_GEO) {
    sv[0] = xmul(a.lj.w[0], (double)q0);
    sv[1] = xmul(a.lj.w[1], (double)q1);
    sv[2] = xmul(a.lj.w[2], (double)q2);
  } else {
    sv[0] = shift_component(a.box.H, 0, (double)q0, (double)q1, (double)q2);
    sv[1] = shift_component(a.box.H, 1, (double)q0, (double)q1, (double)q2);
    sv[2] = shift_component(a.box.H, 2, (double)q0, (double)q1, (double)q2);
  }
}
// an atom's own wrap counts (each in [-1, 1]) as an offset in the 5 x 5 x 5 image-code space
RB_D int wrap_offset(unsigned long long m) { return meta_shift(m, 2) * 25 + meta_shift(m, 1) * 5 + meta_shift(m, 0); }

// ---- pair terms shared by the brick kernel and its slow path ------------------------------------
// One candidate at exact separation (dx, dy, dz), d2 = |d|^2 in the reference's arithmetic; `other`
// is false for the atom itself.  acc: force components (density pass: acc[0] = rho).
template <int POT, class DfdJ, class DfsJ>
RB_D void brick_pair(const GatherArgs& a, double rc2, double rc2_phys, bool other, int zi, int zj, double dfd_i,
                     DfdJ&& dfd_j, double dfs_i, DfsJ&& dfs_j, double dx, double dy, double dz, double d2, double& e,
                     double* acc, bool& coincident) {
  if (tile_is_fehe(POT)) {
    // rgpot_fehe.f90: every branch is a comparison on r = sqrt(d2), the table's IEEE square root
    if (d2 < rc2 && other) {
      const double r = sqrt(d2);
      double rho, drho;
      const int ch = fehe_density(zi, zj, r, rho, drho);  // 0 d-band, 1 s-band, -1 none
      if (tile_is_density(POT)) {  // atom_densities :480-500
        if (ch == 0) acc[0] += rho;
        if (ch == 1) acc[1] += rho;
      } else {  // atom_contribution :509-538
        double v, dv;
        fehe_pair(zi, zj, r, v, dv);
        e += 0.5 * v;
        double radial = dv;
        if (ch == 0) radial += (dfd_i + dfd_j()) * drho;
        if (ch == 1) radial += (dfs_i + dfs_j()) * drho;
        acc[0] += radial * (dx / r);
        acc[1] += radial * (dy / r);
        acc[2] += radial * (dz / r);
      }
    }
    return;
  }
  if (POT == TILE_LJ) {
    if (d2 <= rc2 && other) {  // vesin_visit.hpp:104
      const double invR2 = rcp_fast(d2);
      const double sr2 = a.lj.psi2 * invR2;
      const double a6 = sr2 * sr2 * sr2;
      const double bb = 4.0 * a.lj.u0 * a6;
      e += 0.5 * (bb * (a6 - 1.0) - a.lj.shiftU);
      const double fs = 6.0 * bb * invR2 * (2.0 * a6 - 1.0);
      acc[0] = fma(-fs, dx, acc[0]);
      acc[1] = fma(-fs, dy, acc[1]);
      acc[2] = fma(-fs, dz, acc[2]);
    }
  } else if (POT == TILE_PAIR) {
    if (d2 <= rc2 && other) {
      double u, fs;
      pair_eval(a.lj, d2, zi, zj, u, fs);
      e += 0.5 * u;
      acc[0] = fma(-fs, dx, acc[0]);
      acc[1] = fma(-fs, dy, acc[1]);
      acc[2] = fma(-fs, dz, acc[2]);
    }
  } else if (d2 < rc2 && other) {  // vesin :1202; the atom itself is no neighbour
    if (tile_is_density(POT)) {
      if (d2 == 0.0) coincident = true;
      if (d2 < rc2_phys) acc[0] += eam_density_entry<tile_model(POT)>(zj, d2);
    } else if (d2 < rc2_phys) {
      eam_force_entry<tile_model(POT)>(zi, zj, dfd_i, dfd_j(), dx, dy, dz, d2, e, acc[0], acc[1], acc[2]);
    }
  }
}

// One candidate of the large-system kernels (brick.cuh, tile.cu) at exact separation (dx, dy, dz), d2 = |d|^2 in the reference's
// arithmetic.  For every potential but plain LJ this is brick_pair.  LJ (LJPot.cc:61-81) keeps RAW sums
// and scales once per atom (tile_finish): with a = (psi^2 / r^2)^3 the pair energy is 4 u0 a (a - 1) - U_c
// and the force factor 24 u0 a (2 a - 1) / r^2, so the loop accumulates a (a - 1), the pair count and
// a (2 a - 1) / r^2 * d: three FP64 instructions fewer per pair than forming b = 4 u0 a first, a few ulp
// from the reference's association, far inside the 1e-10 budget; accept / reject is untouched.
template <int POT, class DfdJ, class DfsJ>
RB_D void tile_pair(const GatherArgs& a, double rc2, double rc2_phys, bool other, int zi, int zj, double dfd_i,
                    DfdJ&& dfd_j, double dfs_i, DfsJ&& dfs_j, double dx, double dy, double dz, double d2, double& e,
                    double* acc, int& npair, bool& coincident) {
  if (POT == TILE_LJ) {
    if (d2 <= rc2 && other) {  // vesin_visit.hpp:104
      const double invR2 = rcp_fast(d2);
      const double sr2 = a.lj.psi2 * invR2;
      const double a6 = sr2 * sr2 * sr2;
      e = fma(a6, a6 - 1.0, e);
      ++npair;
      const double t = (a6 * invR2) * fma(2.0, a6, -1.0);
      acc[0] = fma(t, dx, acc[0]);
      acc[1] = fma(t, dy, acc[1]);
      acc[2] = fma(t, dz, acc[2]);
    }
  } else {
    brick_pair<POT>(a, rc2, rc2_phys, other, zi, zj, dfd_i, dfd_j, dfs_i, dfs_j, dx, dy, dz, d2, e, acc, coincident);
  }
}
// raw LJ sums -> half the pair energy and the force on the atom (d = r_j - r_i, so F_i = -24 u0 sum)
template <int POT>
RB_D void tile_finish(const GatherArgs& a, double* acc, double& e, int npair) {
  if (POT == TILE_LJ) {
    const double c = -24.0 * a.lj.u0;
    acc[0] *= c;
    acc[1] *= c;
    acc[2] *= c;
    e = fma(2.0 * a.lj.u0, e, -(0.5 * a.lj.shiftU) * (double)npair);
  }
}

// ---- per-atom walk straight from global memory (the in-kernel slow path) -------------------------
// For the atoms of a single cell whose stencil does not fit the staged capacity.  It visits the
// candidates in the order the brick kernel does -- stencil row `row` belongs to partial sum
// row % TPA, rows ascending, cells along x, atoms in sorted order -- and folds the TPA partial sums
// in the kernel's shuffle tree, so an atom gets the same bits whichever path evaluated it.
template <int POT, int TPA>
__device__ __noinline__ void brick_slow_atom(const GatherArgs& a, const BrickPlan& bp, const TileOut& o,
                                             const double (*svtab)[3], bool ashift, int i, bool& coincident) {
  constexpr bool kPairGeo = POT == TILE_LJ || POT == TILE_PAIR;
  constexpr bool kForce = tile_is_force(POT);
  constexpr bool kDensity = tile_is_density(POT);
  constexpr bool kFeHe = tile_is_fehe(POT);
  constexpr int NV = kDensity ? (kFeHe ? 2 : 1) : 3;
  const Grid& g = a.grid;
  const int cell = a.scell[i];
  const int c[3] = {cell % g.nc[0], (cell / g.nc[0]) % g.nc[1], cell / (g.nc[0] * g.nc[1])};
  const double4 pi = a.spos[i];
  const unsigned long long mi = (unsigned long long)__double_as_longlong(pi.w);
  const int zi = meta_species(mi);
  const double rc2 = kPairGeo ? a.lj.rc2 : a.rc2;
  const double rc2_phys = kPairGeo ? 0.0 : eam_par<tile_model(POT)>().rc2_phys;
  double dfd_i = 0.0, dfs_i = 0.0;
  if (kForce) dfd_i = o.dfd_sorted[i];
  if (kForce && kFeHe) dfs_i = o.dfs_sorted[i];
  const int nrx = 2 * bp.ns[0] + 1, nry = 2 * bp.ns[1] + 1, nrows = nry * (2 * bp.ns[2] + 1);
  double part[TPA][4];
  int npair = 0;
#pragma unroll
  for (int sub = 0; sub < TPA; ++sub) {
    double acc[3] = {0.0, 0.0, 0.0}, e = 0.0;
    for (int row = sub; row < nrows; row += TPA) {
      const int hy = row % nry, hz = row / nry;
      for (int hx = 0; hx < nrx; ++hx) {
        const int hc[3] = {hx, hy, hz};
        int q[3], r[3];
        bool valid = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          q[k] = 0;
          r[k] = c[k] - bp.ns[k] + hc[k];
          if (a.box.periodic[k]) {
            while (r[k] < 0) {
              r[k] += g.nc[k];
              --q[k];
            }
            while (r[k] >= g.nc[k]) {
              r[k] -= g.nc[k];
              ++q[k];
            }
          } else if (r[k] < 0 || r[k] >= g.nc[k]) {
            valid = false;
          }
        }
        if (!valid) continue;
        const int nb = (r[2] * g.nc[1] + r[1]) * g.nc[0] + r[0];
        const int jb = a.starts[nb], je = a.starts[nb + 1];
        const int code = (q[2] + 2) * 25 + (q[1] + 2) * 5 + (q[0] + 2);
        double svx = svtab[code][0], svy = svtab[code][1], svz = svtab[code][2];
        for (int j = jb; j < je; ++j) {
          const double4 pj = a.spos[j];
          if (ashift) {  // image = cell image + s_i - s_j (vesin :1497)
            const unsigned long long mj = (unsigned long long)__double_as_longlong(pj.w);
            double sv[3];
            image_vector<kPairGeo>(a, q[0] + meta_shift(mi, 0) - meta_shift(mj, 0), q[1] + meta_shift(mi, 1) - meta_shift(mj, 1),
                                   q[2] + meta_shift(mi, 2) - meta_shift(mj, 2), sv);
            svx = sv[0];
            svy = sv[1];
            svz = sv[2];
          }
          // (p_j - p_i) + 0 is p_j - p_i: the kernel's unshifted boxes skip the addition
          const double dx = xadd(xsub(pj.x, pi.x), svx), dy = xadd(xsub(pj.y, pi.y), svy),
                       dz = xadd(xsub(pj.z, pi.z), svz);
          const double d2 = xnorm2(dx, dy, dz);
          const int zj = meta_species((unsigned long long)__double_as_longlong(pj.w));
          tile_pair<POT>(a, rc2, rc2_phys, j != i || code != kShiftCodes / 2, zi, zj, dfd_i,
                         [&]() { return o.dfd_sorted[j]; }, dfs_i, [&]() { return o.dfs_sorted[j]; }, dx, dy, dz, d2,
                         e, acc, npair, coincident);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) part[sub][k] = acc[k];
    part[sub][3] = e;
  }
#pragma unroll
  for (int off = 1; off < TPA; off <<= 1)
#pragma unroll
    for (int sub = 0; sub < TPA; sub += 2 * off)
#pragma unroll
      for (int k = 0; k < 4; ++k) part[sub][k] += part[sub + off][k];
  tile_finish<POT>(a, part[0], part[0][3], npair);
  if (kDensity) {
    double emb, demb;
    if (kFeHe) {
      double dems;
      fehe_embed(zi, part[0][0], part[0][1], emb, demb, dems);
      o.dfs_sorted[i] = dems;
    } else {
      eam_embed<tile_model(POT)>(zi, part[0][0], emb, demb);
    }
    o.dfd_sorted[i] = demb;
    o.eemb_sorted[i] = emb;
    return;
  }
  const int orig = meta_orig(mi);
#pragma unroll
  for (int k = 0; k < NV; ++k) o.F[3 * (size_t)orig + k] = part[0][k];
  double e = part[0][3];
  if (kForce) e += o.eemb_sorted[i];
  o.eatom[i] = e;
}

// ---- the brick kernel ------------------------------------------------------------------------
template <int POT, int TPA>
__global__ void __launch_bounds__(kBrickMaxThreads, 2)
k_brick(const __grid_constant__ GatherArgs a, const __grid_constant__ BrickPlan bp,
        const __grid_constant__ TileOut o) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned short hstart[kBrickMaxHalo + 2];  // staged start of every halo cell
  __shared__ int hgstart[kBrickMaxHalo];                // start of the halo cell in the sorted atom array
  __shared__ unsigned char hcode[kBrickMaxHalo];        // image code of the halo cell
  __shared__ double svtab[kShiftCodes][3];              // image vector per code (reference arithmetic)
  __shared__ unsigned short bc_astart[kBrickMaxCells + 2];  // brick-local atom prefix
  __shared__ int s_shifted, s_ntot;
  __shared__ int box_stack[8][6];  // boxes of cells still to do: first cell, extent

  if (a.st_in->flags & kFlagBigShift) return;  // the gather kernels of gather.cuh take over
  // atoms up to one cell vector outside the cell: the wrap count s_j of a staged atom is folded into
  // its image code (q - s_j), that of the atom being evaluated is added pair by pair
  // (re-read where needed instead of being kept in a register through the whole kernel)
#define RB_ASHIFT ((a.st_in->flags & kFlagAtomShift) != 0)
  if (tile_is_density(POT) || tile_is_force(POT)) exp_table_init();
  constexpr bool kPairGeo = POT == TILE_LJ || POT == TILE_PAIR;  // LJ pair search (w * q image vectors)
  constexpr bool kForce = tile_is_force(POT);
  constexpr bool kDensity = tile_is_density(POT);
  constexpr int kModel = tile_model(POT);
  constexpr bool kFeHe = tile_is_fehe(POT);
  constexpr int NV = kDensity ? (kFeHe ? 2 : 1) : 3;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nthreads = blockDim.x;
  const Grid& g = a.grid;

  double* c64 = reinterpret_cast<double*>(smem_raw);  // x, y, z of staged atom s at c64[3 s ..]
  double* c64w = c64 + 3 * bp.cap_cand;               // EAM force pass: F'(rho_j)
  double* c64w2 = c64 + 4 * bp.cap_cand;              // Fe-He force pass: F'(rho_j) of the s-band
  // FP16 offsets from the brick centre (x, y | z, 0): 8 bytes per candidate, two shared-memory
  // wavefronts per warp load where an FP32 record needs four (the search is bound by them)
  uint2* c16 = reinterpret_cast<uint2*>(c64 + tile_words(POT) * bp.cap_cand);
  unsigned short* cinfo = reinterpret_cast<unsigned short*>(c16 + bp.cap_cand);   // image code | species << 8
  unsigned short* lists = cinfo + bp.cap_cand;

  // image vectors; the brick of this block as the first box
  for (int t = tid; t < kShiftCodes; t += nthreads) {
    image_vector<kPairGeo>(a, t % 5 - 2, (t / 5) % 5 - 2, t / 25 - 2, svtab[t]);
  }
  if (tid == 0) {
    const int b[3] = {(int)blockIdx.x % bp.nb[0], ((int)blockIdx.x / bp.nb[0]) % bp.nb[1],
                      (int)blockIdx.x / (bp.nb[0] * bp.nb[1])};
    for (int k = 0; k < 3; ++k) {
      box_stack[0][k] = b[k] * bp.bd[k];
      box_stack[0][3 + k] = min(bp.bd[k], g.nc[k] - b[k] * bp.bd[k]);
    }
  }
  bool coincident = false;
#ifdef RB_BRICK_TIMING
  long long t_mark = clock64();
  unsigned long long t_acc[4] = {0, 0, 0, 0};
#define RB_TICK(slot) do { const long long t_now = clock64(); t_acc[slot] += (unsigned long long)(t_now - t_mark); t_mark = t_now; } while (0)
#else
#define RB_TICK(slot) do { } while (0)
#endif

  // A brick whose halo exceeds the staged capacity (a locally dense region) is cut in two and the
  // halves are done one after the other; every atom sees the same candidates in the same order
  // either way, so the results do not depend on the capacities.
  int sp = 1;
  __syncthreads();
  for (;;) {
    int c0[3], bw[3], hd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      c0[k] = box_stack[sp - 1][k];
      bw[k] = box_stack[sp - 1][3 + k];
      hd[k] = bw[k] + 2 * bp.ns[k];
    }
    const int nh = hd[0] * hd[1] * hd[2];
    const int nbc = bw[0] * bw[1] * bw[2];
    if (tid == 0) s_shifted = 0;
    __syncthreads();

    // ---- halo cells ---------------------------------------------------------------------------------
    const float inv_hd0 = 1.0f / (float)hd[0], inv_hd1 = 1.0f / (float)hd[1];
    for (int h = tid; h < nh; h += nthreads) {
      // h = (hz * hd1 + hy) * hd0 + hx; small numbers: exact float division
      const int hyz = (int)(((float)h + 0.5f) * inv_hd0);
      const int hz = (int)(((float)hyz + 0.5f) * inv_hd1);
      const int hc[3] = {h - hyz * hd[0], hyz - hz * hd[1], hz};
      int q[3], r[3];
      bool valid = true;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        q[k] = 0;
        r[k] = c0[k] - bp.ns[k] + hc[k];
        if (a.box.periodic[k]) {
          while (r[k] < 0) {
            r[k] += g.nc[k];
            --q[k];
          }
          while (r[k] >= g.nc[k]) {
            r[k] -= g.nc[k];
            ++q[k];
          }
        } else if (r[k] < 0 || r[k] >= g.nc[k]) {
          valid = false;
        }
      }
      int hcount = 0, jb = 0;
      if (valid) {
        const int cell = (r[2] * g.nc[1] + r[1]) * g.nc[0] + r[0];
        jb = a.starts[cell];
        hcount = min(a.starts[cell + 1] - jb, 65535);
        if (hcount > 0 && (q[0] | q[1] | q[2]) != 0) s_shifted = 1;
      }
      hgstart[h] = jb;
      hcode[h] = (unsigned char)((q[2] + 2) * 25 + (q[1] + 2) * 5 + (q[0] + 2));
      hstart[h + 1] = (unsigned short)hcount;  // counts first, prefix below
    }
    __syncthreads();
    if (wid == 0) {  // exclusive prefix over <= 512 counts: sixteen per lane
      int sum = 0;
      for (int t = 0; t < 16; ++t) {
        const int h = lane * 16 + t;
        if (h < nh) sum += hstart[h + 1];
      }
      int incl = sum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      int run = incl - sum;
      if (lane == 0) hstart[0] = 0;
      for (int t = 0; t < 16; ++t) {
        const int h = lane * 16 + t;
        if (h < nh) {
          run += hstart[h + 1];
          hstart[h + 1] = (unsigned short)min(run, 65535);
        }
      }
      if (lane == 31) {
        s_ntot = incl;
        if (sp == 1) atomicMax(&o.st->max_halo, incl);  // whole bricks only: sizes the next call
      }
    }
    __syncthreads();
    const int ntot = s_ntot;
    if (ntot > bp.cap_cand) {
      int kmax = 2;
      for (int k = 1; k >= 0; --k)
        if (bw[k] > bw[kmax]) kmax = k;
      if (bw[kmax] > 1) {  // cut the box in two along its longest direction
        if (tid == 0) {
          const int half = bw[kmax] / 2;
          for (int part = 0; part < 2; ++part) {
            for (int k = 0; k < 3; ++k) {
              box_stack[sp - 1 + part][k] = c0[k] + (k == kmax && part == 0 ? half : 0);
              box_stack[sp - 1 + part][3 + k] = k == kmax ? (part == 0 ? bw[k] - half : half) : bw[k];
            }
          }
        }
        sp += 1;
        __syncthreads();
        continue;
      }
      // a single cell whose stencil does not fit: per-atom gather walk straight from global memory
      int r[3] = {c0[0], c0[1], c0[2]};
      const int cell = (r[2] * g.nc[1] + r[1]) * g.nc[0] + r[0];
      const int jb = a.starts[cell], je = a.starts[cell + 1];
      if (tid == 0 && je > jb) atomicAdd(&o.st->slow_atoms, je - jb);
      for (int j = jb + tid; j < je; j += nthreads) brick_slow_atom<POT, TPA>(a, bp, o, svtab, RB_ASHIFT, j, coincident);
      if (--sp == 0) break;
      __syncthreads();
      continue;
    }
    // brick-local atom prefix over the box's own cells (x fastest)
    const float inv_bw0 = 1.0f / (float)bw[0], inv_bw1 = 1.0f / (float)bw[1];
    auto center_of = [&](int c) {
      const int cyz = (int)(((float)c + 0.5f) * inv_bw0);
      const int lz = (int)(((float)cyz + 0.5f) * inv_bw1);
      const int lx = c - cyz * bw[0], ly = cyz - lz * bw[1];
      return ((lz + bp.ns[2]) * hd[1] + (ly + bp.ns[1])) * hd[0] + lx + bp.ns[0];
    };
    if (wid == 0) {
      int cnt[2], sum = 0;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c = lane * 2 + t;
        cnt[t] = 0;
        if (c < nbc) {
          const int hc = center_of(c);
          cnt[t] = (int)hstart[hc + 1] - (int)hstart[hc];
        }
        sum += cnt[t];
      }
      int incl = sum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      int run = incl - sum;
      if (lane == 0) bc_astart[0] = 0;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c = lane * 2 + t;
        run += cnt[t];
        if (c < nbc) bc_astart[c + 1] = (unsigned short)run;
      }
    }
    RB_TICK(0);
    // brick origin for the FP32 offsets (any nearby point serves; it only has to be common)
    double org[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double fr = ((double)c0[d] + 0.5 * (double)bw[d]) / (double)g.nc[d];  // box centre
      if (a.box.periodic[d]) {
#pragma unroll
        for (int k = 0; k < 3; ++k) org[k] += fr * a.box.H[3 * d + k];
      } else {
        org[d] += a.box.lo[d] + a.box.ext[d] * fr;
      }
    }
    // ---- stage the halo: a thread copies whole halo cells ----------------------------------------------
    auto stage = [&](auto ashift_tag) {
      constexpr bool ASHIFT = decltype(ashift_tag)::value;
      for (int h = tid; h < nh; h += nthreads) {
        const int sb = hstart[h], se = hstart[h + 1], jb = hgstart[h] - sb;
        const int code = hcode[h];
        const double svx = svtab[code][0] - org[0], svy = svtab[code][1] - org[1], svz = svtab[code][2] - org[2];
        for (int s = sb; s < se; ++s) {
          const double4 p = a.spos[jb + s];
          const unsigned long long mj = (unsigned long long)__double_as_longlong(p.w);
          int cj = code;
          double sx = svx, sy = svy, sz = svz;
          if (ASHIFT) {
            const int wo = wrap_offset(mj);
            if (wo != 0) {  // vesin :1497: shift = cell shift + s_i - s_j
              cj = code - wo;
              sx = svtab[cj][0] - org[0];
              sy = svtab[cj][1] - org[1];
              sz = svtab[cj][2] - org[2];
              s_shifted = 1;
            }
          }
          c64[3 * s] = p.x;
          c64[3 * s + 1] = p.y;
          c64[3 * s + 2] = p.z;
          if (kForce) c64w[s] = o.dfd_sorted[jb + s];
          if (kForce && kFeHe) c64w2[s] = o.dfs_sorted[jb + s];
          const float fx = (float)(p.x + sx), fy = (float)(p.y + sy), fz = (float)(p.z + sz);
          const __half2 hxy = __floats2half2_rn(fx, fy), hz0 = __floats2half2_rn(fz, 0.0f);
          c16[s] = make_uint2(*reinterpret_cast<const unsigned*>(&hxy), *reinterpret_cast<const unsigned*>(&hz0));
          cinfo[s] = (unsigned short)(cj | (meta_species(mj) << 8));
        }
      }
    };
    if (RB_ASHIFT)
      stage(BoolTag<true>{});
    else
      stage(BoolTag<false>{});
    __syncthreads();
    RB_TICK(1);
    const int natoms = bc_astart[nbc];

    const double rc2 = kPairGeo ? a.lj.rc2 : a.rc2;
    const double rc2_phys = (kPairGeo || kFeHe) ? 0.0 : eam_par<kModel>().rc2_phys;
    const int apb = nthreads / TPA;  // atoms per pass
    const int nrx = 2 * bp.ns[0] + 1, nry = 2 * bp.ns[1] + 1, nrows = nry * (2 * bp.ns[2] + 1);
    // private survivor list of this thread: entry k at shared byte address lsa0 + k * lstride
    const unsigned lsa0 = (unsigned)__cvta_generic_to_shared(lists + tid), lstride = 2u * (unsigned)nthreads;
    const unsigned lsa_end = lsa0 + (unsigned)bp.cap_list * lstride;
    const float inv_lstride = 1.0f / (float)lstride;
    // SHIFTED: some halo cell of this box is a periodic image (only bricks at the box faces);
    // elsewhere (p_j - p_i) + 0 is p_j - p_i and the image-vector lookup is skipped
    auto evaluate = [&](auto shifted_tag, auto ashift_tag) {
      constexpr bool SHIFTED = decltype(shifted_tag)::value;
      constexpr bool ASHIFT = decltype(ashift_tag)::value;  // some atoms carry wrap counts (SHIFTED then)
      constexpr bool kNeedInfo = SHIFTED || POT != TILE_LJ;
      for (int a0 = 0; a0 < natoms; a0 += apb) {
        const int ai = a0 + tid / TPA;
        const int sub = tid % TPA;
        double acc[3] = {0.0, 0.0, 0.0};
        double e_thr = 0.0;  // this thread's share of the atom's energy
        int npair = 0;       // LJ: accepted pairs (the cutoff shift is applied once per atom)
        int si = 0, gi = 0;
        const bool active = ai < natoms;
        unsigned lsa = lsa0;  // end of the list of the last round
        int rounds = 0;
        if (active) {
          int c = 0;
#pragma unroll
          for (int step = 32; step > 0; step >>= 1)
            if (c + step < nbc && bc_astart[c + step] <= ai) c += step;
          const int hc = center_of(c), il = ai - bc_astart[c];
          si = hstart[hc] + il;
          gi = hgstart[hc] + il;
          const uint2 fi = c16[si];
          const __half2 fxy = *reinterpret_cast<const __half2*>(&fi.x), fz0 = *reinterpret_cast<const __half2*>(&fi.y);
          const __half thr = __float2half_ru(bp.thr2f);
          const double pix = c64[3 * si], piy = c64[3 * si + 1], piz = c64[3 * si + 2];
          const int zi = cinfo[si] >> 8;
          double dfd_i = 0.0, dfs_i = 0.0;
          if (kForce) dfd_i = c64w[si];
          if (kForce && kFeHe) dfs_i = c64w2[si];
          int wi = 0;  // own wrap counts (rare): the image vector is then worked out pair by pair
          if (ASHIFT) wi = wrap_offset((unsigned long long)__double_as_longlong(a.spos[gi].w));
          const int hlow = hc - (bp.ns[2] * hd[1] + bp.ns[1]) * hd[0] - bp.ns[0];  // stencil corner
          int ry = sub % nry, rz = sub / nry, row = sub;
          int s = 0, s1 = 0;  // cursor in the current stencil row
          // bounds of the next row, fetched one row ahead (their shared-memory latency then hides
          // behind the candidates of the current row)
          int ns0 = 0, ns1 = 0;
          auto fetch_row = [&]() {
            if (row < nrows) {
              const int h0 = hlow + (rz * hd[1] + ry) * hd[0];
              ns0 = hstart[h0];
              ns1 = hstart[h0 + nrx];
            }
            row += TPA;
            ry += TPA;
            while (ry >= nry) {
              ry -= nry;
              ++rz;
            }
          };
          fetch_row();
          bool more = true;
          // force pass: the list the density pass left for this thread, if it did
          bool have = false;
          unsigned lsa_pre = lsa0;
          if (kForce && o.pool != nullptr) {
            const unsigned long long ref = o.lref[(size_t)gi * TPA + sub];
            if (ref != kNoList) {
              const unsigned cnt = (unsigned)(ref >> 40);
              const unsigned short* src = o.pool + (ref & ((1ull << 40) - 1ull)) + lane;
#pragma unroll 4
              for (unsigned k = 0; k < cnt; ++k, lsa_pre += lstride) {
                const unsigned short v = __ldg(src + 32u * k);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(lsa_pre), "h"(v));
              }
              have = true;
            }
          }
          // rounds of search-then-physics: one round unless the private list fills up first
          do {
            // ---- search: this thread's rows of the atom's stencil, FP32 ----------------------------------
            lsa = lsa0;
            ++rounds;
            if (have) {
              lsa = lsa_pre;
              more = false;
            } else
            for (;;) {
              if (s == s1) {  // next row: 2 ns + 1 cells that are contiguous in the staged order
                if (row - TPA >= nrows) {
                  more = false;
                  break;
                }
                s = ns0;
                s1 = ns1;
                fetch_row();
                if (s == s1) continue;
              }
              int se = s1;
              if (lsa + (unsigned)(s1 - s) * lstride > lsa_end) {  // the row may not fit: count the free slots
                const int room = bp.cap_list - (int)(((float)(lsa - lsa0) + 0.5f) * inv_lstride);
                se = min(s1, s + room);
                if (se == s) break;  // list full: evaluate it, then come back
              }
#pragma unroll 4
              for (; s < se; ++s) {
                const uint2 f = c16[s];
                const __half2 dxy = __hsub2(*reinterpret_cast<const __half2*>(&f.x), fxy);
                const __half2 dz0 = __hsub2(*reinterpret_cast<const __half2*>(&f.y), fz0);
                const __half2 q = __hfma2(dz0, dz0, __hmul2(dxy, dxy));  // (dx^2 + dz^2, dy^2)
                const __half t = __hadd(__low2half(q), __high2half(q));
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(lsa), "h"((unsigned short)s));  // kept only if it passes
                lsa += __hle(t, thr) ? lstride : 0u;
              }
              if (s < s1) break;  // the row continues after the list has been evaluated
            }
            RB_TICK(2);
            // ---- physics: exact FP64 test and pair terms for the listed candidates ------------------------
#pragma unroll kPhysUnroll
            for (unsigned la = lsa0; la < lsa; la += lstride) {
              unsigned short s16;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(s16) : "r"(la));
              const int sj = s16;
              int info = 0;
              if (kNeedInfo) info = cinfo[sj];
              const double* pj = c64 + 3 * sj;
              double dx = xsub(pj[0], pix), dy = xsub(pj[1], piy), dz = xsub(pj[2], piz);
              if (SHIFTED) {
                const int code = info & 127;
                if (!ASHIFT || wi == 0) {
                  dx = xadd(dx, svtab[code][0]);
                  dy = xadd(dy, svtab[code][1]);
                  dz = xadd(dz, svtab[code][2]);
                } else {
                  const int wc = wi + kShiftCodes / 2;  // the wrap counts as a code
                  double sv[3];
                  image_vector<kPairGeo>(a, code % 5 + wc % 5 - 4, (code / 5) % 5 + (wc / 5) % 5 - 4, code / 25 + wc / 25 - 4, sv);
                  dx = xadd(dx, sv[0]);
                  dy = xadd(dy, sv[1]);
                  dz = xadd(dz, sv[2]);
                }
              }
              const double d2 = xnorm2(dx, dy, dz);
              tile_pair<POT>(a, rc2, rc2_phys, sj != si, zi, info >> 8, dfd_i, [&]() { return c64w[sj]; }, dfs_i,
                             [&]() { return c64w2[sj]; }, dx, dy, dz, d2, e_thr, acc, npair, coincident);
            }
            RB_TICK(3);
          } while (more);
        }
        __syncwarp();
        if (kDensity && o.pool != nullptr) {
          // hand the lists of this warp to the force pass: one chunk, as long as its longest list
          const unsigned cnt = (active && rounds == 1) ? (unsigned)(((float)(lsa - lsa0) + 0.5f) * inv_lstride) : 0u;
          unsigned wmax = cnt;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
          unsigned long long chunk = 0;
          if (lane == 0 && wmax) chunk = atomicAdd(&o.st->pool_used, (unsigned long long)wmax * 32ull);
          chunk = __shfl_sync(0xffffffffu, chunk, 0);
          const bool fits = chunk + (unsigned long long)wmax * 32ull <= o.pool_cap;
          if (fits) {
            unsigned short* dst = o.pool + chunk + lane;
            unsigned la = lsa0;
#pragma unroll 4
            for (unsigned k = 0; k < wmax; ++k, la += lstride) {
              unsigned short v;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(la));
              dst[32u * k] = v;
            }
          }
          if (active) o.lref[(size_t)gi * TPA + sub] = (fits && rounds == 1) ? (chunk | (unsigned long long)cnt << 40) : kNoList;
        }
        // the TPA threads of an atom fold their partial sums in a fixed tree
#pragma unroll
        for (int off = 1; off < TPA; off <<= 1) {
#pragma unroll
          for (int k = 0; k < NV; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
          if (!kDensity) e_thr += __shfl_xor_sync(0xffffffffu, e_thr, off);
          if (POT == TILE_LJ) npair += __shfl_xor_sync(0xffffffffu, npair, off);
        }
        if (active && sub == 0) {
          if (!kDensity) tile_finish<POT>(a, acc, e_thr, npair);
          if (kDensity) {
            double emb, demb;
            if (kFeHe) {
              double dems;
              fehe_embed(cinfo[si] >> 8, acc[0], acc[1], emb, demb, dems);
              o.dfs_sorted[gi] = dems;
            } else {
              eam_embed<kModel>(cinfo[si] >> 8, acc[0], emb, demb);
            }
            o.dfd_sorted[gi] = demb;
            o.eemb_sorted[gi] = emb;
          } else {
            const int orig = meta_orig((unsigned long long)__double_as_longlong(a.spos[gi].w));
            o.F[3 * (size_t)orig] = acc[0];
            o.F[3 * (size_t)orig + 1] = acc[1];
            o.F[3 * (size_t)orig + 2] = acc[2];
            if (kForce) e_thr += o.eemb_sorted[gi];
            // per-atom energy in sorted order: the fold over them has one order, whatever brick or
            // box evaluated the atom
            o.eatom[gi] = e_thr;
          }
        }
      }
    };
    if (natoms > 0) {
      if (!s_shifted)
        evaluate(BoolTag<false>{}, BoolTag<false>{});
      else if (!RB_ASHIFT)
        evaluate(BoolTag<true>{}, BoolTag<false>{});
      else
        evaluate(BoolTag<true>{}, BoolTag<true>{});
    }
    // the stack only changes when a box is cut (all threads, before any evaluation), so every
    // thread knows whether boxes remain: the last box needs no barrier and the warps retire alone
    if (--sp == 0) break;
    __syncthreads();
  }

#ifdef RB_BRICK_TIMING
  if (lane == 0)
    for (int k = 0; k < 4; ++k) atomicAdd(&o.st->prof[k], t_acc[k]);
#endif
#undef RB_ASHIFT
  if (kDensity && coincident) atomicOr(&o.st->flags, kFlagCoincident);
}

}  // namespace rb
