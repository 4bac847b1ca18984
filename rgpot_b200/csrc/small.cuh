// small.cuh -- batched all-pairs kernels: one thread block per (small) system.
//
// This is the native body of Potential<D>::forceBatchImpl (CppCore/rgpot/Potential.hpp:235-240):
// NEB / dimer / minimisation clients hand over many small geometries per call; each block keeps
// one whole system in shared memory and runs neighbor search, density, embedding and force
// passes back to back, so a 65 536-configuration batch is ONE launch.
//
//  k_lj_small   LJPot / LJClusterPot reference semantics: literal fold, one nearest image per
//               pair, r2 <= rc2.                      (vesin_visit.hpp:82-110, LJPot.cc:61-81)
//  k_eam_small  CuH2: every image with d2 < rc2 (vesin), self images dropped; per-atom neighbor
//               lists are compacted into shared memory once and reused by the density and the
//               force pass (rgpot_cuh2.f90:436-459).
#pragma once

#include "cells.cuh"
#include "physics.cuh"

namespace rb {

#ifndef RB_SMALL_UNROLL
#define RB_SMALL_UNROLL 1
#endif
constexpr int kSmallUnroll = RB_SMALL_UNROLL;  // list entries per trip of the density / force loops (tuning knob)

struct SmallBox {
  double H[9], Hinv[9];
  double q[3];  // rc / face distance (+ margin): |fractional separation| bound per direction
  int fast;     // every q < 0.5: at most one image per pair and direction
  int bad;      // singular box
};

// vesin Matrix::determinant / inverse (:100-128) and face distances (:258-275)
RB_HD double small_box_det(const double* m) {
  return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
// element k of the adjugate (the inverse is this over the determinant)
RB_HD double small_box_cofactor(const double* m, int k) {
  switch (k) {
    case 0: return m[4] * m[8] - m[7] * m[5];
    case 1: return m[2] * m[7] - m[1] * m[8];
    case 2: return m[1] * m[5] - m[2] * m[4];
    case 3: return m[5] * m[6] - m[3] * m[8];
    case 4: return m[0] * m[8] - m[2] * m[6];
    case 5: return m[3] * m[2] - m[0] * m[5];
    case 6: return m[3] * m[7] - m[6] * m[4];
    case 7: return m[6] * m[1] - m[0] * m[7];
    default: return m[0] * m[4] - m[3] * m[1];
  }
}
// rc / face distance k (+ margin); face distance k = 1 / |column k of Hinv| (equivalent to |n_k . H_k|, n_k the unit normal)
RB_HD double small_box_q(const double* v, int k, double rc) {
  const double cn = sqrt(v[k] * v[k] + v[3 + k] * v[3 + k] + v[6 + k] * v[6 + k]);
  return rc * cn * (1.0 + 1e-9) + 1e-9;
}
RB_HD void small_box_init(const double* b, double rc, SmallBox& sb) {
  for (int k = 0; k < 9; ++k) sb.H[k] = b[k];
  const double* m = sb.H;
  const double det = small_box_det(m);
  sb.bad = !(fabs(det) >= 1e-30);
  if (sb.bad) {
    sb.fast = 0;
    return;
  }
  double* v = sb.Hinv;
  for (int k = 0; k < 9; ++k) v[k] = small_box_cofactor(m, k) / det;
  sb.fast = 1;
  for (int k = 0; k < 3; ++k) {
    sb.q[k] = small_box_q(v, k, rc);
    if (!(sb.q[k] < 0.4999)) sb.fast = 0;
  }
}

struct SmallBoxF {  // FP32 mirror for the prefilter
  float h[9];
  float thr2;  // (rc + margin)^2
  int ortho;   // diagonal cell: three multiplies instead of a 3x3 product
};

// the FP32 mirror and the search threshold of a cell (host: once per single-system call; device: once per block)
RB_HD void small_boxf_init(const SmallBox& sb, double rc, SmallBoxF& sbf) {
  double hmax = 0.0;
  for (int k = 0; k < 9; ++k) {
    sbf.h[k] = (float)sb.H[k];
    hmax += fabs(sb.H[k]);
  }
  sbf.ortho = sb.H[1] == 0.0 && sb.H[2] == 0.0 && sb.H[3] == 0.0 && sb.H[5] == 0.0 && sb.H[6] == 0.0 && sb.H[7] == 0.0;
  // FP32 error of the wrapped fractional separation (~3 ulp of 1) times the cell size,
  // plus the rounding of the Cartesian norm: generous, the exact test follows anyway
  const double margin = 1.0e-6 * hmax + 1.0e-5 * rc;
  const double thr = rc + margin;
  sbf.thr2 = (float)(thr * thr * (1.0 + 1.0e-6));
}

struct SmallArgs {
  const double* pos;       // packed positions of all systems
  const int* Z;            // packed atomic numbers (may be null for LJ)
  const long long* off;    // nsys + 1 atom offsets
  const double* boxes;     // 9 per system
  double* F;               // packed forces
  double* E;               // per-system energy
  unsigned* sys_flags;     // per-system status flags
  int* sys_bad;            // per-system {first bad atom, its Z} (2 ints)
  int nsys;
  int maxnb;               // neighbor-list capacity per atom (EAM)
  int z_shared;            // homogeneous batch: Z holds the atomic numbers of ONE system, shared by all
  int single_n;            // > 0: ONE system of this many atoms starting at atom 0 (the offsets are not read:
                           //      one dependent fetch less when the inputs sit in mapped host memory)
  int split;               // k_lj_small, single system: blocks the atoms are spread over (0 = batch mapping)
  int lanes;               // ... and the lanes that share an atom there (a power of two <= 32)
  double* e_parts;         // split: per-block energy partials ...
  int* ticket;             // ... and the counter that tells the last block to fold them (left at zero)
  double* f_stage;         // split > 1: device scratch for the forces; the last block moves them to F (mapped host
                           //            memory) in one sweep, so the call pays ONE system-scope fence, not one per block
  int half;                // k_eam_small: every pair once (half lists, fixed-point scatter to the partner); see the kernel
  int box_ready;           // k_eam_small, single system: the cell has been prepared by the host (sb, sbf below)
  SmallBox sb;
  SmallBoxF sbf;
  int cluster;             // LJ: free boundaries
  double u0, psi2, shiftU, cutoff;  // LJ
  int morse;                        // 1 Morse, 2 ZBL instead of LJ
  double m_de, m_a, m_re, m_ecut;
  const double* zbl_table;          // ZBL coefficient table (device)
  int zbl_nt;
  double zbl_rin;
  EmitArgs emit;           // optional (single system only)
};

// ------------------------------------------------------------------------------------------ LJ
static __global__ void __launch_bounds__(256)
k_lj_small(SmallArgs a) {
  extern __shared__ double smem[];
  __shared__ double sh32[32];
  const bool split = a.split > 0;
  const int sys = split ? 0 : blockIdx.x;
  const long long base = a.single_n > 0 ? 0 : a.off[sys];
  const int n = a.single_n > 0 ? a.single_n : (int)(a.off[sys + 1] - base);
  RB_STAMP(0);
  double* px = smem;  // 3n interleaved
  int* tp = reinterpret_cast<int*>(smem + 3 * n);  // ZBL: type index per atom
  if (split && ((reinterpret_cast<uintptr_t>(a.pos) & 15) == 0)) {  // (every block of a spread system reads all of it)
    const double2* src = reinterpret_cast<const double2*>(a.pos);
    double2* dst = reinterpret_cast<double2*>(px);
    const int n2 = (3 * n) >> 1;
    for (int t = threadIdx.x; t < n2; t += blockDim.x) dst[t] = src[t];
    if (((3 * n) & 1) && threadIdx.x == 0) px[3 * n - 1] = a.pos[3 * n - 1];
  } else {
    for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) px[t] = a.pos[3 * base + t];
  }
  if (a.morse == 2)
    for (int t = threadIdx.x; t < n; t += blockDim.x) tp[t] = a.Z[(a.z_shared ? 0 : base) + t];

  // fold parameters: PairListCache.hpp:176-184 ; LJClusterPot.cc:46-55 (no fold at all)
  LJParams p;
  p.u0 = a.u0;
  p.psi2 = a.psi2;
  p.shiftU = a.shiftU;
  p.rc2 = xmul(a.cutoff, a.cutoff);
  p.morse = a.morse;
  p.m_de = a.m_de;
  p.m_a = a.m_a;
  p.m_re = a.m_re;
  p.m_ecut = a.m_ecut;
  p.m_two_de_a = 2.0 * a.m_de * a.m_a;
  p.zbl_table = a.zbl_table;
  p.zbl_nt = a.zbl_nt;
  p.zbl_rin = a.zbl_rin;
  p.zbl_rin2 = a.zbl_rin * a.zbl_rin;
  const double* box = a.boxes + 9 * (size_t)sys;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p.w[k] = box[4 * k];
    p.inv[k] = (!a.cluster && p.w[k] != 0.0) ? 1.0 / p.w[k] : 0.0;
  }
  __syncthreads();
  RB_STAMP(1);

  double e = 0.0;
  if (!split) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double xi = px[3 * i], yi = px[3 * i + 1], zi = px[3 * i + 2];
      const int ti = a.morse == 2 ? tp[i] : 0;
      double fx = 0.0, fy = 0.0, fz = 0.0;
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const double dx = fold_component(xsub(px[3 * j], xi), p.w[0], p.inv[0]);
        const double dy = fold_component(xsub(px[3 * j + 1], yi), p.w[1], p.inv[1]);
        const double dz = fold_component(xsub(px[3 * j + 2], zi), p.w[2], p.inv[2]);
        const double r2 = xnorm2(dx, dy, dz);
        if (!(r2 <= p.rc2)) continue;
        lj_pair(p, ti, a.morse == 2 ? tp[j] : 0, dx, dy, dz, r2, e, fx, fy, fz);
        if (a.emit.pairs != nullptr && i < j) {
          const unsigned long long k = atomicAdd(a.emit.count, 1ull);
          if ((long)k < a.emit.cap) {
            a.emit.pairs[2 * k] = i;
            a.emit.pairs[2 * k + 1] = j;
            a.emit.shifts[3 * k] = a.emit.shifts[3 * k + 1] = a.emit.shifts[3 * k + 2] = 0;
          }
        }
      }
      a.F[3 * (base + i)] = fx;
      a.F[3 * (base + i) + 1] = fy;
      a.F[3 * (base + i) + 2] = fz;
    }
  } else {
    // ONE system spread over the lanes of the grid (a single 500-atom system in one block would leave 147
    // SMs idle, and thirteen atoms in thirteen threads a dependent chain of twelve pairs each): a.lanes
    // lanes share an atom and take every a.lanes-th partner, their partial sums are folded by shuffles
    // (a fixed tree), block b owns atoms [b T / lanes, (b + 1) T / lanes)
    const int lanes = a.lanes, apb = blockDim.x / lanes, sub = threadIdx.x & (lanes - 1);
    const int i = blockIdx.x * apb + (int)threadIdx.x / lanes;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (i < n) {
      const double xi = px[3 * i], yi = px[3 * i + 1], zi = px[3 * i + 2];
      const int ti = a.morse == 2 ? tp[i] : 0;
      for (int j = sub; j < n; j += lanes) {
        if (j == i) continue;
        const double dx = fold_component(xsub(px[3 * j], xi), p.w[0], p.inv[0]);
        const double dy = fold_component(xsub(px[3 * j + 1], yi), p.w[1], p.inv[1]);
        const double dz = fold_component(xsub(px[3 * j + 2], zi), p.w[2], p.inv[2]);
        const double r2 = xnorm2(dx, dy, dz);
        if (!(r2 <= p.rc2)) continue;
        lj_pair(p, ti, a.morse == 2 ? tp[j] : 0, dx, dy, dz, r2, e, fx, fy, fz);
      }
    }
    for (int off = 1; off < lanes; off <<= 1) {
      fx += __shfl_xor_sync(0xffffffffu, fx, off);
      fy += __shfl_xor_sync(0xffffffffu, fy, off);
      fz += __shfl_xor_sync(0xffffffffu, fz, off);
    }
    if (i < n && sub == 0) {
      double* Fo = gridDim.x > 1 ? a.f_stage : a.F;
      Fo[3 * i] = fx;
      Fo[3 * i + 1] = fy;
      Fo[3 * i + 2] = fz;
    }
  }
  RB_STAMP(2);
  const double bs = block_sum(e, sh32);  // (its barriers also put every thread's force stores before thread 0's fence)
  RB_STAMP(3);
  if (!split || gridDim.x == 1) {
    if (threadIdx.x == 0) {
      a.E[sys] = bs;
      if (a.single_n > 0) __threadfence_system();  // (a host that polls the flag word must find the results behind it)
      RB_STAMP(4);
      a.sys_flags[sys] = 0u;   // every exit of the batch kernels writes the flag word: no reset copy needed
    }
    return;
  }
  // split: the block that finishes last adds the block sums in block order, moves the forces to the host and signs off
  __shared__ int last;
  if (threadIdx.x == 0) {
    a.e_parts[blockIdx.x] = bs;
    __threadfence();  // this block's staged forces (ordered before this thread by the barrier) before its ticket
    last = atomicAdd(a.ticket, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) a.F[t] = __ldcg(a.f_stage + t);
  double tot = 0.0;
  if (threadIdx.x < 32) {
    // block order, one fixed chain; the loads go out together
    for (unsigned b0 = 0; b0 < gridDim.x; b0 += 32) {
      const unsigned b = b0 + threadIdx.x;
      const double v = b < gridDim.x ? __ldcg(a.e_parts + b) : 0.0;
      for (int l = 0; l < 32; ++l) tot += __shfl_sync(0xffffffffu, v, l);
    }
  }
  __syncthreads();  // every thread's force stores before thread 0's fence
  if (threadIdx.x == 0) {
    a.E[0] = tot;
    *a.ticket = 0;
    __threadfence_system();
    a.sys_flags[0] = 0u;
  }
}

// ---------------------------------------------------------------------------------------- CuH2
// list entry: j (16 bits) | (S0+16) << 16 | (S1+16) << 21 | (S2+16) << 26
RB_D unsigned pack_entry(int j, int s0, int s1, int s2) {
  return (unsigned)j | ((unsigned)(s0 + 16) << 16) | ((unsigned)(s1 + 16) << 21) |
         ((unsigned)(s2 + 16) << 26);
}

// exact separation of entry (j, S) seen from i: (p_j - p_i) + S.H ; vesin :1199, :1972
RB_D double entry_vec(const SmallBox& sb, const double* px, const double* py, const double* pz,
                      int i, int j, int s0, int s1, int s2, double& vx, double& vy, double& vz) {
  vx = xsub(px[j], px[i]);
  vy = xsub(py[j], py[i]);
  vz = xsub(pz[j], pz[i]);
  if ((s0 | s1 | s2) != 0) {  // adding the zero shift would not change a bit
    vx = xadd(vx, shift_component(sb.H, 0, (double)s0, (double)s1, (double)s2));
    vy = xadd(vy, shift_component(sb.H, 1, (double)s0, (double)s1, (double)s2));
    vz = xadd(vz, shift_component(sb.H, 2, (double)s0, (double)s1, (double)s2));
  }
  return xnorm2(vx, vy, vz);
}

// --------------------------------------------------------------------------------------------
// k_eam_small: one block = one system.
//
//  phase A  load, classify (rgpot_cuh2.f90:463-489), wrap into the cell (fractional coordinates,
//           FP64 + an FP32 copy for the prefilter); table of the 27 nearest image vectors S.H
//  phase B  neighbor search, TPA threads per atom (the mapping of the physics phases): a thread
//           tests its share of the other atoms -- FP32, nearest image in fractional space,
//           Cartesian distance with a conservative margin; every lane of a warp reads the same
//           candidate (a shared-memory broadcast) and runs the same trip count -- and appends the
//           survivors to its PRIVATE 16-bit list with a predicated store: no ballots, no queues.
//           A second sweep over the (short) list adds the image index and the species.  The
//           exact FP64 test of vesin ((p_j - p_i) + S.H, d2 < rc2) is the one the physics phases
//           evaluate anyway, so accept / reject stays bit-exact without a separate pass.
//           Boxes narrower than two cutoffs (several images per pair) take a serial enumeration
//           by lane 0 instead.
//  phase C  atoms ordered by neighbor count so the lanes of a warp run equally long loops
//  phase D  density + embedding (rgpot_cuh2.f90:436-451), TPA lanes per atom over its list
//  phase E  pair energy + forces (rgpot_cuh2.f90:454-457), same mapping
// Every sum has a fixed order: results are bit-reproducible.
//
// list entry (16 bits): j (9 bits, n <= 512) | image index (5 bits, 0..26 = (S0+1) + 3 (S1+1)
//             + 9 (S2+1)) << 9 | species of j << 14.  Entry k of thread t = i * TPA + sub sits at
//             list[k * n * TPA + t].  Images further than one cell away (atoms far outside the
//             box) or a full list raise kFlagListOverflow and the system reruns on the cell-list
//             kernels.
// --------------------------------------------------------------------------------------------


// 64-bit fixed-point accumulation in shared memory.  A 64-bit shared-memory atomic add compiles to a
// compare-and-swap loop; two native 32-bit adds with the carry taken from the value the first one returns give
// the same sum (every wrap of the low word is seen by exactly one caller) without a loop.
RB_D void fixed_add(long long* acc, long long q) {
  unsigned* p = reinterpret_cast<unsigned*>(acc);
  const unsigned lo = (unsigned)q, hi = (unsigned)((unsigned long long)q >> 32);
  const unsigned old = atomicAdd(p, lo);
  atomicAdd(p + 1, hi + ((old + lo) < lo ? 1u : 0u));
}

// three accumulators at once (a force vector): the three returning adds go out back to back, so their latencies
// overlap instead of each carry waiting for its own
RB_D void fixed_add3(long long* acc, long long qx, long long qy, long long qz) {
  unsigned* p = reinterpret_cast<unsigned*>(acc);
  const unsigned lx = (unsigned)qx, ly = (unsigned)qy, lz = (unsigned)qz;
  const unsigned ox = atomicAdd(p, lx);
  const unsigned oy = atomicAdd(p + 2, ly);
  const unsigned oz = atomicAdd(p + 4, lz);
  atomicAdd(p + 1, (unsigned)((unsigned long long)qx >> 32) + ((ox + lx) < lx ? 1u : 0u));
  atomicAdd(p + 3, (unsigned)((unsigned long long)qy >> 32) + ((oy + ly) < ly ? 1u : 0u));
  atomicAdd(p + 5, (unsigned)((unsigned long long)qz >> 32) + ((oz + lz) < lz ? 1u : 0u));
}

template <int TPA, int M>
__global__ void __launch_bounds__(448, 2)
k_eam_small(SmallArgs a) {
  constexpr bool kFeHe = M == EAM_FEHE;  // two density channels, branches on r = sqrt(d2) (rgpot_fehe.f90)
  // (compile-time selections: for the two analytic models every bound below stays the constant-memory
  // operand it always was)
  const double model_rc = kFeHe ? 5.4 : eam_par<kFeHe ? EAM_CUH2 : M>().rcut;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sh32[32];
  __shared__ SmallBox sb;
  __shared__ SmallBoxF sbf;
  __shared__ unsigned blk_flags;
  __shared__ int blk_bad[2];
  __shared__ int hist[256];
  // half-list mode: pairs with a hydrogen partner are set aside and evaluated afterwards by consecutive threads
  // (inside the list walk a single such entry drags its whole warp through the mixed-species code at one or two
  // active lanes); their energies go through a fixed-point accumulator, so nothing depends on the queue order
  constexpr int kDeferCap = 512;
  __shared__ unsigned dq[kDeferCap];
  __shared__ int dq_n;
  __shared__ long long e_acc;
  // half-list mode: every thread's list is split in two (even / odd entries); the 2 NTL half lists are sorted by
  // length and a thread walks a long one and a short one, so that warps (and the lanes inside them) finish together
  // -- the owned pairs per atom range from 8 to 38 in the 218-atom slab, and a barrier waits for the longest list
  __shared__ unsigned short perm2[1024];  // half lists by decreasing length (half list v = 2 * column + parity)
  __shared__ unsigned short vcnt[1024];   // entries of every half list that passed the exact test (density pass)

  const int sys = blockIdx.x;
  const long long base = a.single_n > 0 ? 0 : a.off[sys];
  const long long zbase = a.z_shared ? 0 : base;
  const int n = a.single_n > 0 ? a.single_n : (int)(a.off[sys + 1] - base);
  const int maxnb = a.maxnb;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarp = blockDim.x >> 5;
  const int NTL = n * TPA;                       // search / physics threads of this system
  const int lcap = (maxnb + TPA - 1) / TPA + 4;  // list entries per thread

  // shared layout
  double* pos3 = reinterpret_cast<double*>(smem_raw);  // x, y, z interleaved
  double* g3 = pos3 + 3 * n;                            // wrapped fractional coordinates
  double* dfd = g3 + 3 * n;
  double* svt = dfd + n;                                // 27 image vectors, 3 doubles each
  long long* facc = reinterpret_cast<long long*>(svt + 81);  // half-list mode: fixed-point force accumulators, 3 per atom
  size_t off_bytes = sizeof(double) * (10 * (size_t)n + 81);
  off_bytes = (off_bytes + 15) & ~(size_t)15;
  float4* g32 = reinterpret_cast<float4*>(smem_raw + off_bytes);  // FP32 copy of the wrapped fractional coordinates
  int* ash = reinterpret_cast<int*>(g32 + n);      // packed wrap counts (10 bits each, biased 512)
  int* cnt = ash + n;                              // listed candidates per atom
  int* spc = cnt + n;                              // species
  int* perm = spc + n;                             // atoms by decreasing neighbor count
  unsigned short* cnt16 = reinterpret_cast<unsigned short*>(perm + n);  // listed candidates per thread
  unsigned short* list = cnt16 + ((NTL + 1) & ~1);                      // lcap * NTL entries

  RB_STAMP(0);
  {  // the table of the exponential (exp_table_init without its barrier: the one after the cell follows)
    double* tab = exp_table();
    for (int t = threadIdx.x; t < 64; t += blockDim.x) tab[t] = g_exp2tab[t];
  }
  RB_STAMP(1);
  // this thread's first atom: the loads go out before the cell is inverted (the barriers below would otherwise
  // sit between the two global-memory latencies)
  const bool early = (int)threadIdx.x < n;
  double ex = 0.0, ey = 0.0, ez = 0.0;
  if (early) {
    ex = a.pos[3 * (base + threadIdx.x)];
    ey = a.pos[3 * (base + threadIdx.x) + 1];
    ez = a.pos[3 * (base + threadIdx.x) + 2];
  }
  if (a.box_ready) {
    if (threadIdx.x == 0) sb = a.sb;  // one system: the host has inverted the cell
  } else {
    // nine threads, one element of the inverse each (a thread alone spends ~3 us on the nine FP64 divisions and three
    // square roots while the rest of the block waits), then three for the face distances
    if (threadIdx.x < 9) {
      const double* b = a.boxes + 9 * (size_t)sys;
      double m[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) m[k] = b[k];
      const double det = small_box_det(m);
      const bool bad = !(fabs(det) >= 1e-30);
      sb.H[threadIdx.x] = m[threadIdx.x];
      sb.Hinv[threadIdx.x] = bad ? 0.0 : small_box_cofactor(m, (int)threadIdx.x) / det;
      if (threadIdx.x == 0) sb.bad = bad;
    }
    __syncthreads();
    if (threadIdx.x < 3) sb.q[threadIdx.x] = sb.bad ? 1.0 : small_box_q(sb.Hinv, (int)threadIdx.x, model_rc);
    __syncthreads();
    if (threadIdx.x == 0) sb.fast = !sb.bad && sb.q[0] < 0.4999 && sb.q[1] < 0.4999 && sb.q[2] < 0.4999;
  }
  if (threadIdx.x == 0) {
    blk_flags = sb.bad ? kFlagSingularBox : 0u;  // vesin: "the box matrix is not invertible"
    blk_bad[0] = INT_MAX;
    blk_bad[1] = 0;
    dq_n = 0;
    e_acc = 0;
    if (!sb.bad) {
      if (a.box_ready)
        sbf = a.sbf;
      else
        small_boxf_init(sb, model_rc, sbf);
    }
  }
  __syncthreads();
  // Half-list mode (one image per pair only): atom i owns the partners i + 1 ... i + (n - 1) / 2 (indices mod n; for
  // even n the pair at distance n / 2 belongs to the smaller index), evaluates each of its pairs ONCE and hands the
  // partner its share through 64-bit FIXED-POINT atomics in shared memory -- integer addition is associative, so
  // the sums do not depend on the order the lanes arrive in and results stay bit-reproducible.  Scales: density
  // 2^-53 (terms < 0.25), forces 2^-40 eV/A (terms < 1024): the rounding, 1e-16 and 9e-13 absolute, is far inside
  // the 1e-10 / 1e-9 budget; a term outside those ranges (atoms almost on top of each other) sends the system to
  // the general path (kFlagListOverflow) like a full list does.
  const bool half = !kFeHe && a.half != 0 && sb.fast != 0 && !sb.bad && a.emit.pairs == nullptr &&
                    model_rc < 8.0;  // (the force range check below assumes |vec| < 8)
  long long* rho_acc = reinterpret_cast<long long*>(g3 + 2 * n);
  long long* e_fix = reinterpret_cast<long long*>(g3 + n);  // pair energies per atom, fixed point (g3[n, 2n): free outside Fe-He)
  if (threadIdx.x < 27 && !sb.bad) {
    const int t = (int)threadIdx.x;
    const double s0 = (double)(t % 3 - 1), s1 = (double)((t / 3) % 3 - 1), s2 = (double)(t / 9 - 1);
    svt[3 * threadIdx.x] = shift_component(sb.H, 0, s0, s1, s2);
    svt[3 * threadIdx.x + 1] = shift_component(sb.H, 1, s0, s1, s2);
    svt[3 * threadIdx.x + 2] = shift_component(sb.H, 2, s0, s1, s2);
  }

  RB_STAMP(2);
  // ---- phase A
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const bool first = i == (int)threadIdx.x;  // (loaded above)
    const double x = first ? ex : a.pos[3 * (base + i)], y = first ? ey : a.pos[3 * (base + i) + 1],
                 z = first ? ez : a.pos[3 * (base + i) + 2];
    pos3[3 * i] = x;
    pos3[3 * i + 1] = y;
    pos3[3 * i + 2] = z;
    cnt[i] = 0;
    const int zn = M == EAM_AL ? 29 : a.Z[zbase + i];  // EAM-Al is single-species: atomic numbers are not read
    int sp = 0;
    if (zn == (kFeHe ? 26 : 29)) {  // rgpot_cuh2.f90:463-489 / rgpot_fehe.f90:541-569
      sp = 0;
    } else if (zn == (kFeHe ? 2 : 1)) {
      sp = 1;
    } else {
      atomicOr(&blk_flags, kFlagForeignZ);
      atomicMin(&blk_bad[0], i);
    }
    spc[i] = sp;
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) {
      atomicOr(&blk_flags, kFlagNonFinite);
      atomicMin(&blk_bad[0], i);
      g3[3 * i] = g3[3 * i + 1] = g3[3 * i + 2] = 0.0;
      g32[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      ash[i] = 512 | (512 << 10) | (512 << 20);
      continue;
    }
    double f[3];
    int w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      f[k] = x * sb.Hinv[k] + y * sb.Hinv[3 + k] + z * sb.Hinv[6 + k];
      double fl = floor(f[k]);
      fl = fmin(fmax(fl, -1.0e9), 1.0e9);
      w[k] = (int)fl;
      f[k] -= fl;
    }
    g3[3 * i] = f[0];
    g3[3 * i + 1] = f[1];
    g3[3 * i + 2] = f[2];
    g32[i] = make_float4((float)f[0], (float)f[1], (float)f[2], 0.f);
    if (abs(w[0]) > 500 || abs(w[1]) > 500 || abs(w[2]) > 500) atomicOr(&blk_flags, kFlagShiftRange);
    if ((w[0] | w[1] | w[2]) != 0) atomicOr(&blk_flags, kFlagAtomShift);
    ash[i] = ((w[0] + 512) & 1023) | (((w[1] + 512) & 1023) << 10) | (((w[2] + 512) & 1023) << 20);
  }
  __syncthreads();
  if (blk_flags & (kFlagForeignZ | kFlagNonFinite | kFlagShiftRange | kFlagSingularBox)) {
    // the reference leaves energy and forces zeroed on failure
    for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) a.F[3 * base + t] = 0.0;
    __syncthreads();  // (every thread's stores before thread 0's fence)
    if (threadIdx.x == 0) {
      a.E[sys] = 0.0;
      a.sys_bad[2 * sys] = blk_bad[0];
      a.sys_bad[2 * sys + 1] =
          ((blk_flags & kFlagForeignZ) && blk_bad[0] != INT_MAX) ? a.Z[zbase + blk_bad[0]] : 0;
      if (a.single_n > 0) __threadfence_system();
      a.sys_flags[sys] = blk_flags;
    }
    return;
  }

  RB_STAMP(3);
  const double rc2 = kFeHe ? xmul(5.4, 5.4) : eam_par<kFeHe ? EAM_CUH2 : M>().rc2;
  const bool wrapped_atoms = (blk_flags & kFlagAtomShift) != 0;

  // ---- phase B
  if (sb.fast) {
    const float h0 = sbf.h[0], h1 = sbf.h[1], h2 = sbf.h[2], h3 = sbf.h[3], h4 = sbf.h[4],
                h5 = sbf.h[5], h6 = sbf.h[6], h7 = sbf.h[7], h8 = sbf.h[8];
    const float thr2 = sbf.thr2;
    const bool ortho = sbf.ortho != 0;
    const unsigned lbase = (unsigned)__cvta_generic_to_shared(list), lstride = 2u * (unsigned)NTL;
    if (half) {  // (the FP64 fractional coordinates in g3 are not read again in this mode)
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        rho_acc[t] = 0;
        e_fix[t] = 0;
      }
      for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) facc[t] = 0;
    }
    for (int t = threadIdx.x; t < NTL; t += blockDim.x) {
      const int i = t / TPA, sub = t - i * TPA;
      const float4 gi = g32[i];
      const unsigned la0 = lbase + 2u * (unsigned)t, la_last = la0 + (unsigned)(lcap - 1) * lstride;
      unsigned la = la0;  // list cursor (shared byte address); the last slot doubles as the overflow sink
      auto search_half = [&](auto ortho_tag) {
        constexpr bool ORTHO = decltype(ortho_tag)::value;
        const int dmax = ((n - 1) >> 1) + (((n & 1) == 0 && i < (n >> 1)) ? 1 : 0);
#pragma unroll 4
        for (int d = 1 + sub; d <= dmax; d += TPA) {
          int j = i + d;
          j -= j >= n ? n : 0;
          const float4 c = g32[j];
          float t0 = c.x - gi.x, t1 = c.y - gi.y, t2 = c.z - gi.z;
          t0 -= rintf(t0);
          t1 -= rintf(t1);
          t2 -= rintf(t2);
          float dx, dy, dz;
          if (ORTHO) {
            dx = t0 * h0;
            dy = t1 * h4;
            dz = t2 * h8;
          } else {
            dx = fmaf(t0, h0, fmaf(t1, h3, t2 * h6));
            dy = fmaf(t0, h1, fmaf(t1, h4, t2 * h7));
            dz = fmaf(t0, h2, fmaf(t1, h5, t2 * h8));
          }
          const bool pass = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= thr2;
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(la), "h"((unsigned short)j));  // kept only if it passes
          la = min(la + (pass ? lstride : 0u), la_last);
        }
      };
      auto search = [&](auto ortho_tag) {
        constexpr bool ORTHO = decltype(ortho_tag)::value;
#pragma unroll 4
        for (int j = sub; j < n; j += TPA) {
          const float4 c = g32[j];  // the same address in every lane of an atom group: a broadcast
          float t0 = c.x - gi.x, t1 = c.y - gi.y, t2 = c.z - gi.z;
          t0 -= rintf(t0);
          t1 -= rintf(t1);
          t2 -= rintf(t2);
          float dx, dy, dz;
          if (ORTHO) {
            dx = t0 * h0;
            dy = t1 * h4;
            dz = t2 * h8;
          } else {
            dx = fmaf(t0, h0, fmaf(t1, h3, t2 * h6));
            dy = fmaf(t0, h1, fmaf(t1, h4, t2 * h7));
            dz = fmaf(t0, h2, fmaf(t1, h5, t2 * h8));
          }
          const bool pass = (fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= thr2) && (j != i);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(la), "h"((unsigned short)j));  // kept only if it passes
          la = min(la + (pass ? lstride : 0u), la_last);
        }
      };
      if (half) {
        if (ortho)
          search_half(BoolTag<true>{});
        else
          search_half(BoolTag<false>{});
      } else if (ortho) {
        search(BoolTag<true>{});
      } else {
        search(BoolTag<false>{});
      }
      if (la == la_last) atomicOr(&blk_flags, kFlagListOverflow);
      const int c = (int)(((float)(la - la0) + 0.5f) / (float)lstride);
      // second sweep: image index S' = -rint(g_j - g_i) (the search's choice, recomputed the same
      // way) plus the atoms' own wrap counts, and the species of j
      // (half-list mode: the density pass does this on the fly -- its lists are balanced, these are not)
      const int ai = ash[i];
      bool far_image = false;
      unsigned short* row = list + t;
      for (int k = 0; k < (half ? 0 : c); ++k) {
        const int j = row[k * NTL];
        const float4 gj = g32[j];
        int S0 = -(int)rintf(gj.x - gi.x), S1 = -(int)rintf(gj.y - gi.y), S2 = -(int)rintf(gj.z - gi.z);
        if (wrapped_atoms) {
          const int aj = ash[j];
          S0 += (ai & 1023) - (aj & 1023);
          S1 += ((ai >> 10) & 1023) - ((aj >> 10) & 1023);
          S2 += ((ai >> 20) & 1023) - ((aj >> 20) & 1023);
          far_image |= abs(S0) > 1 || abs(S1) > 1 || abs(S2) > 1;
        }
        const int tidx = (S0 + 1) + 3 * (S1 + 1) + 9 * (S2 + 1);
        row[k * NTL] = (unsigned short)(j | ((tidx & 31) << 9) | (spc[j] << 14));
      }
      if (far_image) atomicOr(&blk_flags, kFlagListOverflow);
      cnt16[t] = (unsigned short)c;
    }
  } else if (lane == 0) {
    // small boxes: several images per pair; one lane enumerates them in (j, image) order
    for (int i = warp; i < n; i += nwarp) {
      const int ai = ash[i];
      int count = 0;
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;  // self images are dropped (rgpot_neighbors.f90:132)
        const double t0 = g3[3 * j] - g3[3 * i], t1 = g3[3 * j + 1] - g3[3 * i + 1],
                     t2 = g3[3 * j + 2] - g3[3 * i + 2];
        const int aj = ash[j];
        const int w0 = (ai & 1023) - (aj & 1023), w1 = ((ai >> 10) & 1023) - ((aj >> 10) & 1023),
                  w2 = ((ai >> 20) & 1023) - ((aj >> 20) & 1023);
        const int lo0 = (int)ceil(-t0 - sb.q[0]), hi0 = (int)floor(-t0 + sb.q[0]);
        const int lo1 = (int)ceil(-t1 - sb.q[1]), hi1 = (int)floor(-t1 + sb.q[1]);
        const int lo2 = (int)ceil(-t2 - sb.q[2]), hi2 = (int)floor(-t2 + sb.q[2]);
        for (int s0 = lo0; s0 <= hi0; ++s0)
          for (int s1 = lo1; s1 <= hi1; ++s1)
            for (int s2 = lo2; s2 <= hi2; ++s2) {
              const int S0 = s0 + w0, S1 = s1 + w1, S2 = s2 + w2;
              double vx, vy, vz;
              vx = xadd(xsub(pos3[3 * j], pos3[3 * i]), shift_component(sb.H, 0, (double)S0, (double)S1, (double)S2));
              vy = xadd(xsub(pos3[3 * j + 1], pos3[3 * i + 1]), shift_component(sb.H, 1, (double)S0, (double)S1, (double)S2));
              vz = xadd(xsub(pos3[3 * j + 2], pos3[3 * i + 2]), shift_component(sb.H, 2, (double)S0, (double)S1, (double)S2));
              if (xnorm2(vx, vy, vz) < rc2) {
                if (count / TPA < lcap && abs(S0) <= 1 && abs(S1) <= 1 && abs(S2) <= 1) {
                  const int tidx = (S0 + 1) + 3 * (S1 + 1) + 9 * (S2 + 1);
                  list[(count / TPA) * NTL + i * TPA + count % TPA] =
                      (unsigned short)(j | (tidx << 9) | (spc[j] << 14));
                } else {
                  atomicOr(&blk_flags, kFlagListOverflow);
                }
                ++count;
              }
            }
      }
      for (int sub = 0; sub < TPA; ++sub) cnt16[i * TPA + sub] = (unsigned short)((count + TPA - 1 - sub) / TPA);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {  // candidates per atom: balances the lanes below
    int c = 0;
#pragma unroll
    for (int sub = 0; sub < TPA; ++sub) c += cnt16[i * TPA + sub];
    cnt[i] = c;
  }
  __syncthreads();
  if (blk_flags & kFlagListOverflow) {
    if (threadIdx.x == 0) a.sys_flags[sys] = kFlagListOverflow;  // host reruns on the cell path
    return;
  }

  // optional emission of the table (rgpot_b200_neighbors)
  if (a.emit.pairs != nullptr) {
    for (int t = threadIdx.x; t < NTL; t += blockDim.x) {
      const int i = t / TPA;
      for (int s = 0; s < cnt16[t]; ++s) {
        const unsigned en = list[s * NTL + t];
        const int j = (int)(en & 511u), tidx = (int)((en >> 9) & 31u);
        const double* pi = pos3 + 3 * i;
        const double* pj = pos3 + 3 * j;
        const double* sv = svt + 3 * tidx;
        const double vx = xadd(xsub(pj[0], pi[0]), sv[0]);
        const double vy = xadd(xsub(pj[1], pi[1]), sv[1]);
        const double vz = xadd(xsub(pj[2], pi[2]), sv[2]);
        if (!(xnorm2(vx, vy, vz) < rc2)) continue;  // the exact test of vesin
        const unsigned long long k = atomicAdd(a.emit.count, 1ull);
        if ((long)k < a.emit.cap) {
          a.emit.pairs[2 * k] = i;
          a.emit.pairs[2 * k + 1] = j;
          a.emit.shifts[3 * k] = tidx % 3 - 1;
          a.emit.shifts[3 * k + 1] = (tidx / 3) % 3 - 1;
          a.emit.shifts[3 * k + 2] = tidx / 9 - 1;
        }
      }
    }
  }

  RB_STAMP(4);
  // ---- phase C: order atoms by decreasing neighbor count (counting sort).  The order only decides
  // which lanes work on which atom; every per-atom result, and the energy sum below (taken in atom
  // order), is independent of it, so the integer atomics do not disturb reproducibility.
  const int NL = 2 * NTL;  // half lists
  if (half) {
    // half lists by decreasing length (counting sort; ties in arrival order: every result below is an integer sum
    // or a per-list sum in list order, so the order of equal lengths changes no bit)
    for (int b = threadIdx.x; b < 64; b += blockDim.x) hist[b] = 0;  // (64 bins: a half list holds a dozen entries; longer ones share the last bin)
    __syncthreads();
    int my_rank[3] = {0, 0, 0};  // (NL <= 1024, >= 352 threads whenever NL > 704)
    {
      int k = 0;
      for (int v = threadIdx.x; v < NL; v += blockDim.x, ++k) {
        const int c = ((int)cnt16[v >> 1] + 1 - (v & 1)) >> 1;
        vcnt[v] = (unsigned short)c;
        my_rank[k] = atomicAdd(&hist[min(c, 63)], 1);
      }
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan over the bins, largest count first
      int carry = 0;
      for (int base = 63; base >= 0; base -= 32) {
        const int b = base - lane;
        const int v = hist[b];
        int incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += t;
        }
        hist[b] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncthreads();
    {
      int k = 0;
      for (int v = threadIdx.x; v < NL; v += blockDim.x, ++k) perm2[hist[min((int)vcnt[v], 63)] + my_rank[k]] = (unsigned short)v;
    }
    __syncthreads();
  } else if (blockDim.x == 32) {
    // one warp: the order cannot change which lanes wait for which (a handful of atoms, a single small system)
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = i;
    __syncthreads();
  } else {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    int my_rank[2] = {0, 0};  // blockDim >= n / 2 for every launch configuration (n <= 511, >= 256 threads)
    {
      int k = 0;
      for (int i = threadIdx.x; i < n; i += blockDim.x, ++k) my_rank[k & 1] = atomicAdd(&hist[min(cnt[i], 255)], 1);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan over the bins, largest count first
      int carry = 0;
      for (int base = 255; base >= 0; base -= 32) {
        const int b = base - lane;
        const int v = hist[b];
        int incl = v;
  #pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += t;
        }
        hist[b] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncthreads();
    {
      int k = 0;
      for (int i = threadIdx.x; i < n; i += blockDim.x, ++k) perm[hist[min(cnt[i], 255)] + my_rank[k & 1]] = i;
    }
    __syncthreads();
  }
  double* eatom = g3;  // per-atom energies (the fractional coordinates are no longer needed)
  double* dfs = g3 + n;  // Fe-He: F'(rho) of the s-band channel

  const int sub = threadIdx.x % TPA;
  const int grp = threadIdx.x / TPA;
  const int ngrp = blockDim.x / TPA;
  const double rc2_phys = kFeHe ? rc2 : eam_par<kFeHe ? EAM_CUH2 : M>().rc2_phys;

  double e = 0.0;
  RB_STAMP(5);
  // ---- phase D: density, embedding
  constexpr int MM = kFeHe ? EAM_CUH2 : M;  // (the half-list branches are never taken for Fe-He)
  const double kFixMagic = 6755399441055744.0;  // 1.5 * 2^52: fma(v, scale, magic) holds rint(v * scale) in its low bits
  if (half) {
    for (int u = threadIdx.x; u < NTL; u += blockDim.x) {
      bool coincident = false, range = false;
#pragma unroll 1
      for (int hl = 0; hl < 2; ++hl) {  // a long half list, then a short one
        const int v = perm2[hl ? NL - 1 - u : u];
        const int col = v >> 1, i = col / TPA, c = vcnt[v];
        const int zi = spc[i];
        const double xi = pos3[3 * i], yi = pos3[3 * i + 1], zi_ = pos3[3 * i + 2];
        const float4 gi = g32[i];
        const int ai = ash[i];
        unsigned short* row = list + col + (v & 1) * NTL;  // entries v & 1, (v & 1) + 2, ... of the column
        double rho = 0.0;
        int kept = 0;  // entries that pass the exact test are written back in place: the force pass walks only those
#pragma unroll kSmallUnroll
        for (int s = 0; s < c; ++s) {
          // the search left the partner's index; image index S' = -rint(g_j - g_i) (the search's choice, recomputed
          // the same way) plus the atoms' own wrap counts, and the species of j, make the entry
          const int j = row[2 * s * NTL];
          const float4 gj = g32[j];
          int S0 = -(int)rintf(gj.x - gi.x), S1 = -(int)rintf(gj.y - gi.y), S2 = -(int)rintf(gj.z - gi.z);
          if (wrapped_atoms) {
            const int aj = ash[j];
            S0 += (ai & 1023) - (aj & 1023);
            S1 += ((ai >> 10) & 1023) - ((aj >> 10) & 1023);
            S2 += ((ai >> 20) & 1023) - ((aj >> 20) & 1023);
            if (abs(S0) > 1 || abs(S1) > 1 || abs(S2) > 1) range = true;  // a far image: the general path's case
          }
          const unsigned tidx = (unsigned)((S0 + 1) + 3 * (S1 + 1) + 9 * (S2 + 1)) & 31u;
          const unsigned en = (unsigned)j | (tidx << 9) | ((unsigned)spc[j] << 14);
          const double* pj = pos3 + 3 * j;
          const double* sv = svt + 3 * tidx;
          const double vx = xadd(xsub(pj[0], xi), sv[0]);
          const double vy = xadd(xsub(pj[1], yi), sv[1]);
          const double vz = xadd(xsub(pj[2], zi_), sv[2]);
          const double d2 = xnorm2(vx, vy, vz);
          if (d2 == 0.0) coincident = true;  // dist <= 0 guard, rgpot_cuh2.f90:422-428
          if (!(d2 < rc2_phys)) continue;    // :522
          const int zj = (int)(en >> 14);
          if (MM != EAM_AL && (zi | zj) != 0) {
            const int slot = atomicAdd(&dq_n, 1);
            if (slot < kDeferCap) {
              dq[slot] = ((unsigned)i << 16) | en;
              continue;
            }  // (queue full: the entry stays in the list and is evaluated in place)
          }
          row[2 * kept * NTL] = (unsigned short)en;  // (kept <= s: never ahead of the read cursor)
          ++kept;
          double r, rinv;
          radial(d2, r, rinv);
          const double on_i = eam_rho<MM>(zj, r);                                       // what j deposits on i
          const double on_j = (MM == EAM_AL || zi == zj) ? on_i : eam_rho<MM>(zi, r);  // what i deposits on j
          rho += on_i;
          range |= !(fabs(on_j) < 0.25);
          const long long q = __double_as_longlong(fma(on_j, 0x1p53, kFixMagic)) - __double_as_longlong(kFixMagic);
          fixed_add(rho_acc + j, q);
        }
        vcnt[v] = (unsigned short)kept;  // (the force pass gives this half list to the same thread)
        range |= !(fabs(rho) < 128.0);
        fixed_add(rho_acc + i, __double2ll_rn(rho * 0x1p53));
      }
      if (coincident) atomicOr(&blk_flags, kFlagCoincident);
      if (range) atomicOr(&blk_flags, kFlagListOverflow);
    }
    if (MM != EAM_AL) {
      __syncthreads();
      const int nq = min(dq_n, kDeferCap);
      for (int k = threadIdx.x; k < nq; k += blockDim.x) {
        const unsigned q = dq[k];
        const int i = (int)(q >> 16), j = (int)(q & 511u), zi = spc[i], zj = (int)((q >> 14) & 3u);
        const double* pi = pos3 + 3 * i;
        const double* pj = pos3 + 3 * j;
        const double* sv = svt + 3 * ((q >> 9) & 31u);
        const double vx = xadd(xsub(pj[0], pi[0]), sv[0]);
        const double vy = xadd(xsub(pj[1], pi[1]), sv[1]);
        const double vz = xadd(xsub(pj[2], pi[2]), sv[2]);
        double r, rinv;
        radial(xnorm2(vx, vy, vz), r, rinv);
        const double on_i = eam_rho<MM>(zj, r);
        const double on_j = zi == zj ? on_i : eam_rho<MM>(zi, r);
        if (!(fabs(on_i) < 0.25 && fabs(on_j) < 0.25)) atomicOr(&blk_flags, kFlagListOverflow);
        fixed_add(rho_acc + i, __double_as_longlong(fma(on_i, 0x1p53, kFixMagic)) - __double_as_longlong(kFixMagic));
        fixed_add(rho_acc + j, __double_as_longlong(fma(on_j, 0x1p53, kFixMagic)) - __double_as_longlong(kFixMagic));
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double emb, demb;
      eam_embed<MM>(spc[i], (double)rho_acc[i] * 0x1p-53, emb, demb);
      dfd[i] = demb;
      eatom[i] = emb;
    }
  } else
  for (int ib = 0; ib < n; ib += ngrp) {
    const int slot = ib + grp;
    const bool vi = slot < n;
    double rho = 0.0, rho_s = 0.0;
    bool coincident = false;
    int i = 0;
    if (vi) {
      i = perm[slot];
      const int c = cnt16[i * TPA + sub];
      const int zi = spc[i];
      const double xi = pos3[3 * i], yi = pos3[3 * i + 1], zi_ = pos3[3 * i + 2];
      const unsigned short* row = list + i * TPA + sub;
#pragma unroll kSmallUnroll
      for (int s = 0; s < c; ++s) {
        const unsigned en = row[s * NTL];
        const double* pj = pos3 + 3 * (en & 511u);
        const double* sv = svt + 3 * ((en >> 9) & 31u);
        const double vx = xadd(xsub(pj[0], xi), sv[0]);
        const double vy = xadd(xsub(pj[1], yi), sv[1]);
        const double vz = xadd(xsub(pj[2], zi_), sv[2]);
        const double d2 = xnorm2(vx, vy, vz);
        if (kFeHe) {
          if (!(d2 < rc2)) continue;  // vesin :1202; every term then checks its own cutoff on sqrt(d2)
          double r1, dr1;
          const int ch = fehe_density(zi, (int)(en >> 14), sqrt(d2), r1, dr1);  // atom_densities :480-500
          if (ch == 0) rho += r1;
          if (ch == 1) rho_s += r1;
        } else {
          if (d2 == 0.0) coincident = true;  // dist <= 0 guard, rgpot_cuh2.f90:422-428
          if (!(d2 < rc2_phys)) continue;    // :522
          rho += eam_density_entry<M == EAM_FEHE ? EAM_CUH2 : M>((int)(en >> 14), d2);
        }
      }
    }
#pragma unroll
    for (int off = TPA / 2; off > 0; off >>= 1) {
      rho += __shfl_xor_sync(0xffffffffu, rho, off);
      if (kFeHe) rho_s += __shfl_xor_sync(0xffffffffu, rho_s, off);
    }
    if (coincident) atomicOr(&blk_flags, kFlagCoincident);
    if (vi && sub == 0) {
      double emb, demb;
      if (kFeHe) {
        double dems;
        fehe_embed(spc[i], rho, rho_s, emb, demb, dems);
        dfs[i] = dems;
      } else {
        eam_embed<M == EAM_FEHE ? EAM_CUH2 : M>(spc[i], rho, emb, demb);
      }
      dfd[i] = demb;
      eatom[i] = emb;
    }
  }
  __syncthreads();
  if (blk_flags & kFlagCoincident) {
    for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) a.F[3 * base + t] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
      a.E[sys] = 0.0;
      if (a.single_n > 0) __threadfence_system();
      a.sys_flags[sys] = kFlagCoincident;
    }
    return;
  }

  RB_STAMP(6);
  // ---- phase E: pair energy and forces
  if (half) {
    for (int u = threadIdx.x; u < NTL; u += blockDim.x) {
      bool range = false;
#pragma unroll 1
      for (int hl = 0; hl < 2; ++hl) {
        const int v = perm2[hl ? NL - 1 - u : u];
        const int col = v >> 1, i = col / TPA, c = vcnt[v];
        const int zi = spc[i];
        const double dfd_i = dfd[i];
        const double xi = pos3[3 * i], yi = pos3[3 * i + 1], zi_ = pos3[3 * i + 2];
        const unsigned short* row = list + col + (v & 1) * NTL;
        double fx = 0.0, fy = 0.0, fz = 0.0, ep = 0.0;
#pragma unroll kSmallUnroll
        for (int s = 0; s < c; ++s) {
          const unsigned en = row[2 * s * NTL];
          const int j = (int)(en & 511u);
          const double* pj = pos3 + 3 * j;
          const double* sv = svt + 3 * ((en >> 9) & 31u);
          const double vx = xadd(xsub(pj[0], xi), sv[0]);
          const double vy = xadd(xsub(pj[1], yi), sv[1]);
          const double vz = xadd(xsub(pj[2], zi_), sv[2]);
          const double d2 = xnorm2(vx, vy, vz);  // (d2 < rc2_phys: the density pass kept only those)
          double phi, w;
          eam_force_pair<MM>(zi, (int)(en >> 14), dfd_i, dfd[j], d2, phi, w);
          ep += phi;  // the whole pair energy: the partner does not see this pair
          const double tx = w * vx, ty = w * vy, tz = w * vz;
          fx += tx;
          fy += ty;
          fz += tz;
          range |= !(fabs(w) < 128.0);  // |v| <= rc < 8: every component term below 1024
          fixed_add3(facc + 3 * j, __double_as_longlong(fma(-tx, 0x1p40, kFixMagic)) - __double_as_longlong(kFixMagic),
                     __double_as_longlong(fma(-ty, 0x1p40, kFixMagic)) - __double_as_longlong(kFixMagic),
                     __double_as_longlong(fma(-tz, 0x1p40, kFixMagic)) - __double_as_longlong(kFixMagic));
        }
        if (c > 0) {
          range |= !(fabs(fx) < 0x1p20 && fabs(fy) < 0x1p20 && fabs(fz) < 0x1p20 && fabs(ep) < 0x1p20);
          fixed_add3(facc + 3 * i, __double2ll_rn(fx * 0x1p40), __double2ll_rn(fy * 0x1p40), __double2ll_rn(fz * 0x1p40));
          fixed_add(e_fix + i, __double2ll_rn(ep * 0x1p40));
        }
      }
      if (range) atomicOr(&blk_flags, kFlagListOverflow);
    }
    if (MM != EAM_AL) {
      const int nq = min(dq_n, kDeferCap);
      for (int k = threadIdx.x; k < nq; k += blockDim.x) {
        const unsigned q = dq[k];
        const int i = (int)(q >> 16), j = (int)(q & 511u);
        const double* pi = pos3 + 3 * i;
        const double* pj = pos3 + 3 * j;
        const double* sv = svt + 3 * ((q >> 9) & 31u);
        const double vx = xadd(xsub(pj[0], pi[0]), sv[0]);
        const double vy = xadd(xsub(pj[1], pi[1]), sv[1]);
        const double vz = xadd(xsub(pj[2], pi[2]), sv[2]);
        double phi, w;
        eam_force_pair<MM>(spc[i], (int)((q >> 14) & 3u), dfd[i], dfd[j], xnorm2(vx, vy, vz), phi, w);
        if (!(fabs(w) < 128.0 && fabs(phi) < 1024.0)) atomicOr(&blk_flags, kFlagListOverflow);
        const double t[3] = {w * vx, w * vy, w * vz};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          fixed_add(facc + 3 * i + c, __double_as_longlong(fma(t[c], 0x1p40, kFixMagic)) - __double_as_longlong(kFixMagic));
          fixed_add(facc + 3 * j + c, __double_as_longlong(fma(-t[c], 0x1p40, kFixMagic)) - __double_as_longlong(kFixMagic));
        }
        fixed_add(&e_acc, __double_as_longlong(fma(phi, 0x1p40, kFixMagic)) - __double_as_longlong(kFixMagic));
      }
    }
    __syncthreads();
    if (blk_flags & kFlagListOverflow) {
      if (threadIdx.x == 0) a.sys_flags[sys] = kFlagListOverflow;  // out of the fixed-point range: the general path takes it
      return;
    }
    for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) a.F[3 * base + t] = (double)facc[t] * 0x1p-40;
  } else
  for (int ib = 0; ib < n; ib += ngrp) {
    const int slot = ib + grp;
    const bool vi = slot < n;
    double fx = 0.0, fy = 0.0, fz = 0.0, ep = 0.0;
    int i = 0;
    if (vi) {
      i = perm[slot];
      const int c = cnt16[i * TPA + sub];
      const int zi = spc[i];
      const double dfd_i = dfd[i];
      const double xi = pos3[3 * i], yi = pos3[3 * i + 1], zi_ = pos3[3 * i + 2];
      const unsigned short* row = list + i * TPA + sub;
#pragma unroll kSmallUnroll
      for (int s = 0; s < c; ++s) {
        const unsigned en = row[s * NTL];
        const int j = (int)(en & 511u);
        const double* pj = pos3 + 3 * j;
        const double* sv = svt + 3 * ((en >> 9) & 31u);
        const double vx = xadd(xsub(pj[0], xi), sv[0]);
        const double vy = xadd(xsub(pj[1], yi), sv[1]);
        const double vz = xadd(xsub(pj[2], zi_), sv[2]);
        const double d2 = xnorm2(vx, vy, vz);
        if (kFeHe) {
          if (!(d2 < rc2)) continue;
          const double r = sqrt(d2);
          const int zj = (int)(en >> 14);
          double v, dv, r1, dr1;
          fehe_pair(zi, zj, r, v, dv);  // atom_contribution :509-538
          const int ch = fehe_density(zi, zj, r, r1, dr1);
          ep += 0.5 * v;
          double radial = dv;
          if (ch == 0) radial += (dfd_i + dfd[j]) * dr1;
          if (ch == 1) radial += (dfs[i] + dfs[j]) * dr1;
          fx += radial * (vx / r);
          fy += radial * (vy / r);
          fz += radial * (vz / r);
        } else {
          if (!(d2 < rc2_phys)) continue;
          eam_force_entry<M == EAM_FEHE ? EAM_CUH2 : M>(zi, (int)(en >> 14), dfd_i, dfd[j], vx, vy, vz, d2, ep, fx, fy, fz);
        }
      }
    }
#pragma unroll
    for (int off = TPA / 2; off > 0; off >>= 1) {
      fx += __shfl_xor_sync(0xffffffffu, fx, off);
      fy += __shfl_xor_sync(0xffffffffu, fy, off);
      fz += __shfl_xor_sync(0xffffffffu, fz, off);
      ep += __shfl_xor_sync(0xffffffffu, ep, off);
    }
    if (vi && sub == 0) {
      a.F[3 * (base + i)] = fx;
      a.F[3 * (base + i) + 1] = fy;
      a.F[3 * (base + i) + 2] = fz;
      eatom[i] += ep;
    }
  }
  RB_STAMP(7);
  __syncthreads();  // (also: every thread's force stores are ordered before thread 0's system fence below)
  RB_STAMP(8);
  for (int i = threadIdx.x; i < n; i += blockDim.x)  // atom order: reproducible
    e += half ? eatom[i] + (double)e_fix[i] * 0x1p-40 : eatom[i];
  const double bs = block_sum(e, sh32);
  if (threadIdx.x == 0) {
    a.E[sys] = half ? bs + (double)e_acc * 0x1p-40 : bs;  // (+ the pair energies of the set-aside entries)
    RB_STAMP(9);
    if (a.single_n > 0) __threadfence_system();  // (only a polling host needs the order; a batch must not pay for it)
    RB_STAMP(10);
    a.sys_flags[sys] = 0u;
  }
}

}  // namespace rb
