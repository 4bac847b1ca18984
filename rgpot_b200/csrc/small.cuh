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
  int cluster;             // LJ: free boundaries
  double u0, psi2, shiftU, cutoff;  // LJ
  EmitArgs emit;           // optional (single system only)
};

// ------------------------------------------------------------------------------------------ LJ
__global__ void __launch_bounds__(256)
k_lj_small(SmallArgs a) {
  extern __shared__ double smem[];
  __shared__ double sh32[32];
  const int sys = blockIdx.x;
  const long long base = a.off[sys];
  const int n = (int)(a.off[sys + 1] - base);
  double* px = smem;  // 3n interleaved
  for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) px[t] = a.pos[3 * base + t];

  // fold parameters: PairListCache.hpp:176-184 ; LJClusterPot.cc:46-55 (no fold at all)
  LJParams p;
  p.u0 = a.u0;
  p.psi2 = a.psi2;
  p.shiftU = a.shiftU;
  p.rc2 = xmul(a.cutoff, a.cutoff);
  const double* box = a.boxes + 9 * (size_t)sys;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p.w[k] = box[4 * k];
    p.inv[k] = (!a.cluster && p.w[k] != 0.0) ? 1.0 / p.w[k] : 0.0;
  }
  __syncthreads();

  double e = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double xi = px[3 * i], yi = px[3 * i + 1], zi = px[3 * i + 2];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      const double dx = fold_component(xsub(px[3 * j], xi), p.w[0], p.inv[0]);
      const double dy = fold_component(xsub(px[3 * j + 1], yi), p.w[1], p.inv[1]);
      const double dz = fold_component(xsub(px[3 * j + 2], zi), p.w[2], p.inv[2]);
      const double r2 = xnorm2(dx, dy, dz);
      if (!(r2 <= p.rc2)) continue;
      lj_pair(p, dx, dy, dz, r2, e, fx, fy, fz);
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
  const double bs = block_sum(e, sh32);
  if (threadIdx.x == 0) a.E[sys] = bs;
}

// ---------------------------------------------------------------------------------------- CuH2
struct SmallBox {
  double H[9], Hinv[9];
  double q[3];  // rc / face distance (+ margin): |fractional separation| bound per direction
  int fast;     // every q < 0.5: at most one image per pair and direction
  int bad;      // singular box
};

// vesin Matrix::determinant / inverse (:100-128) and face distances (:258-275)
RB_D void small_box_init(const double* b, double rc, SmallBox& sb) {
  for (int k = 0; k < 9; ++k) sb.H[k] = b[k];
  const double* m = sb.H;
  const double det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                     m[2] * (m[3] * m[7] - m[4] * m[6]);
  sb.bad = !(fabs(det) >= 1e-30);
  if (sb.bad) {
    sb.fast = 0;
    return;
  }
  double* v = sb.Hinv;
  v[0] = (m[4] * m[8] - m[7] * m[5]) / det;
  v[1] = (m[2] * m[7] - m[1] * m[8]) / det;
  v[2] = (m[1] * m[5] - m[2] * m[4]) / det;
  v[3] = (m[5] * m[6] - m[3] * m[8]) / det;
  v[4] = (m[0] * m[8] - m[2] * m[6]) / det;
  v[5] = (m[3] * m[2] - m[0] * m[5]) / det;
  v[6] = (m[3] * m[7] - m[6] * m[4]) / det;
  v[7] = (m[6] * m[1] - m[0] * m[7]) / det;
  v[8] = (m[0] * m[4] - m[3] * m[1]) / det;
  // face distance k = 1 / |column k of Hinv|  (equivalent to |n_k . H_k| with n_k the unit normal)
  sb.fast = 1;
  for (int k = 0; k < 3; ++k) {
    const double cn = sqrt(v[k] * v[k] + v[3 + k] * v[3 + k] + v[6 + k] * v[6 + k]);
    sb.q[k] = rc * cn * (1.0 + 1e-9) + 1e-9;
    if (!(sb.q[k] < 0.4999)) sb.fast = 0;
  }
}

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

// TPA = lanes cooperating on one atom (power of two, <= 32)
template <int TPA>
__global__ void __launch_bounds__(512, 2)
k_eam_small(SmallArgs a) {
  extern __shared__ double smem[];
  __shared__ double sh32[32];
  __shared__ SmallBox sb;
  __shared__ unsigned blk_flags;
  __shared__ int blk_bad[2];

  const int sys = blockIdx.x;
  const long long base = a.off[sys];
  const int n = (int)(a.off[sys + 1] - base);
  const int maxnb = a.maxnb;

  // shared layout
  double* px = smem;
  double* py = px + n;
  double* pz = py + n;
  double* gx = pz + n;  // wrapped fractional coordinates
  double* gy = gx + n;
  double* gz = gy + n;
  double* dfd = gz + n;
  int* ash = (int*)(dfd + n);        // packed wrap counts (10 bits each, biased 512)
  int* cnt = ash + n;                // neighbors per atom
  int* spc = cnt + n;                // species
  unsigned* list = (unsigned*)(spc + n);  // n * maxnb entries

  if (threadIdx.x == 0) {
    small_box_init(a.boxes + 9 * (size_t)sys, c_eam.rcut, sb);
    blk_flags = sb.bad ? kFlagSingularBox : 0u;  // vesin: "the box matrix is not invertible"
    blk_bad[0] = INT_MAX;
    blk_bad[1] = 0;
  }
  __syncthreads();

  // ---- load, classify (rgpot_cuh2.f90:463-489), wrap into the cell
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = a.pos[3 * (base + i)], y = a.pos[3 * (base + i) + 1], z = a.pos[3 * (base + i) + 2];
    px[i] = x;
    py[i] = y;
    pz[i] = z;
    const int zn = a.Z[base + i];
    int sp = 0;
    if (zn == 29) {
      sp = 0;
    } else if (zn == 1) {
      sp = 1;
    } else {
      atomicOr(&blk_flags, kFlagForeignZ);
      atomicMin(&blk_bad[0], i);
    }
    spc[i] = sp;
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) {
      atomicOr(&blk_flags, kFlagNonFinite);
      gx[i] = gy[i] = gz[i] = 0.0;
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
    gx[i] = f[0];
    gy[i] = f[1];
    gz[i] = f[2];
    if (abs(w[0]) > 500 || abs(w[1]) > 500 || abs(w[2]) > 500) atomicOr(&blk_flags, kFlagShiftRange);
    ash[i] = ((w[0] + 512) & 1023) | (((w[1] + 512) & 1023) << 10) | (((w[2] + 512) & 1023) << 20);
  }
  __syncthreads();
  if (blk_flags & (kFlagForeignZ | kFlagNonFinite | kFlagShiftRange | kFlagSingularBox)) {
    // the reference leaves energy and forces zeroed on failure
    for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) a.F[3 * base + t] = 0.0;
    if (threadIdx.x == 0) {
      a.E[sys] = 0.0;
      a.sys_flags[sys] = blk_flags;
      a.sys_bad[2 * sys] = blk_bad[0];
      a.sys_bad[2 * sys + 1] = (blk_bad[0] != INT_MAX) ? a.Z[base + blk_bad[0]] : 0;
    }
    return;
  }

  const int sub = threadIdx.x % TPA;
  const int grp = threadIdx.x / TPA;
  const int ngrp = blockDim.x / TPA;
  const unsigned grp_shift = (threadIdx.x & 31) - sub;  // first lane of this group in the warp
  const unsigned grp_mask = (TPA == 32) ? 0xffffffffu : (((1u << TPA) - 1u) << grp_shift);
  const double rc2 = c_eam.rc2;
  const double rcut = c_eam.rcut;

  // ---- neighbor search: all pairs, candidate image from the wrapped fractional separation
  for (int ib = 0; ib < n; ib += ngrp) {
    const int i = ib + grp;
    const bool vi = i < n;
    int count = 0;
    if (sb.fast) {
      const double gxi = vi ? gx[i] : 0.0, gyi = vi ? gy[i] : 0.0, gzi = vi ? gz[i] : 0.0;
      const int ai = vi ? ash[i] : 0;
      for (int jb = 0; jb < n; jb += TPA) {
        const int j = jb + sub;
        bool acc = false;
        int S0 = 0, S1 = 0, S2 = 0;
        if (vi && j < n && j != i) {
          double t0 = gx[j] - gxi, t1 = gy[j] - gyi, t2 = gz[j] - gzi;
          if (t0 > 0.5) { t0 -= 1.0; S0 = -1; } else if (t0 < -0.5) { t0 += 1.0; S0 = 1; }
          if (t1 > 0.5) { t1 -= 1.0; S1 = -1; } else if (t1 < -0.5) { t1 += 1.0; S1 = 1; }
          if (t2 > 0.5) { t2 -= 1.0; S2 = -1; } else if (t2 < -0.5) { t2 += 1.0; S2 = 1; }
          if (fabs(t0) <= sb.q[0] && fabs(t1) <= sb.q[1] && fabs(t2) <= sb.q[2]) {
            const int aj = ash[j];
            S0 += (ai & 1023) - (aj & 1023);
            S1 += ((ai >> 10) & 1023) - ((aj >> 10) & 1023);
            S2 += ((ai >> 20) & 1023) - ((aj >> 20) & 1023);
            double vx, vy, vz;
            acc = entry_vec(sb, px, py, pz, i, j, S0, S1, S2, vx, vy, vz) < rc2;
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        const unsigned gb = (bal & grp_mask) >> grp_shift;
        if (acc) {
          const int slot = count + __popc(gb & ((1u << sub) - 1u));
          if (slot < maxnb && abs(S0) < 16 && abs(S1) < 16 && abs(S2) < 16) {
            list[(size_t)i * maxnb + slot] = pack_entry(j, S0, S1, S2);
          } else {
            atomicOr(&blk_flags, kFlagListOverflow);
          }
        }
        count += __popc(gb);
      }
    } else if (vi && sub == 0) {
      // small boxes: several images per pair; the group leader enumerates them in order
      const int ai = ash[i];
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;  // self images are dropped (rgpot_neighbors.f90:132)
        const double t0 = gx[j] - gx[i], t1 = gy[j] - gy[i], t2 = gz[j] - gz[i];
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
              if (entry_vec(sb, px, py, pz, i, j, S0, S1, S2, vx, vy, vz) < rc2) {
                if (count < maxnb && abs(S0) < 16 && abs(S1) < 16 && abs(S2) < 16) {
                  list[(size_t)i * maxnb + count] = pack_entry(j, S0, S1, S2);
                } else {
                  atomicOr(&blk_flags, kFlagListOverflow);
                }
                ++count;
              }
            }
      }
    }
    if (!sb.fast) count = __shfl_sync(0xffffffffu, count, grp_shift);  // leader's count
    if (vi && sub == 0) cnt[i] = count;
  }
  __syncthreads();
  if (blk_flags & kFlagListOverflow) {
    if (threadIdx.x == 0) a.sys_flags[sys] = kFlagListOverflow;  // host reruns on the cell path
    return;
  }

  // optional emission of the table (rgpot_b200_neighbors)
  if (a.emit.pairs != nullptr) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      for (int s = 0; s < cnt[i]; ++s) {
        const unsigned en = list[(size_t)i * maxnb + s];
        const unsigned long long k = atomicAdd(a.emit.count, 1ull);
        if ((long)k < a.emit.cap) {
          a.emit.pairs[2 * k] = i;
          a.emit.pairs[2 * k + 1] = (int)(en & 0xffffu);
          a.emit.shifts[3 * k] = (int)((en >> 16) & 31u) - 16;
          a.emit.shifts[3 * k + 1] = (int)((en >> 21) & 31u) - 16;
          a.emit.shifts[3 * k + 2] = (int)((en >> 26) & 31u) - 16;
        }
      }
    }
  }

  double e = 0.0;
  // ---- pass 1 + 2: density, embedding (rgpot_cuh2.f90:436-451)
  for (int ib = 0; ib < n; ib += ngrp) {
    const int i = ib + grp;
    const bool vi = i < n;
    double rho = 0.0;
    bool coincident = false;
    if (vi) {
      const int c = cnt[i];
      for (int s = sub; s < c; s += TPA) {
        const unsigned en = list[(size_t)i * maxnb + s];
        const int j = (int)(en & 0xffffu);
        double vx, vy, vz;
        const double d2 = entry_vec(sb, px, py, pz, i, j, (int)((en >> 16) & 31u) - 16,
                                    (int)((en >> 21) & 31u) - 16, (int)((en >> 26) & 31u) - 16, vx,
                                    vy, vz);
        const double r = sqrt(d2);
        if (r <= 0.0) coincident = true;
        if (r >= rcut) continue;
        rho += eam_rho(spc[j], r);
      }
    }
#pragma unroll
    for (int off = TPA / 2; off > 0; off >>= 1) rho += __shfl_xor_sync(0xffffffffu, rho, off);
    if (coincident) atomicOr(&blk_flags, kFlagCoincident);
    if (vi && sub == 0) {
      double emb, demb;
      eam_embed(spc[i], rho, emb, demb);
      dfd[i] = demb;
      e += emb;
    }
  }
  __syncthreads();
  if (blk_flags & kFlagCoincident) {
    for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) a.F[3 * base + t] = 0.0;
    if (threadIdx.x == 0) {
      a.E[sys] = 0.0;
      a.sys_flags[sys] = kFlagCoincident;
    }
    return;
  }

  // ---- pass 3: pair energy and forces (rgpot_cuh2.f90:454-457)
  for (int ib = 0; ib < n; ib += ngrp) {
    const int i = ib + grp;
    const bool vi = i < n;
    double fx = 0.0, fy = 0.0, fz = 0.0, ep = 0.0;
    if (vi) {
      const int c = cnt[i];
      const int zi = spc[i];
      const double dfd_i = dfd[i];
      for (int s = sub; s < c; s += TPA) {
        const unsigned en = list[(size_t)i * maxnb + s];
        const int j = (int)(en & 0xffffu);
        double vx, vy, vz;
        const double d2 = entry_vec(sb, px, py, pz, i, j, (int)((en >> 16) & 31u) - 16,
                                    (int)((en >> 21) & 31u) - 16, (int)((en >> 26) & 31u) - 16, vx,
                                    vy, vz);
        const double r = sqrt(d2);
        if (r >= rcut) continue;
        eam_force_entry(zi, spc[j], dfd_i, dfd[j], vx, vy, vz, r, ep, fx, fy, fz);
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
      e += ep;
    }
  }
  const double bs = block_sum(e, sh32);
  if (threadIdx.x == 0) {
    a.E[sys] = bs;
    a.sys_flags[sys] = 0u;
  }
}

}  // namespace rb
