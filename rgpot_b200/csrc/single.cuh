// single.cuh -- ONE small EAM system (about 64 .. 512 atoms) spread over many blocks.
//
// The batch kernel (small.cuh) gives a system one thread block: right for 65 536 systems per launch,
// wrong for the single 218-atom slab of BASELINE configs[1] or the 130-atom slab of the reference's own
// benchmark (CppCore/tests/PotBench.cc:262-321), where one SM works and 147 watch -- 40+ microseconds of
// kernel for 11 954 pair terms.  Here a block owns 16 atoms, sixteen lanes share an atom and walk every
// sixteenth partner, and search and physics are one loop (no neighbor lists: with 14 partners per lane
// there is nothing to amortise):
//
//   k_eam_one   setup (once) -> rho_i, F(rho_i), F'(rho_i)      (rgpot_cuh2.f90:436-451, 506-530)
//               -> grid barrier (a cooperative launch: every block is resident)
//               -> 0.5 phi, forces, E = sum                      (rgpot_cuh2.f90:454-459, 536-587)
//
// Pair selection is the batch kernel's: nearest image in fractional space (FP32) proposes the image,
// the exact vesin expression (p_j - p_i) + S.H with d2 < rc2 (FP64, unfused) decides.  The kernels only
// take the clean case -- a cell wide enough that a pair has at most one image inside the cutoff, known
// species, finite coordinates, atoms within one cell vector of the cell, no coincident atoms; anything
// else raises `fallback` and the host reruns the system on the batch kernel, which owns the error
// reporting.  Sums have a fixed order (partner index mod 16, then a four-level shuffle tree; block
// energies folded in block order by the last block): bit-reproducible.
#pragma once

#include <cooperative_groups.h>

#include "small.cuh"

namespace rb {

struct OneArgs {
  const double* pos;   // 3 n (device copy of the caller's positions)
  const int* Z;        // n atomic numbers (CuH2), unused for EAM-Al
  const double* box;   // 9 (device copy; the kernels use the prepared copy below)
  SmallBox sb;         // cell, inverse, |fractional separation| bounds: prepared on the host (vesin's formulas)
  SmallBoxF sbf;       // FP32 mirror for the prefilter
  int n;
  double* dfd;         // n: F'(rho)       (density -> force)
  double* eemb;        // n: F(rho)
  unsigned* fallback;  // device word: == call_id means "the batch kernel has to take this system"
  unsigned call_id;    // increases with every call, so the word never needs a reset
  double* F;           // 3 n, mapped host memory
  double* f_stage;     // 3 n, device scratch: the last block moves the forces to F (one system-scope fence per call)
  double* E;           // 1, mapped host memory
  unsigned* flags;     // mapped host status word (written last, behind a system fence)
  double* e_parts;     // per-block energies
  int* ticket;         // blocks that have delivered (left at zero)
};

constexpr int kOneAtomsPerBlock = 8;
constexpr int kOneLanes = 32;
constexpr unsigned kOneFallback = 0x40000000u;  // flags value: "rerun on the batch kernel"

// shared setup of both kernels: positions, species, wrapped fractional coordinates (FP32), wrap counts,
// the 27 image vectors; returns false (uniformly) if the system is not the clean case
template <int M>
RB_D bool one_setup(const OneArgs& a, double* pos3, float4* g32, int* ash, int* spc, double* svt, const SmallBox& sb,
                    unsigned& bad) {
  const int n = a.n;
  if (threadIdx.x == 0) bad = 0u;
  __syncthreads();
  if (threadIdx.x < 27) {
    const int t = (int)threadIdx.x;
    const double s0 = (double)(t % 3 - 1), s1 = (double)((t / 3) % 3 - 1), s2 = (double)(t / 9 - 1);
    svt[3 * t] = shift_component(sb.H, 0, s0, s1, s2);
    svt[3 * t + 1] = shift_component(sb.H, 1, s0, s1, s2);
    svt[3 * t + 2] = shift_component(sb.H, 2, s0, s1, s2);
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = a.pos[3 * i], y = a.pos[3 * i + 1], z = a.pos[3 * i + 2];
    pos3[3 * i] = x;
    pos3[3 * i + 1] = y;
    pos3[3 * i + 2] = z;
    int sp = 0;
    if (M != EAM_AL) {
      const int zn = a.Z[i];
      if (zn == 1)
        sp = 1;
      else if (zn != 29)
        bad = 1u;  // (a benign race: every writer stores the same value)
    }
    spc[i] = sp;
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) {
      bad = 1u;
      g32[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      ash[i] = 0;
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
    g32[i] = make_float4((float)f[0], (float)f[1], (float)f[2], 0.f);
    if (abs(w[0]) > 1 || abs(w[1]) > 1 || abs(w[2]) > 1) bad = 1u;  // far outside the cell: the general kernels' case
    ash[i] = (w[0] + 1) | ((w[1] + 1) << 2) | ((w[2] + 1) << 4);
  }
  __syncthreads();
  return bad == 0u;
}

// walk the partners j = sub, sub + kOneLanes, ... of atom i: body(j, vx, vy, vz, d2) for every accepted entry
template <class Body>
RB_D bool one_walk(int n, int i, int sub, const double* pos3, const float4* g32, const int* ash, const double* svt,
                   const SmallBoxF& sbf, double rc2, Body&& body) {
  const float4 gi = g32[i];
  const int ai = ash[i];
  const double xi = pos3[3 * i], yi = pos3[3 * i + 1], zi = pos3[3 * i + 2];
  bool far = false;
  for (int j = sub; j < n; j += kOneLanes) {
    const float4 c = g32[j];
    float t0 = c.x - gi.x, t1 = c.y - gi.y, t2 = c.z - gi.z;
    const float r0 = rintf(t0), r1 = rintf(t1), r2 = rintf(t2);
    t0 -= r0;
    t1 -= r1;
    t2 -= r2;
    const float dx = fmaf(t0, sbf.h[0], fmaf(t1, sbf.h[3], t2 * sbf.h[6]));
    const float dy = fmaf(t0, sbf.h[1], fmaf(t1, sbf.h[4], t2 * sbf.h[7]));
    const float dz = fmaf(t0, sbf.h[2], fmaf(t1, sbf.h[5], t2 * sbf.h[8]));
    if (!(fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= sbf.thr2) || j == i) continue;
    // image of j seen from i: the search's choice plus the atoms' own wrap counts (vesin :1497)
    const int aj = ash[j];
    const int S0 = -(int)r0 + (ai & 3) - (aj & 3), S1 = -(int)r1 + ((ai >> 2) & 3) - ((aj >> 2) & 3),
              S2 = -(int)r2 + ((ai >> 4) & 3) - ((aj >> 4) & 3);
    if (abs(S0) > 1 || abs(S1) > 1 || abs(S2) > 1) {
      far = true;
      continue;
    }
    const double* sv = svt + 3 * ((S0 + 1) + 3 * (S1 + 1) + 9 * (S2 + 1));
    const double vx = xadd(xsub(pos3[3 * j], xi), sv[0]);
    const double vy = xadd(xsub(pos3[3 * j + 1], yi), sv[1]);
    const double vz = xadd(xsub(pos3[3 * j + 2], zi), sv[2]);
    const double d2 = xnorm2(vx, vy, vz);
    if (d2 < rc2) body(j, vx, vy, vz, d2);
  }
  return far;
}

RB_HD size_t one_smem_bytes(int n) {
  return sizeof(double) * (4 * (size_t)n + 81 + 2) + sizeof(float4) * (size_t)n + sizeof(int) * 2 * (size_t)n + 64;
}

// host side of small_box_init (the same formulas; used once per call instead of once per block)
inline bool one_prepare_box(const double* b, double rc, SmallBox& sb, SmallBoxF& sbf) {
  for (int k = 0; k < 9; ++k) sb.H[k] = b[k];
  const double* m = sb.H;
  const double det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  sb.bad = !(fabs(det) >= 1e-30);
  sb.fast = 0;
  if (sb.bad) return false;
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
  sb.fast = 1;
  double hmax = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double cn = sqrt(v[k] * v[k] + v[3 + k] * v[3 + k] + v[6 + k] * v[6 + k]);
    sb.q[k] = rc * cn * (1.0 + 1e-9) + 1e-9;
    if (!(sb.q[k] < 0.4999)) sb.fast = 0;
  }
  for (int k = 0; k < 9; ++k) {
    sbf.h[k] = (float)sb.H[k];
    hmax += fabs(sb.H[k]);
  }
  sbf.ortho = 0;
  const double margin = 1.0e-6 * hmax + 1.0e-5 * rc;
  const double thr = rc + margin;
  sbf.thr2 = (float)(thr * thr * (1.0 + 1.0e-6));
  return sb.fast != 0;
}

template <int M>
__global__ void __launch_bounds__(kOneAtomsPerBlock* kOneLanes) k_eam_one(const __grid_constant__ OneArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned bad;
  __shared__ double sh32[32];
  __shared__ int last;
  const int n = a.n;
  double* pos3 = reinterpret_cast<double*>(smem_raw);
  double* dfd = pos3 + 3 * n;
  double* svt = dfd + n;
  float4* g32 = reinterpret_cast<float4*>(svt + 82);
  int* ash = reinterpret_cast<int*>(g32 + n);
  int* spc = ash + n;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  RB_STAMP(0);
  exp_table_init();
  RB_STAMP(1);
  const bool ok = one_setup<M>(a, pos3, g32, ash, spc, svt, a.sb, bad);
  const EamParams& P = eam_par<M>();
  const int i = blockIdx.x * kOneAtomsPerBlock + (int)(threadIdx.x / kOneLanes), sub = threadIdx.x % kOneLanes;
  RB_STAMP(2);
  // ---- density, embedding
  {
    double rho = 0.0;
    bool coincident = false, far = false;
    if (ok && i < n) {
      far = one_walk(n, i, sub, pos3, g32, ash, svt, a.sbf, P.rc2, [&](int j, double, double, double, double d2) {
        if (d2 == 0.0) coincident = true;  // dist <= 0 guard, rgpot_cuh2.f90:422-428: the batch kernel reports it
        if (d2 < P.rc2_phys) rho += eam_density_entry<M>(spc[j], d2);
      });
    }
#pragma unroll
    for (int off = 1; off < kOneLanes; off <<= 1) rho += __shfl_xor_sync(0xffffffffu, rho, off);
    if (!ok || far || coincident) atomicMax(a.fallback, a.call_id);
    if (ok && i < n && sub == 0) {
      double emb, demb;
      eam_embed<M>(spc[i], rho, emb, demb);
      a.dfd[i] = demb;
      a.eemb[i] = emb;
    }
  }
  RB_STAMP(3);
  grid.sync();  // every atom's F'(rho) is in memory
  RB_STAMP(4);
  if (__ldcg(a.fallback) == a.call_id) {
    // not this kernel's case: the batch kernel takes the system (and owns the error reporting)
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.flags = kOneFallback;
    return;
  }
  for (int t = threadIdx.x; t < n; t += blockDim.x) dfd[t] = __ldcg(a.dfd + t);
  __syncthreads();
  RB_STAMP(5);
  // ---- pair energy, forces
  double fx = 0.0, fy = 0.0, fz = 0.0, ep = 0.0;
  if (i < n) {
    const int zi = spc[i];
    const double dfd_i = dfd[i];
    one_walk(n, i, sub, pos3, g32, ash, svt, a.sbf, P.rc2, [&](int j, double vx, double vy, double vz, double d2) {
      if (d2 < P.rc2_phys) eam_force_entry<M>(zi, spc[j], dfd_i, dfd[j], vx, vy, vz, d2, ep, fx, fy, fz);
    });
  }
#pragma unroll
  for (int off = 1; off < kOneLanes; off <<= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, off);
    fy += __shfl_xor_sync(0xffffffffu, fy, off);
    fz += __shfl_xor_sync(0xffffffffu, fz, off);
    ep += __shfl_xor_sync(0xffffffffu, ep, off);
  }
  double e = 0.0;
  if (i < n && sub == 0) {
    a.f_stage[3 * i] = fx;
    a.f_stage[3 * i + 1] = fy;
    a.f_stage[3 * i + 2] = fz;
    e = ep + __ldcg(a.eemb + i);
  }
  RB_STAMP(6);
  const double bs = block_sum(e, sh32);  // (its barriers put every thread's force stores before thread 0's fence)
  RB_STAMP(7);
  if (threadIdx.x == 0) {
    a.e_parts[blockIdx.x] = bs;
    __threadfence();  // this block's staged forces before its ticket
    RB_STAMP(8);
    last = atomicAdd(a.ticket, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) RB_STAMP_ANY(9);
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
    RB_STAMP_ANY(10);
    __threadfence_system();
    RB_STAMP_ANY(11);
    *a.flags = 0u;
  }
}

}  // namespace rb
