// tile.cuh -- warp-per-cell tile kernels for one large system (the >= 100k-atom configurations).
//
// The thread-per-atom gather of gather.cuh pays the full FP64 pair math under a 15 % accept rate
// (divergence) and re-reads every neighbour record once per atom.  Here one WARP owns one cell:
//
//   * the cell's atoms (<= 32 per pass) sit in shared memory, FP64 and as FP32 offsets from a
//     local origin;
//   * the candidates of the 27-cell stencil are flattened into one virtual array (prefix table
//     over the stencil cells), eight candidates per lane are kept in registers as FP32 offsets;
//   * every (atom, candidate) pair takes a 7-instruction FP32 prefilter (FP32 pipe; the FP64 pipe
//     stays free); survivors are compacted through a per-warp queue and take the EXACT FP64 test
//     and the pair math 32 at a time on full lanes -- the accept/reject decision is still the
//     reference's expression bit for bit ((p_j - p_i) + shift, unfused norm);
//   * per-atom sums are formed by a fixed-order segmented scan over the queue order and added
//     to warp-private shared accumulators: no atomics, bit-reproducible.
//
// Valid for the fast geometry only (every periodic direction >= 3 cells, no atom outside the
// cell, one image per stencil cell); the kernels return immediately when the binning flags say
// otherwise and the gather kernels of gather.cuh take over.
#pragma once

#include "gather.cuh"

namespace rb {

constexpr int kTileWarps = 8;    // warps per block
constexpr int kTileChunk = 256;  // candidates per chunk (8 per lane)

struct TileWarp {
  double pix[32], piy[32], piz[32];  // the cell's atoms (current pass of <= 32)
  double fax[32], fay[32], faz[32];  // their force accumulators
  double acc2[32];                   // second accumulator (EAM: rho)
  float4 ri[32];                     // FP32 offsets from the local origin
  double seg_sv[27][3];              // image vector of each stencil cell
  int seg_start[27];
  int seg_prefix[28];
  int cand_j[kTileChunk];
  unsigned char cand_seg[kTileChunk];
};

// 1/a for normal positive a: hardware seed + two Newton steps (~1 ulp), no slow path
RB_D double rcp_fast(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  return fma(y, e, y);
}

// stencil table of cell (cx, cy, cz): lane s < 27 describes stencil cell s.  LJ: image vector
// w*q (exact); EAM: S.H with vesin's operation order.  Returns the number of candidates.
template <bool LJ>
RB_D int tile_build_segments(const GatherArgs& a, TileWarp& w, int cx, int cy, int cz, int lane) {
  int count = 0;
  if (lane < 27) {
    const int d[3] = {lane % 3 - 1, (lane / 3) % 3 - 1, lane / 9 - 1};
    const int c[3] = {cx, cy, cz};
    int q[3], r[3];
    bool valid = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      q[k] = 0;
      r[k] = c[k] + d[k];
      if (a.box.periodic[k]) {
        divmod_i(c[k] + d[k], a.grid.nc[k], q[k], r[k]);
      } else if (r[k] < 0 || r[k] >= a.grid.nc[k]) {
        valid = false;
      }
    }
    if (valid) {
      const int nb = (r[2] * a.grid.nc[1] + r[1]) * a.grid.nc[0] + r[0];
      const int jb = a.starts[nb];
      count = a.starts[nb + 1] - jb;
      w.seg_start[lane] = jb;
      if (LJ) {
        w.seg_sv[lane][0] = xmul(a.lj.w[0], (double)q[0]);
        w.seg_sv[lane][1] = xmul(a.lj.w[1], (double)q[1]);
        w.seg_sv[lane][2] = xmul(a.lj.w[2], (double)q[2]);
      } else {
        w.seg_sv[lane][0] = shift_component(a.box.H, 0, (double)q[0], (double)q[1], (double)q[2]);
        w.seg_sv[lane][1] = shift_component(a.box.H, 1, (double)q[0], (double)q[1], (double)q[2]);
        w.seg_sv[lane][2] = shift_component(a.box.H, 2, (double)q[0], (double)q[1], (double)q[2]);
      }
    } else {
      w.seg_start[lane] = 0;
    }
  }
  int incl = count;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane < 27) w.seg_prefix[lane] = incl - count;
  if (lane == 26) w.seg_prefix[27] = incl;
  const int total = __shfl_sync(0xffffffffu, incl, 26);
  __syncwarp();
  return total;
}

// stencil cell that holds virtual candidate v
RB_D int tile_find_segment(const TileWarp& w, int v) {
  int s = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const int t = s + step;
    if (t <= 26 && w.seg_prefix[t] <= v) s = t;
  }
  return s;
}

// fixed-order inclusive segmented scan over runs of equal keys (runs are contiguous)
RB_D double seg_scan(double v, int key, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, off);
    const int k = __shfl_up_sync(0xffffffffu, key, off);
    if (lane >= off && k == key) v += t;
  }
  return v;
}

// ---- bitmap tiles ------------------------------------------------------------------------------
// Per chunk of 256 candidates the FP32 prefilter leaves, for every atom of the cell, a 256-bit
// survivor mask (8 ballots).  The exact phase then gives every atom its own lanes (TPA of them),
// which walk the set bits in canonical (word, bit) order, redo the separation in FP64 exactly as
// the reference does, and accumulate in registers: no queue, no cross-lane reduction per batch,
// a fixed summation order.

// ------------------------------------------------------------------------------------------ LJ
__global__ void __launch_bounds__(kTileWarps * 32)
k_lj_tile(GatherArgs a, float thr2f, double* __restrict__ F, double* __restrict__ partials) {
  __shared__ TileWarp sm[kTileWarps];
  __shared__ unsigned masks[kTileWarps][32][8];  // [atom][word]: survivors of the prefilter
  if (a.st_in->flags & kFlagAtomShift) return;   // the literal-fold gather kernel takes over
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int cell = blockIdx.x * kTileWarps + wid;
  if (cell >= a.grid.ncells) return;
  TileWarp& w = sm[wid];
  unsigned(*mask)[8] = masks[wid];
  const int ib = a.starts[cell], ni_total = a.starts[cell + 1] - ib;
  if (ni_total == 0) {
    if (lane == 0) partials[cell] = 0.0;
    return;
  }
  const Grid& g = a.grid;
  const int cx = cell % g.nc[0], cy = (cell / g.nc[0]) % g.nc[1], cz = cell / (g.nc[0] * g.nc[1]);
  const int ncand = tile_build_segments<true>(a, w, cx, cy, cz, lane);
  const double4 o4 = a.spos[ib];  // local origin of the FP32 offsets
  const double rc2 = a.lj.rc2;
  double e_lane = 0.0;

  for (int i0 = 0; i0 < ni_total; i0 += 32) {
    const int ni = min(32, ni_total - i0);
    // lanes per atom in the exact phase
    const int tpa = ni <= 4 ? 8 : (ni <= 8 ? 4 : (ni <= 16 ? 2 : 1));
    const int my_il = lane / tpa, my_sub = lane % tpa;
    const bool have = my_il < ni;
    double pix = 0.0, piy = 0.0, piz = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    int orig_i = 0;
    if (lane < ni) {
      const double4 p = a.spos[ib + i0 + lane];
      w.ri[lane] = make_float4((float)(p.x - o4.x), (float)(p.y - o4.y), (float)(p.z - o4.z), 0.f);
    }
    if (have) {
      const double4 p = a.spos[ib + i0 + my_il];
      pix = p.x;
      piy = p.y;
      piz = p.z;
      orig_i = meta_orig((unsigned long long)__double_as_longlong(p.w));
    }
    __syncwarp();

    for (int ch = 0; ch < ncand; ch += kTileChunk) {
      float cxr[8], cyr[8], czr[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int v = ch + r * 32 + lane;
        cxr[r] = cyr[r] = czr[r] = __int_as_float(0x7fc00000);  // NaN never passes
        if (v < ncand) {
          const int s = tile_find_segment(w, v);
          const int j = w.seg_start[s] + (v - w.seg_prefix[s]);
          const double4 pj = a.spos[j];
          cxr[r] = (float)((pj.x + w.seg_sv[s][0]) - o4.x);
          cyr[r] = (float)((pj.y + w.seg_sv[s][1]) - o4.y);
          czr[r] = (float)((pj.z + w.seg_sv[s][2]) - o4.z);
          w.cand_j[r * 32 + lane] = j;
          w.cand_seg[r * 32 + lane] = (unsigned char)s;
        }
      }
      // FP32 prefilter: 8 ballots per atom, lane r keeps word r
      for (int il = 0; il < ni; ++il) {
        const float4 f = w.ri[il];
        unsigned keep = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float dx = cxr[r] - f.x, dy = cyr[r] - f.y, dz = czr[r] - f.z;
          const unsigned bal = __ballot_sync(0xffffffffu, fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= thr2f);
          if (lane == r) keep = bal;
        }
        if (lane < 8) mask[il][lane] = keep;
      }
      __syncwarp();
      // exact phase: TPA lanes per atom walk its survivor bits
      if (have) {
        const int self = ib + i0 + my_il;
        // one flattened walk over this lane's words: every trip of the loop is one survivor, so
        // the lanes of the warp only diverge by their survivor counts
        int wd = my_sub;
        unsigned m = mask[my_il][wd];
        while (true) {
          while (m == 0u && wd + tpa < 8) {
            wd += tpa;
            m = mask[my_il][wd];
          }
          if (m == 0u) break;
          const int idx = wd * 32 + __ffs((int)m) - 1;
          m &= m - 1;
          const int j = w.cand_j[idx];
          const int s = w.cand_seg[idx];
          const double4 pj = a.spos[j];
          const double dx = xadd(xsub(pj.x, pix), w.seg_sv[s][0]);
          const double dy = xadd(xsub(pj.y, piy), w.seg_sv[s][1]);
          const double dz = xadd(xsub(pj.z, piz), w.seg_sv[s][2]);
          const double r2 = xnorm2(dx, dy, dz);
          if (r2 <= rc2 && j != self) {  // vesin_visit.hpp:104, and i != j
            const double invR2 = rcp_fast(r2);
            const double sr2 = a.lj.psi2 * invR2;
            const double a6 = sr2 * sr2 * sr2;
            const double b = 4.0 * a.lj.u0 * a6;
            e_lane += 0.5 * (b * (a6 - 1.0) - a.lj.shiftU);
            const double fs = 6.0 * b * invR2 * (2.0 * a6 - 1.0);
            fx = fma(-fs, dx, fx);
            fy = fma(-fs, dy, fy);
            fz = fma(-fs, dz, fz);
          }
        }
      }
      __syncwarp();
    }
    // fold the TPA lanes of each atom (fixed order) and store
    for (int off = tpa >> 1; off > 0; off >>= 1) {
      fx += __shfl_xor_sync(0xffffffffu, fx, off);
      fy += __shfl_xor_sync(0xffffffffu, fy, off);
      fz += __shfl_xor_sync(0xffffffffu, fz, off);
    }
    if (have && my_sub == 0) {
      F[3 * (size_t)orig_i] = fx;
      F[3 * (size_t)orig_i + 1] = fy;
      F[3 * (size_t)orig_i + 2] = fz;
    }
    __syncwarp();
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) e_lane += __shfl_down_sync(0xffffffffu, e_lane, off);
  if (lane == 0) partials[cell] = e_lane;
}

// ---------------------------------------------------------------------------------------- CuH2
// PASS 0: density + embedding (rgpot_cuh2.f90:436-451) -> dfd_sorted, eemb_sorted
// PASS 1: pair energy + forces (rgpot_cuh2.f90:454-457) -> F, per-cell energy partials
// Same tile walk as k_lj_tile with vesin's accept rule (d2 < rc2, self images dropped); the stencil
// table lists every (cell, image) of the 3x3x3 neighbourhood, so cells that wrap onto themselves
// (fewer than three cells in a direction) are still enumerated image by image.
template <int PASS>
__global__ void __launch_bounds__(kTileWarps * 32)
k_eam_tile(GatherArgs a, float thr2f, double* __restrict__ dfd_sorted, double* __restrict__ eemb_sorted,
           double* __restrict__ F, double* __restrict__ partials, Status* __restrict__ st) {
  __shared__ TileWarp sm[kTileWarps];
  __shared__ unsigned masks[kTileWarps][32][8];
  if (a.st_in->flags & kFlagAtomShift) return;  // per-pair image shifts: the gather kernels take over
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int cell = blockIdx.x * kTileWarps + wid;
  if (cell >= a.grid.ncells) return;
  TileWarp& w = sm[wid];
  unsigned(*mask)[8] = masks[wid];
  const int ib = a.starts[cell], ni_total = a.starts[cell + 1] - ib;
  if (ni_total == 0) {
    if (PASS == 1 && lane == 0) partials[cell] = 0.0;
    return;
  }
  const Grid& g = a.grid;
  const int cx = cell % g.nc[0], cy = (cell / g.nc[0]) % g.nc[1], cz = cell / (g.nc[0] * g.nc[1]);
  const int ncand = tile_build_segments<false>(a, w, cx, cy, cz, lane);
  const double4 o4 = a.spos[ib];
  const double rc2 = a.rc2;
  const double rc2_phys = c_eam.rc2_phys;
  double e_lane = 0.0;
  bool coincident = false;

  for (int i0 = 0; i0 < ni_total; i0 += 32) {
    const int ni = min(32, ni_total - i0);
    const int tpa = ni <= 4 ? 8 : (ni <= 8 ? 4 : (ni <= 16 ? 2 : 1));
    const int my_il = lane / tpa, my_sub = lane % tpa;
    const bool have = my_il < ni;
    double pix = 0.0, piy = 0.0, piz = 0.0, fx = 0.0, fy = 0.0, fz = 0.0, rho = 0.0, dfd_i = 0.0;
    int orig_i = 0, zi = 0;
    if (lane < ni) {
      const double4 p = a.spos[ib + i0 + lane];
      w.ri[lane] = make_float4((float)(p.x - o4.x), (float)(p.y - o4.y), (float)(p.z - o4.z), 0.f);
    }
    if (have) {
      const double4 p = a.spos[ib + i0 + my_il];
      pix = p.x;
      piy = p.y;
      piz = p.z;
      const unsigned long long mi = (unsigned long long)__double_as_longlong(p.w);
      orig_i = meta_orig(mi);
      zi = meta_species(mi);
      if (PASS == 1) dfd_i = dfd_sorted[ib + i0 + my_il];
    }
    __syncwarp();

    for (int ch = 0; ch < ncand; ch += kTileChunk) {
      float cxr[8], cyr[8], czr[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int v = ch + r * 32 + lane;
        cxr[r] = cyr[r] = czr[r] = __int_as_float(0x7fc00000);
        if (v < ncand) {
          const int s = tile_find_segment(w, v);
          const int j = w.seg_start[s] + (v - w.seg_prefix[s]);
          const double4 pj = a.spos[j];
          cxr[r] = (float)((pj.x + w.seg_sv[s][0]) - o4.x);
          cyr[r] = (float)((pj.y + w.seg_sv[s][1]) - o4.y);
          czr[r] = (float)((pj.z + w.seg_sv[s][2]) - o4.z);
          w.cand_j[r * 32 + lane] = j;
          w.cand_seg[r * 32 + lane] = (unsigned char)s;
        }
      }
      for (int il = 0; il < ni; ++il) {
        const float4 f = w.ri[il];
        unsigned keep = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float dx = cxr[r] - f.x, dy = cyr[r] - f.y, dz = czr[r] - f.z;
          const unsigned bal = __ballot_sync(0xffffffffu, fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= thr2f);
          if (lane == r) keep = bal;
        }
        if (lane < 8) mask[il][lane] = keep;
      }
      __syncwarp();
      if (have) {
        const int self = ib + i0 + my_il;
        int wd = my_sub;
        unsigned m = mask[my_il][wd];
        while (true) {
          while (m == 0u && wd + tpa < 8) {
            wd += tpa;
            m = mask[my_il][wd];
          }
          if (m == 0u) break;
          const int idx = wd * 32 + __ffs((int)m) - 1;
          m &= m - 1;
          const int j = w.cand_j[idx];
          const int s = w.cand_seg[idx];
          const double4 pj = a.spos[j];
          const double dx = xadd(xsub(pj.x, pix), w.seg_sv[s][0]);
          const double dy = xadd(xsub(pj.y, piy), w.seg_sv[s][1]);
          const double dz = xadd(xsub(pj.z, piz), w.seg_sv[s][2]);
          const double d2 = xnorm2(dx, dy, dz);
          if (d2 < rc2 && j != self) {  // vesin :1202 ; self images dropped, rgpot_neighbors.f90:132
            const int zj = meta_species((unsigned long long)__double_as_longlong(pj.w));
            if (PASS == 0) {
              if (d2 == 0.0) coincident = true;
              if (d2 < rc2_phys) rho += eam_density_entry(zj, d2);
            } else if (d2 < rc2_phys) {
              eam_force_entry(zi, zj, dfd_i, dfd_sorted[j], dx, dy, dz, d2, e_lane, fx, fy, fz);
            }
          }
        }
      }
      __syncwarp();
    }
    if (PASS == 0) {
      for (int off = tpa >> 1; off > 0; off >>= 1) rho += __shfl_xor_sync(0xffffffffu, rho, off);
      if (have && my_sub == 0) {
        double emb, demb;
        eam_embed(zi, rho, emb, demb);
        dfd_sorted[ib + i0 + my_il] = demb;
        eemb_sorted[ib + i0 + my_il] = emb;
      }
    } else {
      for (int off = tpa >> 1; off > 0; off >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, off);
        fy += __shfl_xor_sync(0xffffffffu, fy, off);
        fz += __shfl_xor_sync(0xffffffffu, fz, off);
      }
      if (have && my_sub == 0) {
        F[3 * (size_t)orig_i] = fx;
        F[3 * (size_t)orig_i + 1] = fy;
        F[3 * (size_t)orig_i + 2] = fz;
        e_lane += eemb_sorted[ib + i0 + my_il];
      }
    }
    __syncwarp();
  }
  if (PASS == 0) {
    if (coincident) atomicOr(&st->flags, kFlagCoincident);
  } else {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e_lane += __shfl_down_sync(0xffffffffu, e_lane, off);
    if (lane == 0) partials[cell] = e_lane;
  }
}

}  // namespace rb
