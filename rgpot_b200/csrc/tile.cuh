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

constexpr int kTileWarps = 4;    // warps per block (static shared memory stays under 48 KB)
constexpr int kTileChunk = 256;  // candidates per chunk (8 per lane)

struct TileWarp {
  double pix[32], piy[32], piz[32];  // the cell's atoms (current pass of <= 32)
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
// survivor mask (8 ballots).  The exact phase cuts the chunk's survivor-bit sequence (atom-major,
// then word, then bit) into 32 ranges of EQUAL LENGTH, one per lane, so every lane runs the same
// number of FP64 pair evaluations whatever the cell's atom count or the survivor distribution.
// A lane redoes each separation in FP64 exactly as the reference does and accumulates in
// registers; the partial sum of every (lane, atom) piece is parked in a shared slot and each
// atom's owner lane adds the pieces that belong to it in lane order.  No queue, no atomics, a
// fixed summation order.
enum { TILE_LJ = 0, TILE_EAM_DENSITY = 1, TILE_EAM_FORCE = 2, TILE_PAIR = 3 };  // TILE_PAIR: Morse / ZBL on the LJ geometry
constexpr int kTilePieces = 4;  // (lane, atom) pieces a lane may produce per chunk

struct TileOut {
  double* F;            // forces, original order (LJ, EAM force)
  double* partials;     // per-cell energy partials (LJ, EAM force)
  double* dfd_sorted;   // EAM: F'(rho), sorted order (written by density, read by force)
  double* eemb_sorted;  // EAM: F(rho)
  Status* st;
};

struct TilePiece {
  double v[3];
  int atom;
  int pad;
};

template <int POT>
__global__ void __launch_bounds__(kTileWarps * 32)
k_tile(GatherArgs a, float thr2f, TileOut o) {
  __shared__ TileWarp sm[kTileWarps];
  __shared__ unsigned masks[kTileWarps][32][8];  // [atom][word]: prefilter survivors
  __shared__ TilePiece pieces[kTileWarps][32][kTilePieces];
  __shared__ int bases[kTileWarps][33];
  if (a.st_in->flags & kFlagAtomShift) return;  // the gather kernels of gather.cuh take over
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int cell = blockIdx.x * kTileWarps + wid;
  if (cell >= a.grid.ncells) return;
  TileWarp& w = sm[wid];
  unsigned(*mask)[8] = masks[wid];
  TilePiece(*piece)[kTilePieces] = pieces[wid];
  int* base = bases[wid];
  const int ib = a.starts[cell], ni_total = a.starts[cell + 1] - ib;
  if (ni_total == 0) {
    if (POT != TILE_EAM_DENSITY && lane == 0) o.partials[cell] = 0.0;
    return;
  }
  const Grid& g = a.grid;
  const int cx = cell % g.nc[0], cy = (cell / g.nc[0]) % g.nc[1], cz = cell / (g.nc[0] * g.nc[1]);
  constexpr bool kPairGeo = POT == TILE_LJ || POT == TILE_PAIR;  // LJ pair search (one image per stencil cell)
  const int ncand = tile_build_segments<kPairGeo>(a, w, cx, cy, cz, lane);
  const double4 o4 = a.spos[ib];  // local origin of the FP32 offsets
  const double rc2 = kPairGeo ? a.lj.rc2 : a.rc2;
  const double rc2_phys = kPairGeo ? 0.0 : c_eam.rc2_phys;
  constexpr int NV = POT == TILE_EAM_DENSITY ? 1 : 3;
  double e_lane = 0.0;
  bool coincident = false;

  for (int i0 = 0; i0 < ni_total; i0 += 32) {
    const int ni = min(32, ni_total - i0);
    double own[3] = {0.0, 0.0, 0.0};  // totals of atom `lane` (owner lanes: lane < ni)
    if (lane < ni) {
      const double4 p = a.spos[ib + i0 + lane];
      w.pix[lane] = p.x;
      w.piy[lane] = p.y;
      w.piz[lane] = p.z;
      w.ri[lane] = make_float4((float)(p.x - o4.x), (float)(p.y - o4.y), (float)(p.z - o4.z), 0.f);
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
#pragma unroll
      for (int k = 0; k < kTilePieces; ++k) piece[lane][k].atom = -1;
      __syncwarp();

      // survivors per atom and their exclusive prefix: lane il owns atom il
      int mine = 0;
      if (lane < ni) {
#pragma unroll
        for (int wd = 0; wd < 8; ++wd) mine += __popc(mask[lane][wd]);
      }
      int incl = mine;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      base[lane] = incl - mine;
      const int B = __shfl_sync(0xffffffffu, incl, 31);
      if (lane == 31) base[32] = B;
      __syncwarp();

      // my range of the survivor sequence
      const int first = (int)(((long long)lane * B) >> 5);
      int todo = (int)(((long long)(lane + 1) * B) >> 5) - first;
      int il = 0, wd = 0, npiece = 0, self = 0, zi = 0;
      unsigned m = 0u;
      double pix = 0.0, piy = 0.0, piz = 0.0, dfd_i = 0.0;
      double acc[3] = {0.0, 0.0, 0.0};
      bool overflow = false;
      if (todo > 0) {
        // atom that holds survivor `first`: last il with base[il] <= first (bases are sorted)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
          if (il + step < ni && base[il + step] <= first) il += step;
        int skip = first - base[il];
        m = mask[il][0];
        while (skip >= __popc(m)) {  // word inside the atom
          skip -= __popc(m);
          m = mask[il][++wd];
        }
        for (; skip > 0; --skip) m &= m - 1;  // bits inside the word
      }
      auto load_atom = [&]() {
        self = ib + i0 + il;
        pix = w.pix[il];
        piy = w.piy[il];
        piz = w.piz[il];
        if (POT == TILE_EAM_FORCE) {
          zi = meta_species((unsigned long long)__double_as_longlong(a.spos[self].w));
          dfd_i = o.dfd_sorted[self];
        } else if (POT == TILE_PAIR) {
          zi = meta_species((unsigned long long)__double_as_longlong(a.spos[self].w));
        }
      };
      bool used = false;  // the current atom received at least one survivor from this lane
      auto park = [&]() {
        if (!used) return;
        used = false;
        if (npiece < kTilePieces) {
#pragma unroll
          for (int k = 0; k < NV; ++k) piece[lane][npiece].v[k] = acc[k];
          piece[lane][npiece].atom = il;
        } else {
          overflow = true;
        }
        ++npiece;
        acc[0] = acc[1] = acc[2] = 0.0;
      };
      if (todo > 0) load_atom();
      for (; todo > 0; --todo) {
        while (m == 0u) {  // next word / next atom that still has survivors
          if (++wd == 8) {
            park();
            wd = 0;
            ++il;
            load_atom();
          }
          m = mask[il][wd];
        }
        used = true;
        const int idx = wd * 32 + __ffs((int)m) - 1;
        m &= m - 1;
        const int j = w.cand_j[idx];
        const int s = w.cand_seg[idx];
        const double4 pj = a.spos[j];
        const double dx = xadd(xsub(pj.x, pix), w.seg_sv[s][0]);
        const double dy = xadd(xsub(pj.y, piy), w.seg_sv[s][1]);
        const double dz = xadd(xsub(pj.z, piz), w.seg_sv[s][2]);
        const double d2 = xnorm2(dx, dy, dz);
        if (POT == TILE_LJ) {
          if (d2 <= rc2 && j != self) {  // vesin_visit.hpp:104, and i != j
            const double invR2 = rcp_fast(d2);
            const double sr2 = a.lj.psi2 * invR2;
            const double a6 = sr2 * sr2 * sr2;
            const double b = 4.0 * a.lj.u0 * a6;
            e_lane += 0.5 * (b * (a6 - 1.0) - a.lj.shiftU);
            const double fs = 6.0 * b * invR2 * (2.0 * a6 - 1.0);
            acc[0] = fma(-fs, dx, acc[0]);
            acc[1] = fma(-fs, dy, acc[1]);
            acc[2] = fma(-fs, dz, acc[2]);
          }
        } else if (POT == TILE_PAIR) {
          if (d2 <= rc2 && j != self) {
            double u, fs;
            pair_eval(a.lj, d2, zi, meta_species((unsigned long long)__double_as_longlong(pj.w)), u, fs);
            e_lane += 0.5 * u;
            acc[0] = fma(-fs, dx, acc[0]);
            acc[1] = fma(-fs, dy, acc[1]);
            acc[2] = fma(-fs, dz, acc[2]);
          }
        } else if (d2 < rc2 && j != self) {  // vesin :1202; self images dropped (rgpot_neighbors.f90:132)
          const int zj = meta_species((unsigned long long)__double_as_longlong(pj.w));
          if (POT == TILE_EAM_DENSITY) {
            if (d2 == 0.0) coincident = true;
            if (d2 < rc2_phys) acc[0] += eam_density_entry(zj, d2);
          } else if (d2 < rc2_phys) {
            eam_force_entry(zi, zj, dfd_i, o.dfd_sorted[j], dx, dy, dz, d2, e_lane, acc[0], acc[1], acc[2]);
          }
        }
      }
      park();
      // A lane that would need more pieces than it has slots (many atoms with very few survivors)
      // cannot happen for condensed systems; it is refused loudly rather than summed wrongly.
      if (__any_sync(0xffffffffu, overflow)) {
        if (lane == 0) atomicOr(&o.st->flags, kFlagListOverflow);
      }
      __syncwarp();
      // owner lanes add the pieces of their atom in lane order (fixed)
      if (lane < ni && mine > 0) {
        const int b0 = base[lane], b1 = b0 + mine - 1;
        auto range_start = [&](int l) { return (int)(((long long)l * B) >> 5); };
        int L = (int)(((long long)b0 << 5) / B);  // lane whose range holds survivor b0
        while (L < 31 && range_start(L + 1) <= b0) ++L;
        while (L > 0 && range_start(L) > b0) --L;
        for (; L < 32 && range_start(L) <= b1; ++L) {
#pragma unroll
          for (int k2 = 0; k2 < kTilePieces; ++k2) {
            if (piece[L][k2].atom == lane) {
#pragma unroll
              for (int k = 0; k < NV; ++k) own[k] += piece[L][k2].v[k];
            }
          }
        }
      }
      __syncwarp();
    }
    // results of this pass of <= 32 atoms
    if (lane < ni) {
      const int self = ib + i0 + lane;
      const unsigned long long mi = (unsigned long long)__double_as_longlong(a.spos[self].w);
      if (POT == TILE_EAM_DENSITY) {
        double emb, demb;
        eam_embed(meta_species(mi), own[0], emb, demb);
        o.dfd_sorted[self] = demb;
        o.eemb_sorted[self] = emb;
      } else {
        const int orig = meta_orig(mi);
        o.F[3 * (size_t)orig] = own[0];
        o.F[3 * (size_t)orig + 1] = own[1];
        o.F[3 * (size_t)orig + 2] = own[2];
        if (POT == TILE_EAM_FORCE) e_lane += o.eemb_sorted[self];
      }
    }
    __syncwarp();
  }
  if (POT == TILE_EAM_DENSITY) {
    if (coincident) atomicOr(&o.st->flags, kFlagCoincident);
  } else {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e_lane += __shfl_down_sync(0xffffffffu, e_lane, off);
    if (lane == 0) o.partials[cell] = e_lane;
  }
}

}  // namespace rb
