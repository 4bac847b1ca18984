// tile.cu -- the large-system kernels (>= 20 000 atoms): cell groups on tensor-core tiles.
//
// What it replaces under the reference path: the pair search (vesin's cell list,
// third_party/vesin/vesin-single-build.cpp:1444-1519, and the O(N^2) scan of
// CppCore/rgpot/nlist/vesin_visit.hpp:82-110) and the per-pair physics of LJPot.cc:61-81 /
// rgpot_cuh2.f90:506-587 that consumes its output.
//
// One thread block owns a BRICK of cells and stages the brick's halo in shared memory once (FP64
// positions for the physics, a 16-byte FP16 record for the search).  A WARP owns a GROUP of adjacent
// cells -- about 14 atoms; groups are aligned to the GRID, not to the brick, so an atom's group, its
// candidate window and every summation order below depend on the cell list alone, never on how the
// planner sized bricks or buffers -- and evaluates its atoms sixteen at a time:
//
//   window   the group's stencil box, (gd + 2 ns)^3 cells, flattened into one list of staged indices
//            in (z, y, x, rank) order: every candidate any atom of the group can reach, ~365 for an
//            LJ fluid at rc = 2.5 sigma.
//   search   |r_i - r_j|^2 = n_i + n_j - 2 r_i.r_j is a Gram matrix: with A = (x, y, z, 1, 1, n_hi,
//            n_lo, 0) per atom and B = (-2x, -2y, -2z, n_hi, n_lo, 1, 1, 0) per candidate one
//            mma.sync.m16n8k8 (FP16 in, FP32 accumulate) yields the 16 x 8 squared distances of
//            sixteen atoms against eight candidates -- 128 pair tests per warp instruction where the
//            scalar search spent fourteen instructions on 32.  Every lane compares its four results
//            with the threshold and keeps four bits; positions are FP16 offsets from the brick centre
//            and the threshold carries their rounding bound, so the survivors are a SUPERSET of the
//            exact accept set.  All 32 lanes run the same trip count on the same candidates: no
//            divergence, no per-thread lists, no shared-memory atomics.
//   physics  lane (g, t) walks its bits for atom g, then for atom g + 8 (the fragment layout makes a
//            quad of lanes own an atom's row): the separation is redone in FP64 with the REFERENCE's
//            expression ((p_j - p_i) + S.H, unfused norm), so accept / reject is bit-exact, and the
//            pair terms are accumulated in registers; the four partial sums of a quad are folded by
//            two shuffles.  Fixed order, no floating-point atomics: bit-reproducible.
//
// The FP64 pair physics stays on the CUDA cores (it is gather/pairwise work); only the all-pairs
// distance prefilter, which IS a small dense contraction, uses the tensor cores.
//
// EAM: the density pass hands its survivor masks (32 bits per lane and 64 candidates) to the force
// pass through a global pool, so the search runs once per call.
#include <cuda_fp16.h>

#include "tile_api.hpp"

namespace rb {

namespace {

// ---- small helpers -----------------------------------------------------------------------------
RB_D unsigned h2_as_u32(__half2 v) { return *reinterpret_cast<unsigned*>(&v); }

// D(16x8, f32) = A(16x8, f16, row) * B(8x8, f16, col), fresh accumulator
RB_D void mma_16x8x8(float c[4], unsigned a0, unsigned a1, unsigned b) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %7, %7, %7};\n"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a0), "r"(a1), "r"(b), "f"(0.0f));
}

// The four FP16 pairs of a candidate's B column: k = (-2x, -2y | -2z, n_hi | n_lo, 1 | 1, 0).
// Positions are offsets from the box centre, rounded to FP16 ONCE; the norm is taken of the rounded
// values in FP32 and split into two halves, so the tensor-core d2 is the distance between the
// rounded points up to FP32 rounding of the accumulation.
RB_D uint4 make_record(float fx, float fy, float fz) {
  const __half hx = __float2half_rn(fx), hy = __float2half_rn(fy), hz = __float2half_rn(fz);
  const float rx = __half2float(hx), ry = __half2float(hy), rz = __half2float(hz);
  const float n = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
  const __half nhi = __float2half_rn(n);
  const __half nlo = __float2half_rn(n - __half2float(nhi));
  const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
  uint4 r;
  r.x = h2_as_u32(__halves2half2(__float2half_rn(-2.0f * rx), __float2half_rn(-2.0f * ry)));
  r.y = h2_as_u32(__halves2half2(__float2half_rn(-2.0f * rz), nhi));
  r.z = h2_as_u32(__halves2half2(nlo, one));
  r.w = h2_as_u32(__halves2half2(one, zero));
  return r;
}

// A-row pair k = 2t, 2t + 1 of an atom from its B record: (x, y | z, 1 | 1, n_hi | n_lo, 0)
RB_D unsigned a_pair_from_record(const uint4& r, int t) {
  const __half2 mhalf = __floats2half2_rn(-0.5f, -0.5f);
  const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
  const __half2 xy = __hmul2(*reinterpret_cast<const __half2*>(&r.x), mhalf);  // exact: a power of two
  const __half2 zn = *reinterpret_cast<const __half2*>(&r.y);                  // (-2z, n_hi)
  const __half2 l1 = *reinterpret_cast<const __half2*>(&r.z);                  // (n_lo, 1)
  if (t == 0) return h2_as_u32(xy);
  if (t == 1) return h2_as_u32(__halves2half2(__hmul(__low2half(zn), __low2half(mhalf)), one));
  if (t == 2) return h2_as_u32(__halves2half2(one, __high2half(zn)));
  return h2_as_u32(__halves2half2(__low2half(l1), zero));
}
// a row that matches nothing: n_hi = 60000 puts every d2 far above any threshold
RB_D unsigned a_pair_empty(int t) {
  const __half zero = __float2half_rn(0.0f);
  if (t == 2) return h2_as_u32(__halves2half2(zero, __float2half_rn(60000.0f)));
  return h2_as_u32(__halves2half2(zero, zero));
}

constexpr unsigned kRowEmpty = 0xffffu;

// ---- per-atom walk straight from global memory (the in-kernel slow path) --------------------------
// For a single group whose window does not fit the staged capacity.  It visits the group's window in
// the kernel's order -- rows (z, y), cells along x, atoms in sorted order -- assigns window position w
// to partial sum (w % 8) / 4 like the kernel's two lanes per atom do, and adds the two partial sums the
// way the kernel's shuffle does: an atom gets the same bits whichever path evaluated it.
template <int POT>
__device__ __noinline__ void tile_slow_atom(const GatherArgs& a, const TilePlan& tp, const TileOut& o,
                                            const double (*svtab)[3], bool ashift, const int g0[3], const int gw[3],
                                            int i, bool& coincident) {
  constexpr bool kPairGeo = POT == TILE_LJ || POT == TILE_PAIR;
  constexpr bool kForce = tile_is_force(POT);
  constexpr bool kDensity = tile_is_density(POT);
  constexpr bool kFeHe = tile_is_fehe(POT);
  constexpr int NV = kDensity ? (kFeHe ? 2 : 1) : 3;
  const Grid& g = a.grid;
  const double4 pi = a.spos[i];
  const unsigned long long mi = (unsigned long long)__double_as_longlong(pi.w);
  const int zi = meta_species(mi);
  const double rc2 = kPairGeo ? a.lj.rc2 : a.rc2;
  const double rc2_phys = (kPairGeo || kFeHe) ? 0.0 : eam_par<tile_model(POT)>().rc2_phys;
  double dfd_i = 0.0, dfs_i = 0.0;
  if (kForce) dfd_i = o.dfd_sorted[i];
  if (kForce && kFeHe) dfs_i = o.dfs_sorted[i];
  double part[2][4];
  int npair[2] = {0, 0};
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int k = 0; k < 4; ++k) part[s][k] = 0.0;
  int wpos = 0;
  const int wz = gw[2] + 2 * tp.ns[2], wy = gw[1] + 2 * tp.ns[1], wx = gw[0] + 2 * tp.ns[0];
  for (int hz = 0; hz < wz; ++hz)
    for (int hy = 0; hy < wy; ++hy)
      for (int hx = 0; hx < wx; ++hx) {
        const int hc[3] = {hx, hy, hz};
        int q[3], r[3];
        bool valid = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          q[k] = 0;
          r[k] = g0[k] - tp.ns[k] + hc[k];
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
        for (int j = jb; j < je; ++j, ++wpos) {
          const double4 pj = a.spos[j];
          const unsigned long long mj = (unsigned long long)__double_as_longlong(pj.w);
          double svx = svtab[code][0], svy = svtab[code][1], svz = svtab[code][2];
          if (ashift) {  // image = cell image + s_i - s_j (vesin :1497)
            double sv[3];
            image_vector<kPairGeo>(a, q[0] + meta_shift(mi, 0) - meta_shift(mj, 0), q[1] + meta_shift(mi, 1) - meta_shift(mj, 1),
                                   q[2] + meta_shift(mi, 2) - meta_shift(mj, 2), sv);
            svx = sv[0];
            svy = sv[1];
            svz = sv[2];
          }
          const double dx = xadd(xsub(pj.x, pi.x), svx), dy = xadd(xsub(pj.y, pi.y), svy),
                       dz = xadd(xsub(pj.z, pi.z), svz);
          const double d2 = xnorm2(dx, dy, dz);
          const int half = (wpos & 7) >> 2;
          double* ps = half ? part[1] : part[0];
          int& np = half ? npair[1] : npair[0];
          tile_pair<POT>(a, rc2, rc2_phys, j != i || code != kShiftCodes / 2, zi, meta_species(mj), dfd_i,
                         [&]() { return o.dfd_sorted[j]; }, dfs_i, [&]() { return o.dfs_sorted[j]; }, dx, dy, dz, d2,
                         ps[3], ps, np, coincident);
        }
      }
#pragma unroll
  for (int k = 0; k < 4; ++k) part[0][k] = part[0][k] + part[1][k];
  tile_finish<POT>(a, part[0], part[0][3], npair[0] + npair[1]);
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

// ---- the kernel ----------------------------------------------------------------------------------
// resident blocks per SM the register budget aims at: the EAM force pass carries the most live state
constexpr int tile_min_blocks(int pot) { return tile_is_force(pot) ? 2 : 3; }

template <int POT>
__global__ void __launch_bounds__(kTileThreads, tile_min_blocks(POT))
k_tile(const __grid_constant__ GatherArgs a, const __grid_constant__ TilePlan tp, const __grid_constant__ TileOut o) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned short hstart[kTileMaxHalo + 2];  // staged start of every halo cell
  __shared__ int hgstart[kTileMaxHalo];                // start of the halo cell in the sorted atom array
  __shared__ unsigned char hcode[kTileMaxHalo];        // image code of the halo cell
  __shared__ double svtab[kShiftCodes][3];             // image vector per code (reference arithmetic)
  __shared__ int s_shifted, s_ntot;
  __shared__ int s_scan[kTileThreads / 32];
  __shared__ int box_stack[12][6];  // boxes of cells still to do: first cell, extent

  if (a.st_in->flags & kFlagBigShift) return;  // the gather kernels of gather.cuh take over
  const bool ashift = (a.st_in->flags & kFlagAtomShift) != 0;
  if (tile_is_density(POT) || tile_is_force(POT)) exp_table_init();
  constexpr bool kPairGeo = POT == TILE_LJ || POT == TILE_PAIR;  // LJ pair search (w * q image vectors)
  constexpr bool kForce = tile_is_force(POT);
  constexpr bool kDensity = tile_is_density(POT);
  constexpr int kModel = tile_model(POT);
  constexpr bool kFeHe = tile_is_fehe(POT);
  constexpr int NV = kDensity ? (kFeHe ? 2 : 1) : 3;
  constexpr int kWords = tile_words(POT);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int fg = lane >> 2, ft = lane & 3;  // fragment row (atom slot) and column pair of this lane
  const Grid& g = a.grid;

  double* c64 = reinterpret_cast<double*>(smem_raw);  // x, y, z of staged atom s at c64[3 s ..]
  double* c64w = c64 + 3 * tp.cap_cand;               // EAM force pass: F'(rho_j)
  double* c64w2 = c64 + 4 * tp.cap_cand;              // Fe-He force pass: F'(rho_j) of the s-band
  uint4* rec = reinterpret_cast<uint4*>(c64 + kWords * tp.cap_cand);  // cap_cand + 1 records
  unsigned short* cinfo = reinterpret_cast<unsigned short*>(rec + tp.cap_cand + 1);  // image code | species << 8
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(cinfo + ((tp.cap_cand + 8) & ~7));
  const unsigned per_warp = 2u * (unsigned)tp.wcap + 4u * 32u * ((unsigned)tp.wcap / 64u) + 128u;
  unsigned short* widx = reinterpret_cast<unsigned short*>(warp_base + (size_t)wid * per_warp);
  unsigned* masks = reinterpret_cast<unsigned*>(widx + tp.wcap);      // word w of lane l at masks[32 w + l]
  int* rowgidx = reinterpret_cast<int*>(masks + 32 * (tp.wcap / 64));  // 16 atom slots: index in the sorted array
  unsigned short* rowsidx = reinterpret_cast<unsigned short*>(rowgidx + 16);  // ... and staged index
  const unsigned* rec32 = reinterpret_cast<const unsigned*>(rec);

  for (int t = tid; t < kShiftCodes; t += nthreads) image_vector<kPairGeo>(a, t % 5 - 2, (t / 5) % 5 - 2, t / 25 - 2, svtab[t]);
  if (tid == 0) {
    const int b[3] = {(int)blockIdx.x % tp.nb[0], ((int)blockIdx.x / tp.nb[0]) % tp.nb[1],
                      (int)blockIdx.x / (tp.nb[0] * tp.nb[1])};
    for (int k = 0; k < 3; ++k) {
      box_stack[0][k] = b[k] * tp.bd[k];
      box_stack[0][3 + k] = min(tp.bd[k], g.nc[k] - b[k] * tp.bd[k]);
    }
    // the far-away dummy record: pads windows to a multiple of eight, matches nothing
    rec[tp.cap_cand] = make_uint4(0u, h2_as_u32(__halves2half2(__float2half_rn(0.0f), __float2half_rn(60000.0f))),
                                  h2_as_u32(__halves2half2(__float2half_rn(0.0f), __float2half_rn(1.0f))),
                                  h2_as_u32(__halves2half2(__float2half_rn(1.0f), __float2half_rn(0.0f))));
  }
  bool coincident = false;
  const double rc2 = kPairGeo ? a.lj.rc2 : a.rc2;
  const double rc2_phys = (kPairGeo || kFeHe) ? 0.0 : eam_par<kModel>().rc2_phys;

  int sp = 1;
  __syncthreads();
  for (;;) {
    int c0[3], bw[3], hd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      c0[k] = box_stack[sp - 1][k];
      bw[k] = box_stack[sp - 1][3 + k];
      hd[k] = bw[k] + 2 * tp.ns[k];
    }
    const int nh = hd[0] * hd[1] * hd[2];
    if (tid == 0) s_shifted = 0;
    __syncthreads();

    // ---- halo cells: where each lives in the sorted atom array, its image code, its population -------
    for (int h = tid; h < nh; h += nthreads) {
      const int hyz = h / hd[0], hz = hyz / hd[1];
      const int hc[3] = {h - hyz * hd[0], hyz - hz * hd[1], hz};
      int q[3], r[3];
      bool valid = true;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        q[k] = 0;
        r[k] = c0[k] - tp.ns[k] + hc[k];
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
    {  // exclusive prefix over nh <= 1024 counts: a run of cells per thread, block scan of the run sums
      const int per = (nh + nthreads - 1) / nthreads;
      const int h0 = tid * per, h1 = min(nh, h0 + per);
      int sum = 0;
      for (int h = h0; h < h1; ++h) sum += hstart[h + 1];
      int incl = sum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      if (lane == 31) s_scan[wid] = incl;
      __syncthreads();
      int base = 0;
      for (int w = 0; w < wid; ++w) base += s_scan[w];
      int run = base + incl - sum;
      if (tid == 0) hstart[0] = 0;
      for (int h = h0; h < h1; ++h) {
        run += hstart[h + 1];
        hstart[h + 1] = (unsigned short)min(run, 65535);
      }
      if (tid == nthreads - 1) {
        s_ntot = base + incl;
        if (sp == 1) atomicMax(&o.st->max_halo, base + incl);  // whole bricks only: sizes the next call
      }
    }
    __syncthreads();
    const int ntot = s_ntot;
    if (ntot > tp.cap_cand) {
      // too crowded for the staged capacity: cut the box in two along its longest direction, on a
      // group boundary (groups stay whole: their windows and summation orders do not change)
      const int ngk[3] = {(bw[0] + tp.gd[0] - 1) / tp.gd[0], (bw[1] + tp.gd[1] - 1) / tp.gd[1],
                          (bw[2] + tp.gd[2] - 1) / tp.gd[2]};
      int kmax = -1, best = 0;
      if (ngk[2] >= 2 && bw[2] >= best) { best = bw[2]; kmax = 2; }
      if (ngk[1] >= 2 && bw[1] >= best) { best = bw[1]; kmax = 1; }
      if (ngk[0] >= 2 && bw[0] >= best) { best = bw[0]; kmax = 0; }
      if (kmax >= 0) {
        if (tid == 0) {
          const int ngm = kmax == 0 ? ngk[0] : (kmax == 1 ? ngk[1] : ngk[2]);
          const int gdm = kmax == 0 ? tp.gd[0] : (kmax == 1 ? tp.gd[1] : tp.gd[2]);
          const int half = (ngm / 2) * gdm;  // cells in the first part
#pragma unroll
          for (int part = 0; part < 2; ++part) {
#pragma unroll
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
      // a single group whose window still does not fit: per-atom walk straight from global memory
      const int gw[3] = {bw[0], bw[1], bw[2]};
      int cnt = 0;
      for (int cz = 0; cz < bw[2]; ++cz)
        for (int cy = 0; cy < bw[1]; ++cy) {
          const int cell = ((c0[2] + cz) * g.nc[1] + c0[1] + cy) * g.nc[0] + c0[0];
          const int jb = a.starts[cell], je = a.starts[cell + bw[0]];  // cells along x are contiguous
          for (int j = jb + tid; j < je; j += nthreads) tile_slow_atom<POT>(a, tp, o, svtab, ashift, c0, gw, j, coincident);
          cnt += je - jb;
        }
      if (tid == 0 && cnt) atomicAdd(&o.st->slow_atoms, cnt);
      if (--sp == 0) break;
      __syncthreads();
      continue;
    }

    // box centre for the FP16 offsets (any nearby point serves; it only has to be common)
    double org[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double fr = ((double)c0[d] + 0.5 * (double)bw[d]) / (double)g.nc[d];
      if (a.box.periodic[d]) {
#pragma unroll
        for (int k = 0; k < 3; ++k) org[k] += fr * a.box.H[3 * d + k];
      } else {
        org[d] += a.box.lo[d] + a.box.ext[d] * fr;
      }
    }
    // ---- stage the halo: every staged slot learns its halo cell (a thread per cell), then one staged
    // atom per thread and step ------------------------------------------------------------------------
    for (int h = tid; h < nh; h += nthreads) {
      const int sb = hstart[h], se = hstart[h + 1];
      for (int s = sb; s < se; ++s) cinfo[s] = (unsigned short)h;
    }
    __syncthreads();
    for (int s = tid; s < ntot; s += nthreads) {
      const int h = cinfo[s];
      const int j = hgstart[h] + (s - hstart[h]);
      const double4 p = a.spos[j];
      const unsigned long long mj = (unsigned long long)__double_as_longlong(p.w);
      int cj = hcode[h];
      if (ashift) {
        const int wo = wrap_offset(mj);
        if (wo != 0) {  // vesin :1497: shift = cell shift + s_i - s_j
          cj -= wo;
          s_shifted = 1;
        }
      }
      c64[3 * s] = p.x;
      c64[3 * s + 1] = p.y;
      c64[3 * s + 2] = p.z;
      if (kForce) c64w[s] = o.dfd_sorted[j];
      if (kForce && kFeHe) c64w2[s] = o.dfs_sorted[j];
      rec[s] = make_record((float)(p.x + (svtab[cj][0] - org[0])), (float)(p.y + (svtab[cj][1] - org[1])),
                           (float)(p.z + (svtab[cj][2] - org[2])));
      cinfo[s] = (unsigned short)(cj | (meta_species(mj) << 8));
    }
    __syncthreads();
    const bool shifted = s_shifted != 0;

    // ---- groups of this box, one per warp and round --------------------------------------------------
    const int ng[3] = {(bw[0] + tp.gd[0] - 1) / tp.gd[0], (bw[1] + tp.gd[1] - 1) / tp.gd[1],
                       (bw[2] + tp.gd[2] - 1) / tp.gd[2]};
    const int ngroups = ng[0] * ng[1] * ng[2];
    for (int grp = wid; grp < ngroups; grp += nwarps) {
      const int gyz = grp / ng[0];
      const int gc[3] = {grp - gyz * ng[0], gyz % ng[1], gyz / ng[1]};
      int g0[3], gw[3];  // first cell (box-local) and extent of the group
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        g0[k] = gc[k] * tp.gd[k];
        gw[k] = min(tp.gd[k], bw[k] - g0[k]);
      }
      // atoms of the group: cells in (z, y, x) order, one cell per lane
      const int ncell = gw[0] * gw[1] * gw[2];
      int cs = 0, cn = 0, cg = 0;  // staged start, population and sorted-array start of this lane's cell
      if (lane < ncell) {
        const int cyz = lane / gw[0];
        const int lx = lane - cyz * gw[0], ly = cyz % gw[1], lz = cyz / gw[1];
        const int hc = ((g0[2] + lz + tp.ns[2]) * hd[1] + (g0[1] + ly + tp.ns[1])) * hd[0] + g0[0] + lx + tp.ns[0];
        cs = hstart[hc];
        cn = (int)hstart[hc + 1] - cs;
        cg = hgstart[hc];
      }
      int cincl = cn;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, cincl, off);
        if (lane >= off) cincl += t;
      }
      const int cexcl = cincl - cn;
      const int natoms = __shfl_sync(0xffffffffu, cincl, 31);
      if (natoms == 0) continue;
      // window: rows (z, y) of gw[0] + 2 ns[0] cells that are contiguous in the staged order
      const int wx = gw[0] + 2 * tp.ns[0], wy = gw[1] + 2 * tp.ns[1], wz = gw[2] + 2 * tp.ns[2];
      const int nrows = wy * wz;  // <= 64: two rows per lane
      int rs[2], rn[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int row = 2 * lane + t;
        rs[t] = rn[t] = 0;
        if (row < nrows) {
          const int ry = row % wy, rz = row / wy;
          const int h0 = ((g0[2] + rz) * hd[1] + g0[1] + ry) * hd[0] + g0[0];
          rs[t] = hstart[h0];
          rn[t] = (int)hstart[h0 + wx] - rs[t];
        }
      }
      int rincl = rn[0] + rn[1];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, rincl, off);
        if (lane >= off) rincl += t;
      }
      const int roff0 = rincl - rn[0] - rn[1], roff1 = roff0 + rn[0];  // window offsets of this lane's rows
      const int wtot = __shfl_sync(0xffffffffu, rincl, 31);

      for (int a0 = 0; a0 < natoms; a0 += 16) {  // sixteen atoms per pass
        // the pass's atoms into the slot tables: staged index and index in the sorted atom array
        __syncwarp();
        if (lane < 16) rowsidx[lane] = (unsigned short)kRowEmpty;
        __syncwarp();
        {
          const int k0 = max(0, a0 - cexcl), k1 = min(cn, a0 + 16 - cexcl);
          for (int k = k0; k < k1; ++k) {
            rowsidx[cexcl + k - a0] = (unsigned short)(cs + k);
            rowgidx[cexcl + k - a0] = cg + k;
          }
        }
        __syncwarp();
        const unsigned si0 = rowsidx[fg], si1 = rowsidx[fg + 8];  // staged index of the atoms of slot fg, fg + 8
        unsigned fa0 = a_pair_empty(ft), fa1 = a_pair_empty(ft);
        if (si0 != kRowEmpty) fa0 = a_pair_from_record(rec[si0], ft);
        if (si1 != kRowEmpty) fa1 = a_pair_from_record(rec[si1], ft);
        const int gfirst = rowgidx[0];  // sorted index of the pass's first atom: the key of the hand-off
        // the masks the density pass left for this pass, if it did
        bool have = false;
        unsigned long long lref_slot = 0;
        if (kForce && o.poolw != nullptr && wtot <= tp.wcap) {
          lref_slot = o.lref[gfirst];
          have = lref_slot != kNoList;
        }

        // physics mapping: TWO lanes per atom.  Lane L works for the atom of slot L / 2 on half L % 2 of
        // every tile's eight candidates (columns 4 h .. 4 h + 3): one stream of ~28 survivors per lane.
        // Its bits sit in the mask words of the two search lanes that own those columns of the slot's
        // row: lanes 4 (slot % 8) + 2 h and + 2 h + 1, top half of each nibble for slots 0..7, bottom
        // half for slots 8..15.
        const int pslot = lane >> 1, phalf = lane & 1, pR = pslot >> 3;
        const unsigned psi = rowsidx[pslot];
        const unsigned* msrc = masks + 4 * (pslot & 7) + 2 * phalf;
        double ac[3] = {0.0, 0.0, 0.0}, ec = 0.0;
        int npair = 0;
        double px = 0.0, py = 0.0, pz = 0.0, dfd_i = 0.0, dfs_i = 0.0;
        int zi = 0, wi = 0;
        if (psi != kRowEmpty) {
          px = c64[3 * psi];
          py = c64[3 * psi + 1];
          pz = c64[3 * psi + 2];
          const int info_i = cinfo[psi];
          zi = info_i >> 8;
          if (kForce) dfd_i = c64w[psi];
          if (kForce && kFeHe) dfs_i = c64w2[psi];
          // own wrap counts (rare): the image vector is then worked out pair by pair; the staged image
          // code of an interior atom is centre - wrap
          if (ashift) wi = kShiftCodes / 2 - (info_i & 127);
        }

        for (int w0 = 0; w0 < wtot; w0 += tp.wcap) {  // window chunks (one, unless the region is crowded)
          const int wlen = min(tp.wcap, wtot - w0);
          const int wpad = (wlen + 7) & ~7;
          __syncwarp();
          // flatten this chunk of the window: every lane copies its own two rows
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int ro = t ? roff1 : roff0;
            const int k0 = max(0, w0 - ro), k1 = min(rn[t], w0 + wlen - ro);
            for (int k = k0; k < k1; ++k) widx[ro + k - w0] = (unsigned short)(rs[t] + k);
          }
          if (lane < wpad - wlen) widx[wlen + lane] = (unsigned short)tp.cap_cand;
          __syncwarp();
          const int ntile = wpad >> 3, nword = (ntile + 7) >> 3;
          if (have) {
            const unsigned* src = o.poolw + (lref_slot & ((1ull << 40) - 1ull)) + lane;
            for (int wd = 0; wd < nword; ++wd) masks[32 * wd + lane] = __ldg(src + 32 * wd);
          } else {
            // ---- search: 16 atoms x 8 candidates per tensor-core instruction ------------------------------
            // the four results of a lane are rounded to FP16 (the threshold carries that rounding too),
            // compared two at a time, and the four outcomes land as one nibble of the lane's mask word;
            // bit 31 - (4 u + i) <-> result i of tile u, so "most significant bit first" walks a word in
            // window order
            const __half2 thr2h = __half2half2(__float2half_ru(tp.thr2 * 1.002f));
            for (int wd = 0; wd < nword; ++wd) {
              unsigned m = 0;
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int t = 8 * wd + u;
                if (t < ntile) {
                  const unsigned idx = widx[8 * t + fg];
                  const unsigned bfrag = rec32[4 * idx + ft];
                  float c[4];
                  mma_16x8x8(c, fa0, fa1, bfrag);
                  const unsigned r01 = __hle2_mask(__floats2half2_rn(c[0], c[1]), thr2h);
                  const unsigned r23 = __hle2_mask(__floats2half2_rn(c[2], c[3]), thr2h);
                  unsigned v = (r01 & 0x00040008u) | (r23 & 0x00010002u);
                  v |= v >> 16;
                  asm("bfi.b32 %0, %1, %0, %2, 4;" : "+r"(m) : "r"(v), "r"(28 - 4 * u));
                }
              }
              masks[32 * wd + lane] = m;
            }
            if (kDensity && o.poolw != nullptr) {
              // hand the masks to the force pass (single-chunk windows only)
              unsigned long long ref = kNoList;
              if (wtot <= tp.wcap) {
                unsigned long long chunk = 0;
                if (lane == 0) chunk = atomicAdd(&o.st->pool_used, (unsigned long long)nword * 32ull);
                chunk = __shfl_sync(0xffffffffu, chunk, 0);
                if (chunk + (unsigned long long)nword * 32ull <= o.pool_cap) {
                  unsigned* dst = o.poolw + chunk + lane;
                  for (int wd = 0; wd < nword; ++wd) dst[32 * wd] = masks[32 * wd + lane];
                  ref = chunk;
                }
              }
              if (lane == 0) o.lref[gfirst] = ref;
            }
          }
          // ---- physics: exact FP64 test and pair terms for the kept candidates -------------------------
          // A lane consumes one survivor per iteration; when its word runs dry it builds the next one in
          // the same iteration (no inner scan loop), and the warp votes on leaving: the lanes stay in
          // step and only wait for each other at the end.
          __syncwarp();  // (the masks were written by other lanes)
          auto physics = [&](auto shifted_tag, auto ashift_tag) {
            constexpr bool SHIFTED = decltype(shifted_tag)::value;
            constexpr bool ASHIFT = decltype(ashift_tag)::value;
            constexpr bool kNeedInfo = SHIFTED || POT != TILE_LJ;
            // window position of bit k (counted from the top) of the combined word of word index wd:
            //   64 wd + 8 (k >> 2) + 4 h + (k & 3) = wq[2 k - (k & 3)] with wq advanced by 64 per word
            const unsigned short* wq = widx + 4 * phalf;
            const unsigned* mp = msrc;
            int wleft = psi != kRowEmpty ? nword : 0;  // words not yet loaded
            unsigned m = 0;
            auto next_word = [&]() {
              const unsigned ma = mp[0], mb = mp[1];
              m = pR ? (((ma & 0x33333333u) << 2) | (mb & 0x33333333u)) : ((ma & 0xccccccccu) | ((mb & 0xccccccccu) >> 2));
              mp += 32;
              --wleft;
            };
            if (wleft > 0) next_word();
            for (;;) {
              if (m == 0 && wleft > 0) {  // this word is used up: take the next one
                next_word();
                wq += 64;
              }
              if (!__any_sync(0xffffffffu, m != 0 || wleft > 0)) break;
              if (m != 0) {
                const int k = __clz((int)m);
                m &= ~(0x80000000u >> k);
                const int sj = wq[2 * k - (k & 3)];
                int info = 0;
                if (kNeedInfo) info = cinfo[sj];
                const double* pj = c64 + 3 * sj;
                double dx = xsub(pj[0], px), dy = xsub(pj[1], py), dz = xsub(pj[2], pz);
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
                tile_pair<POT>(a, rc2, rc2_phys, (unsigned)sj != psi, zi, info >> 8, dfd_i, [&]() { return c64w[sj]; }, dfs_i,
                               [&]() { return c64w2[sj]; }, dx, dy, dz, d2, ec, ac, npair, coincident);
              }
            }
          };
          if (!shifted)
            physics(BoolTag<false>{}, BoolTag<false>{});
          else if (!ashift)
            physics(BoolTag<true>{}, BoolTag<false>{});
          else
            physics(BoolTag<true>{}, BoolTag<true>{});
        }
        // the two lanes of an atom add their partial sums
#pragma unroll
        for (int k = 0; k < NV; ++k) ac[k] += __shfl_xor_sync(0xffffffffu, ac[k], 1);
        if (!kDensity) ec += __shfl_xor_sync(0xffffffffu, ec, 1);
        if (POT == TILE_LJ) npair += __shfl_xor_sync(0xffffffffu, npair, 1);
        if (phalf == 0 && psi != kRowEmpty) {
          const int gidx = rowgidx[pslot];
          if (kDensity) {
            double emb, demb;
            if (kFeHe) {
              double dems;
              fehe_embed(zi, ac[0], ac[1], emb, demb, dems);
              o.dfs_sorted[gidx] = dems;
            } else {
              eam_embed<kModel>(zi, ac[0], emb, demb);
            }
            o.dfd_sorted[gidx] = demb;
            o.eemb_sorted[gidx] = emb;
          } else {
            tile_finish<POT>(a, ac, ec, npair);
            const int orig = meta_orig((unsigned long long)__double_as_longlong(a.spos[gidx].w));
            o.F[3 * (size_t)orig] = ac[0];
            o.F[3 * (size_t)orig + 1] = ac[1];
            o.F[3 * (size_t)orig + 2] = ac[2];
            if (kForce) ec += o.eemb_sorted[gidx];
            o.eatom[gidx] = ec;
          }
        }
      }
    }
    if (--sp == 0) break;
    __syncthreads();
  }
  if (kDensity && coincident) atomicOr(&o.st->flags, kFlagCoincident);
}

template <int POT>
cudaError_t launch_one(const GatherArgs& ga, const TilePlan& tp, const TileOut& to, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_tile<POT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp.smem);
  if (e != cudaSuccess) return e;
  k_tile<POT><<<tp.nbricks, tp.threads, tp.smem, st>>>(ga, tp, to);
  return cudaGetLastError();
}

}  // namespace

cudaError_t tile_upload_constants(int kind, const void* params, size_t bytes, const double* exp2tab) {
  cudaError_t e = cudaSuccess;
  if (exp2tab) e = cudaMemcpyToSymbol(c_exp2tab, exp2tab, sizeof(double) * 64);
  if (exp2tab && e == cudaSuccess) e = cudaMemcpyToSymbol(g_exp2tab, exp2tab, sizeof(double) * 64);
  if (e != cudaSuccess) return e;
  if (kind == 0 && bytes == sizeof(EamParams)) return cudaMemcpyToSymbol(c_eam, params, bytes);
  if (kind == 1 && bytes == sizeof(EamParams)) return cudaMemcpyToSymbol(c_eam_al, params, bytes);
  if (kind == 2 && bytes == sizeof(FeHeParams)) return cudaMemcpyToSymbol(c_fehe, params, bytes);
  return cudaErrorInvalidValue;
}

cudaError_t tile_launch(int pot, const GatherArgs& ga, const TilePlan& tp, const TileOut& to, cudaStream_t st) {
  switch (pot) {
    case TILE_LJ: return launch_one<TILE_LJ>(ga, tp, to, st);
    case TILE_PAIR: return launch_one<TILE_PAIR>(ga, tp, to, st);
    case TILE_EAM_DENSITY: return launch_one<TILE_EAM_DENSITY>(ga, tp, to, st);
    case TILE_EAM_FORCE: return launch_one<TILE_EAM_FORCE>(ga, tp, to, st);
    case TILE_AL_DENSITY: return launch_one<TILE_AL_DENSITY>(ga, tp, to, st);
    case TILE_AL_FORCE: return launch_one<TILE_AL_FORCE>(ga, tp, to, st);
    case TILE_FEHE_DENSITY: return launch_one<TILE_FEHE_DENSITY>(ga, tp, to, st);
    case TILE_FEHE_FORCE: return launch_one<TILE_FEHE_FORCE>(ga, tp, to, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace rb
