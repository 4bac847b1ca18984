// cells.cuh -- GPU cell-list build: bin -> count -> scan -> scatter -> per-cell order.
//
// Replaces the neighbor search under the reference path:
//   * vesin CellList::add_point / cells_ buckets  (third_party/vesin/vesin-single-build.cpp:1421-1442)
//   * the O(N^2) scan of CppCore/rgpot/nlist/vesin_visit.hpp:82-110 (LJ)
// The result is a cell-sorted copy of the atoms (double4: x, y, z, packed meta) plus the CSR
// cell starts.  Atoms inside a cell are ordered by original index, so every later summation
// has a fixed order and results are bit-reproducible run to run (no floating-point atomics).
#pragma once

#include <limits.h>

#include "common.cuh"

namespace rb {

// ---- bounding box of free (non-periodic) directions ---------------------------------------
// order-preserving map double -> unsigned 64 so min/max can use integer atomics
RB_D unsigned long long dkey(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
RB_HD double dkey_inv(unsigned long long k) {
  unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d;
  memcpy(&d, &u, sizeof d);
  return d;
#endif
}

// mm[0..2] = min keys, mm[3..5] = max keys (pre-initialised to ~0 / 0)
static __global__ void k_bbox(const double* __restrict__ pos, int n, unsigned long long* __restrict__ mm) {
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned long long key = dkey(pos[3 * (size_t)i + k]);
      lo[k] = key < lo[k] ? key : lo[k];
      hi[k] = key > hi[k] ? key : hi[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo[k], off);
      const unsigned long long b = __shfl_xor_sync(0xffffffffu, hi[k], off);
      lo[k] = a < lo[k] ? a : lo[k];
      hi[k] = b > hi[k] ? b : hi[k];
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&mm[k], lo[k]);
      atomicMax(&mm[3 + k], hi[k]);
    }
  }
}

// ---- binning ----------------------------------------------------------------------------
// cell coordinate + periodic wrap count of one position (vesin add_point :1421-1442; the
// bounding-box form of BoundingBox::cartesian_to_fractional :244-260 for free directions).
RB_D void locate(const Box& box, const Grid& g, double x, double y, double z, int c[3], int s[3]) {
  const double p[3] = {x, y, z};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double f;
    if (box.periodic[k]) {
      f = p[0] * box.Hinv[k] + p[1] * box.Hinv[3 + k] + p[2] * box.Hinv[6 + k];
    } else {
      f = (p[k] - box.lo[k]) / box.ext[k];
    }
    // clamp before the int conversion so absurd coordinates cannot overflow
    double cf = floor(f * (double)g.nc[k]);
    cf = fmin(fmax(cf, -2.0e9), 2.0e9);
    int ci = (int)cf;
    if (box.periodic[k]) {
      divmod_i(ci, g.nc[k], s[k], c[k]);
    } else {
      c[k] = min(max(ci, 0), g.nc[k] - 1);
      s[k] = 0;
    }
  }
}

static __global__ void k_bin(const double* __restrict__ pos, int n, Box box, Grid g,
                      int* __restrict__ cell_of, int* __restrict__ rank, int* __restrict__ ashift,
                      int* __restrict__ counts, Status* __restrict__ st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = pos[3 * (size_t)i], y = pos[3 * (size_t)i + 1], z = pos[3 * (size_t)i + 2];
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) {
    atomicOr(&st->flags, kFlagNonFinite);
    atomicMin(&st->first_bad_atom, i);
    cell_of[i] = 0;
    ashift[i] = 0;
    rank[i] = atomicAdd(&counts[0], 1);
    return;
  }
  int c[3], s[3];
  locate(box, g, x, y, z, c, s);
  unsigned fl = 0;
  if ((s[0] | s[1] | s[2]) != 0) fl |= kFlagAtomShift;
  if (abs(s[0]) > 1 || abs(s[1]) > 1 || abs(s[2]) > 1) fl |= kFlagBigShift;  // the brick kernels cover +-1
  if (abs(s[0]) > 511 || abs(s[1]) > 511 || abs(s[2]) > 511) fl |= kFlagShiftRange;
  if (fl) atomicOr(&st->flags, fl);
  const int cell = (c[2] * g.nc[1] + c[1]) * g.nc[0] + c[0];
  cell_of[i] = cell;
  ashift[i] = ((s[0] + 512) & 1023) | (((s[1] + 512) & 1023) << 10) | (((s[2] + 512) & 1023) << 20);
  rank[i] = atomicAdd(&counts[cell], 1);  // integer atomic: the COUNT is order-independent
}

// ---- exclusive scan over the cell counts (three small kernels) ------------------------------
constexpr int kScanBlock = 1024;
constexpr int kScanItems = 4;  // per thread -> 4096 cells per block

static __global__ void k_scan_partial(const int* __restrict__ counts, int n, int* __restrict__ block_sums) {
  __shared__ int warp_sums[32];
  const int base = blockIdx.x * kScanBlock * kScanItems;
  int v = 0;
  for (int t = 0; t < kScanItems; ++t) {
    const int idx = base + threadIdx.x * kScanItems + t;
    if (idx < n) v += counts[idx];
  }
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = warp_sums[threadIdx.x];
    for (int off = 16; off > 0; off >>= 1) w += __shfl_down_sync(0xffffffffu, w, off);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = w;
  }
}

// starts[c] = exclusive prefix of counts; starts[n] = total.  Every block adds up the sums of the
// blocks before it on its own (a few hundred integers at most) instead of waiting for a scan launch.
static __global__ void k_scan_final(const int* __restrict__ counts, int n, const int* __restrict__ block_sums,
                             int* __restrict__ starts) {
  __shared__ int sh[kScanBlock];
  __shared__ int wsum[32];
  __shared__ int block_off;
  {
    int part = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += kScanBlock) part += block_sums[j];
    for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x];
      for (int off = 16; off > 0; off >>= 1) w += __shfl_down_sync(0xffffffffu, w, off);
      if (threadIdx.x == 0) block_off = w;
    }
    __syncthreads();
  }
  const int base = blockIdx.x * kScanBlock * kScanItems;
  int v[kScanItems];
  int sum = 0;
  for (int t = 0; t < kScanItems; ++t) {
    const int idx = base + threadIdx.x * kScanItems + t;
    v[t] = idx < n ? counts[idx] : 0;
    sum += v[t];
  }
  sh[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < kScanBlock; off <<= 1) {
    const int add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += add;
    __syncthreads();
  }
  int run = block_off + sh[threadIdx.x] - sum;
  for (int t = 0; t < kScanItems; ++t) {
    const int idx = base + threadIdx.x * kScanItems + t;
    if (idx < n) starts[idx] = run;
    run += v[t];
    if (idx == n - 1) starts[n] = run;
  }
}

// ---- scatter + deterministic per-cell order --------------------------------------------------
static __global__ void k_scatter(const int* __restrict__ cell_of, const int* __restrict__ rank,
                          const int* __restrict__ starts, int n, int* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) order[starts[cell_of[i]] + rank[i]] = i;
}

// deterministic order inside a cell: every atom finds its rank among the atoms of its cell (the
// number of cell mates with a smaller original index) and moves there.  One thread per atom, so
// the cost does not depend on how few cells there are.
static __global__ void k_cell_order(const int* __restrict__ starts, const int* __restrict__ cell_of,
                             const int* __restrict__ order_in, int n, int* __restrict__ order_out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int i = order_in[s];
  const int c = cell_of[i];
  const int b = starts[c], e = starts[c + 1];
  int rank = 0;
  for (int t = b; t < e; ++t) rank += order_in[t] < i;
  order_out[b + rank] = i;
}

// k_cell_order and k_gather_sorted in one pass: the atom finds its rank inside its cell and writes
// its sorted record there directly
static __global__ void k_order_gather(const double* __restrict__ pos, const int* __restrict__ Z,
                               const int* __restrict__ starts, const int* __restrict__ cell_of,
                               const int* __restrict__ order_in, const int* __restrict__ ashift, int n,
                               int check_z, double4* __restrict__ spos, int* __restrict__ scell,
                               Status* __restrict__ st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int i = order_in[s];
  const int c = cell_of[i];
  const int b = starts[c], e = starts[c + 1];
  int rank = 0;
  for (int t = b; t < e; ++t) rank += order_in[t] < i;
  int species = 0;
  if (check_z == 2) {
    species = Z[i] & 15;  // ZBL: the host already mapped atomic numbers to type indices
  } else if (check_z) {
    // check_z 1: CuH2 (29 -> 0, 1 -> 1, rgpot_cuh2.f90:463-489); 3: FeHe (26 -> 0, 2 -> 1, rgpot_fehe.f90:541-569)
    const int z = Z[i];
    if (z == (check_z == 3 ? 26 : 29)) {
      species = 0;
    } else if (z == (check_z == 3 ? 2 : 1)) {
      species = 1;
    } else {
      atomicOr(&st->flags, kFlagForeignZ);
      atomicMin(&st->first_bad_atom, i);
    }
  }
  const int sh = ashift[i];
  const unsigned long long meta =
      pack_meta(i, species, (sh & 1023) - 512, ((sh >> 10) & 1023) - 512, ((sh >> 20) & 1023) - 512);
  spos[b + rank] = make_double4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2],
                                __longlong_as_double((long long)meta));
  scell[b + rank] = c;
}

// sorted records; species from atomic numbers (rgpot_cuh2.f90:463-489 classify) when check_z
static __global__ void k_gather_sorted(const double* __restrict__ pos, const int* __restrict__ Z,
                                const int* __restrict__ order, const int* __restrict__ cell_of,
                                const int* __restrict__ ashift, int n, int check_z,
                                double4* __restrict__ spos, int* __restrict__ scell,
                                Status* __restrict__ st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int i = order[s];
  int species = 0;
  if (check_z == 2) {
    species = Z[i] & 15;  // ZBL: the host already mapped atomic numbers to type indices
  } else if (check_z) {
    const int z = Z[i];
    if (z == 29) {
      species = 0;
    } else if (z == 1) {
      species = 1;
    } else {
      atomicOr(&st->flags, kFlagForeignZ);
      atomicMin(&st->first_bad_atom, i);
    }
  }
  const int sh = ashift[i];
  const unsigned long long meta =
      pack_meta(i, species, (sh & 1023) - 512, ((sh >> 10) & 1023) - 512, ((sh >> 20) & 1023) - 512);
  spos[s] = make_double4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2],
                         __longlong_as_double((long long)meta));
  scell[s] = cell_of[i];
}

// ---- species bookkeeping for table-driven pair potentials (ZBL) on device-resident input ---------
// which atomic numbers occur (bit z of an 8-word bitmap; bit 0 of word 8 flags a number outside 0..255)
static __global__ void k_species_present(const int* __restrict__ Z, int n, unsigned* __restrict__ bitmap) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int z = Z[i];
    if (z < 0 || z > 255)
      atomicOr(&bitmap[8], 1u);
    else if (!((bitmap[z >> 5] >> (z & 31)) & 1u))  // (a stale read only costs a redundant atomic)
      atomicOr(&bitmap[z >> 5], 1u << (z & 31));
  }
}

// atomic number -> type index through a 256-entry table
static __global__ void k_map_types(const int* __restrict__ Z, int n, const int* __restrict__ lut, int* __restrict__ types) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) types[i] = lut[Z[i] & 255];
}

// ---- deterministic energy reduction --------------------------------------------------------
// per-block partial sums were written by the force kernels in block order; one block folds them
// in a fixed tree.  out[0] = sum
static __global__ void k_reduce_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// long arrays, one launch: block b folds the fixed slice [b * per, (b + 1) * per); the block that
// finishes last (a ticket in the status word) folds the slice sums the way k_reduce_partials would
// with 128 threads -- the order of every addition is a function of n alone
static __global__ void k_reduce_slices(const double* __restrict__ partials, int n, int per, double* __restrict__ slice_sums,
                                Status* __restrict__ st, double* __restrict__ out) {
  __shared__ double sh[256];
  __shared__ int last;
  const int lo = blockIdx.x * per, hi = min(n, lo + per);
  double acc = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += partials[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    slice_sums[blockIdx.x] = sh[0];
    __threadfence();
    last = atomicAdd(&st->ticket, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int nb = gridDim.x;
  acc = 0.0;
  if (threadIdx.x < 128)
    for (int i = threadIdx.x; i < nb; i += 128) acc += __ldcg(slice_sums + i);
  sh[threadIdx.x] = threadIdx.x < 128 ? acc : 0.0;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = sh[0];
    st->ticket = 0;  // the ticket serves the next fold of this call (density / force passes share the status)
  }
}

// fixed-order block sum helper used by the force kernels (blockDim multiple of 32, <= 1024)
RB_D double block_sum(double v, double* sh32) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh32[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < (int)((blockDim.x + 31) >> 5) ? sh32[lane] : 0.0;
    for (int off = 16; off > 0; off >>= 1) r += __shfl_down_sync(0xffffffffu, r, off);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

}  // namespace rb
