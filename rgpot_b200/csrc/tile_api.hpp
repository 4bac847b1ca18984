// tile_api.hpp -- interface between the engine (engine.cu) and the tile kernels (tile.cu, a
// translation unit of its own so the two compile side by side).
#pragma once

#include <cuda_runtime.h>

#include "brick.cuh"  // GatherArgs, TileOut, TILE_* (templates only: nothing is instantiated by including it)

namespace rb {

constexpr int kTileMaxHalo = 1024;   // halo cells of one box (e.g. (8 + 4) x (4 + 4) x (4 + 4))
constexpr int kTileThreads = 256;    // eight warps: one cell group each per round
constexpr int kTileMaxGroupCells = 32;

struct TilePlan {
  int bd[3];     // cells per brick and direction (multiples of gd)
  int nb[3];     // bricks per direction
  int ns[3];     // stencil half width per direction (cells)
  int gd[3];     // cells per GROUP: the unit one warp evaluates (<= 16 atoms per pass)
  int nbricks;
  int cap_cand;  // staged halo atoms a block can hold (multiple of 8; slot cap_cand is the far-away dummy)
  int wcap;      // window entries a warp searches per chunk (multiple of 64)
  int threads;
  float thr2;    // threshold on the tensor-core d2 (a superset of the exact accept test)
  unsigned smem;
};

// dynamic shared memory of one block: FP64 words per staged atom, the 16-byte FP16 record, the info
// word; per warp the window (16-bit staged indices), the survivor bit masks and the row table
inline unsigned tile_smem_bytes(int cap_cand, int wcap, int threads, int words) {
  const unsigned nwarps = (unsigned)threads / 32u;
  unsigned b = (unsigned)cap_cand * (8u * (unsigned)words) + 16u * ((unsigned)cap_cand + 1u) +
               2u * (((unsigned)cap_cand + 8u) & ~7u);
  b += nwarps * (2u * (unsigned)wcap + 4u * 32u * ((unsigned)wcap / 64u) + 128u);
  return b + 64u;
}

// coefficient tables of this translation unit (one copy of the __constant__ blocks per unit)
cudaError_t tile_upload_constants(int kind, const void* params, size_t bytes, const double* exp2tab);

// launch k_tile<pot>; returns the launch error, if any
cudaError_t tile_launch(int pot, const GatherArgs& ga, const TilePlan& tp, const TileOut& to, cudaStream_t st);

}  // namespace rb
