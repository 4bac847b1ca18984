// hash.cuh -- the result-cache key of rgpot's Potential (CppCore/rgpot/Potential.hpp:324-336) on
// the GPU: XXH3_64bits (xxHash 0.8.3, the reference's pinned subproject, default secret, seed 0)
// over a system's positions, atomic numbers and box, XOR the hashes of the potential type and the
// parameter fingerprint.  Bit-exact integer work: for the same type and fingerprint a key made
// here equals the key the reference makes on the CPU, so `forceBatch`'s hit / miss compaction
// (Potential.hpp:251-306) can hash a whole batch in one launch from device-resident inputs.
//
// Eight lanes per system (four systems per warp): the lanes are the eight accumulators of the
// long-input loop, first over the positions, then over the atomic numbers; the first lane hashes
// the box; inputs of at most 240 bytes take the short forms on the first lane of the group.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "common.cuh"

namespace rb {

#define RB_XXH_SECRET_BYTES                                                                                    \
  0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c, 0xde, 0xd4,  \
      0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f, 0xcb, 0x79, 0xe6,    \
      0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21, 0xb8, 0x08, 0x46, 0x74,    \
      0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c, 0x3c, 0x28, 0x52, 0xbb, 0x91,    \
      0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3, 0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d,    \
      0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8, 0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34,    \
      0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d, 0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31,    \
      0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64, 0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5,    \
      0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb, 0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16,    \
      0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e, 0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8,    \
      0xd1, 0x7a, 0xd0, 0x31, 0xce, 0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca,    \
      0xbb, 0x4b, 0x40, 0x7e

__constant__ unsigned char c_xxh_secret[192] = {RB_XXH_SECRET_BYTES};
static const unsigned char h_xxh_secret[192] = {RB_XXH_SECRET_BYTES};

constexpr uint64_t kP32_1 = 0x9E3779B1ull, kP32_2 = 0x85EBCA77ull, kP32_3 = 0xC2B2AE3Dull;
constexpr uint64_t kP64_1 = 0x9E3779B185EBCA87ull, kP64_2 = 0xC2B2AE3D27D4EB4Full, kP64_3 = 0x165667B19E3779F9ull,
                   kP64_4 = 0x85EBCA77C2B2AE63ull, kP64_5 = 0x27D4EB2F165667C5ull;
constexpr uint64_t kMX1 = 0x165667919E3779F9ull, kMX2 = 0x9FB21C651E98DF25ull;

RB_HD const unsigned char* xxh_secret() {
#ifdef __CUDA_ARCH__
  return c_xxh_secret;
#else
  return h_xxh_secret;
#endif
}
// little-endian reads at any alignment (the inputs are 4-byte aligned at least, the secret is not)
RB_HD uint32_t rd32(const unsigned char* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
RB_HD uint64_t rd64(const unsigned char* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
RB_HD uint64_t in64(const unsigned char* p) {
#ifdef __CUDA_ARCH__
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p);  // device inputs are 4-byte aligned: two word loads
  return (uint64_t)q[0] | ((uint64_t)q[1] << 32);
#else
  uint64_t v;  // host: any alignment, any length
  memcpy(&v, p, 8);
  return v;
#endif
}
RB_HD uint64_t mul_fold(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return (a * b) ^ __umul64hi(a, b);
#else
  const unsigned __int128 m = (unsigned __int128)a * b;
  return (uint64_t)m ^ (uint64_t)(m >> 64);
#endif
}
RB_HD uint64_t swap32(uint32_t x) { return ((x & 0xffu) << 24) | ((x & 0xff00u) << 8) | ((x >> 8) & 0xff00u) | (x >> 24); }
RB_HD uint64_t swap64(uint64_t x) { return (swap32((uint32_t)x) << 32) | swap32((uint32_t)(x >> 32)); }
RB_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
RB_HD uint64_t xxh64_avalanche(uint64_t h) {
  h ^= h >> 33;
  h *= kP64_2;
  h ^= h >> 29;
  h *= kP64_3;
  h ^= h >> 32;
  return h;
}
RB_HD uint64_t xxh3_avalanche(uint64_t h) {
  h ^= h >> 37;
  h *= kMX1;
  h ^= h >> 32;
  return h;
}
RB_HD uint64_t mix16(const unsigned char* in, const unsigned char* sec) {
  return mul_fold(in64(in) ^ rd64(sec), in64(in + 8) ^ rd64(sec + 8));
}

// inputs of at most 240 bytes (xxhash.h XXH3_len_0to16 / 17to128 / 129to240, seed 0); on the device
// len % 4 == 0 (arrays of doubles and ints)
RB_HD uint64_t xxh3_short(const unsigned char* in, size_t len) {
  const unsigned char* s = xxh_secret();
  if (len == 0) return xxh64_avalanche(rd64(s + 56) ^ rd64(s + 64));
  if (len < 4) {
    const uint32_t c = ((uint32_t)in[0] << 16) | ((uint32_t)in[len >> 1] << 24) | (uint32_t)in[len - 1] | ((uint32_t)len << 8);
    return xxh64_avalanche((uint64_t)c ^ (uint64_t)(rd32(s) ^ rd32(s + 4)));
  }
  if (len <= 8) {
    const uint32_t a = rd32(in), b = rd32(in + len - 4);
    uint64_t h = ((uint64_t)b + ((uint64_t)a << 32)) ^ (rd64(s + 8) ^ rd64(s + 16));
    h ^= rotl64(h, 49) ^ rotl64(h, 24);
    h *= kMX2;
    h ^= (h >> 35) + len;
    h *= kMX2;
    return h ^ (h >> 28);
  }
  if (len <= 16) {
    const uint64_t lo = in64(in) ^ (rd64(s + 24) ^ rd64(s + 32)), hi = in64(in + len - 8) ^ (rd64(s + 40) ^ rd64(s + 48));
    return xxh3_avalanche(len + swap64(lo) + hi + mul_fold(lo, hi));
  }
  if (len <= 128) {
    uint64_t acc = len * kP64_1;
    int i = (int)((len - 1) / 32);
    do {
      acc += mix16(in + 16 * i, s + 32 * i);
      acc += mix16(in + len - 16 * (i + 1), s + 32 * i + 16);
    } while (i-- != 0);
    return xxh3_avalanche(acc);
  }
  uint64_t acc = len * kP64_1;
  const int rounds = (int)len / 16;
  for (int i = 0; i < 8; ++i) acc += mix16(in + 16 * i, s + 16 * i);
  uint64_t acc_end = mix16(in + len - 16, s + 136 - 17);
  acc = xxh3_avalanche(acc);
  for (int i = 8; i < rounds; ++i) acc_end += mix16(in + 16 * i, s + 16 * (i - 8) + 3);
  return xxh3_avalanche(acc + acc_end);
}

// host: any length (used for the two 8-byte words every key shares, and by the engine's self-test)
inline uint64_t xxh3_host(const unsigned char* in, size_t len) {
  if (len <= 240) return xxh3_short(in, len);
  const unsigned char* s = h_xxh_secret;
  uint64_t acc[8] = {kP32_3, kP64_1, kP64_2, kP64_3, kP64_4, kP32_2, kP64_5, kP32_1};
  auto stripe = [&](const unsigned char* p, const unsigned char* sec) {
    uint64_t v[8];
    for (int l = 0; l < 8; ++l) v[l] = in64(p + 8 * l);
    for (int l = 0; l < 8; ++l) {
      const uint64_t k = v[l] ^ rd64(sec + 8 * l);
      acc[l] += v[l ^ 1] + (k & 0xffffffffull) * (k >> 32);
    }
  };
  const size_t block = 64 * 16, nblocks = (len - 1) / block;
  for (size_t b = 0; b < nblocks; ++b) {
    for (int t = 0; t < 16; ++t) stripe(in + b * block + 64 * t, s + 8 * t);
    for (int l = 0; l < 8; ++l) acc[l] = ((acc[l] ^ (acc[l] >> 47)) ^ rd64(s + 128 + 8 * l)) * kP32_1;
  }
  const size_t nstripes = ((len - 1) - block * nblocks) / 64;
  for (size_t t = 0; t < nstripes; ++t) stripe(in + nblocks * block + 64 * t, s + 8 * t);
  stripe(in + len - 64, s + 192 - 64 - 7);
  uint64_t r = len * kP64_1;
  for (int i = 0; i < 4; ++i) r += mul_fold(acc[2 * i] ^ rd64(s + 11 + 16 * i), acc[2 * i + 1] ^ rd64(s + 11 + 16 * i + 8));
  return xxh3_avalanche(r);
}

// long input (> 240 bytes), eight cooperating lanes l = 0..7 of one warp (xxhash.h
// XXH3_hashLong_internal_loop, XXH3_scalarRound, XXH3_scalarScrambleRound, XXH3_mergeAccs);
// `base` is the group's first lane.  Every lane of the group returns the hash.
__device__ inline uint64_t xxh3_long_group(const unsigned char* in, size_t len, int l, int base, unsigned mask,
                                           const uint64_t* sw /* the secret as 24 words, shared memory */) {
  const unsigned char* s = c_xxh_secret;
  const uint64_t init[8] = {kP32_3, kP64_1, kP64_2, kP64_3, kP64_4, kP32_2, kP64_5, kP32_1};
  uint64_t acc = init[l];
  auto absorb = [&](uint64_t v, uint64_t key) {
    const uint64_t other = __shfl_xor_sync(mask, v, 1);  // adjacent lanes swap their input words
    const uint64_t k = v ^ key;
    acc += other + (k & 0xffffffffull) * (k >> 32);
  };
  const bool aligned8 = (reinterpret_cast<uintptr_t>(in) & 7) == 0;
  const size_t block = 64 * 16, nblocks = (len - 1) / block;
  const uint64_t scramble_key = sw[16 + l];
  for (size_t b = 0; b < nblocks; ++b) {
    // a block of 16 stripes: all sixteen loads of this lane are issued before the first is used
    uint64_t v[16];
    const unsigned char* p = in + b * block + 8 * l;
    if (aligned8) {
#pragma unroll
      for (int t = 0; t < 16; ++t) v[t] = __ldg(reinterpret_cast<const unsigned long long*>(p + 64 * t));
    } else {
#pragma unroll
      for (int t = 0; t < 16; ++t) v[t] = in64(p + 64 * t);
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) absorb(v[t], sw[t + l]);
    acc = ((acc ^ (acc >> 47)) ^ scramble_key) * kP32_1;
  }
  const int nstripes = (int)(((len - 1) - block * nblocks) / 64);
  {
    // the partial block (up to 15 stripes) and the last stripe: loads first, as above
    uint64_t v[16];
    const unsigned char* p = in + nblocks * block + 8 * l;
#pragma unroll
    for (int t = 0; t < 15; ++t) v[t] = t < nstripes ? in64(p + 64 * t) : 0ull;
    v[15] = in64(in + len - 64 + 8 * l);
#pragma unroll
    for (int t = 0; t < 15; ++t)
      if (t < nstripes) absorb(v[t], sw[t + l]);  // (uniform over the eight lanes of the group)
    absorb(v[15], rd64(s + 192 - 64 - 7 + 8 * l));
  }
  // merge: pairs of accumulators, then the sum over the four pairs in order
  const uint64_t keyed = acc ^ rd64(s + 11 + 8 * l);
  const uint64_t partner = __shfl_xor_sync(mask, keyed, 1);
  const uint64_t mixed = mul_fold(keyed, partner);  // even lanes hold the pair's term
  uint64_t r = len * kP64_1;
#pragma unroll
  for (int i = 0; i < 4; ++i) r += __shfl_sync(mask, mixed, base + 2 * i);
  return xxh3_avalanche(r);
}

struct KeyArgs {
  const size_t* offsets;  // nsys + 1 atom offsets
  const double* pos;      // packed positions
  const int* Z;           // packed atomic numbers
  const double* boxes;    // 9 per system
  uint64_t tail;          // XXH3(type) ^ XXH3(paramsKey), the same for every system
  uint64_t* keys;
  size_t nsys;
};

static __global__ void __launch_bounds__(256, 4) k_cache_keys(const KeyArgs a) {
  __shared__ uint64_t sw[24];
  if (threadIdx.x < 24) sw[threadIdx.x] = rd64(c_xxh_secret + 8 * threadIdx.x);
  __syncthreads();
  // four systems per warp, eight lanes each (the eight accumulators of the long form)
  const int lane = threadIdx.x & 31, grp = lane >> 3, l = lane & 7;
  const size_t sys = ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + grp;
  if (sys >= a.nsys) return;  // whole groups leave together
  const unsigned mask = 0xffu << (8 * grp);
  const size_t o = a.offsets[sys], n = a.offsets[sys + 1] - o;
  const unsigned char* pp = reinterpret_cast<const unsigned char*>(a.pos + 3 * o);
  const unsigned char* zp = reinterpret_cast<const unsigned char*>(a.Z + o);
  const size_t plen = 24 * n, zlen = 4 * n;
  // the branch conditions are uniform over a group
  uint64_t hp = 0, hz = 0;
  if (plen > 240)
    hp = xxh3_long_group(pp, plen, l, 8 * grp, mask, sw);
  else if (l == 0)
    hp = xxh3_short(pp, plen);
  if (zlen > 240)
    hz = xxh3_long_group(zp, zlen, l, 8 * grp, mask, sw);
  else if (l == 0)
    hz = xxh3_short(zp, zlen);
  if (l == 0) {
    const uint64_t hb = xxh3_short(reinterpret_cast<const unsigned char*>(a.boxes + 9 * sys), 72);
    a.keys[sys] = hp ^ hz ^ hb ^ a.tail;
  }
}

}  // namespace rb
