// engine.cu -- host side of the B200 engine: handle state, path selection, launches, C ABI.
//
// One handle = one CUDA device + one stream + its own scratch buffers (no process-global state:
// the reference's Fortran entry keeps a module-saved table and is ProcessSerial,
// CppCore/rgpot/fortran/FortranPots.hpp:14-18; this engine is PerInstance).
//
// There is deliberately no CPU path in this file: if CUDA is unusable every entry point fails.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rgpot_b200.h"
#include "cells.cuh"
#include "common.cuh"
#include "gather.cuh"
#include "physics.cuh"
#include "small.cuh"
#include "single.cuh"
#include "brick.cuh"
#include "hash.cuh"
#include "hostpipe.hpp"
#include "tile_api.hpp"

using namespace rb;

namespace {

constexpr int kStatusCuda = 100;
constexpr int kStatusArgs = 101;
constexpr int kStatusRange = 102;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct HostBuf {  // pinned
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

// CuH2 defaults (rgpot_cuh2.f90:58-105) with the five cutoff shifts of cut_shifted (:118-135)
// physics threshold on d2: d2 < rc2 (vesin) and sqrt(d2) < r_phys (the kernels' own radial test)
// folded into one compare, plus the identities the fast radial terms rely on
void finish_eam_params(EamParams& p, double r_phys) {
  double T = r_phys * r_phys;
  while (sqrt(T) >= r_phys) T = nextafter(T, 0.0);
  while (sqrt(nextafter(T, INFINITY)) < r_phys) T = nextafter(T, INFINITY);
  T = nextafter(T, INFINITY);  // smallest d2 whose rounded sqrt reaches r_phys
  p.rc2_phys = T < p.rc2 ? T : p.rc2;
  p.cu_delta = 2.0 * p.ba[0] - p.bb[0];
  p.cucu_square = (p.aa[0] == 2.0 * p.ab[0]) ? 1 : 0;
  p.cu_taylor = (fabs(p.cu_delta) * p.rcut < 1.0e-3) ? 1 : 0;
}

// evaluated on the host with libm, exactly like the reference does on every call.
EamParams make_eam_params() {
  EamParams p{};
  const double da[3] = {2862.3, 86.1495, 79.5013};
  const double aa[3] = {3.51236, 4.21154, 2.47961};
  const double db[3] = {-109.107, 15355.5, -107.555};
  const double ab[3] = {1.75618, 6.07604, 2.99918};
  for (int k = 0; k < 3; ++k) {
    p.da[k] = da[k];
    p.aa[k] = aa[k];
    p.db[k] = db[k];
    p.ab[k] = ab[k];
  }
  p.scale[0] = 0.273072;
  p.ba[0] = 3.69051;
  p.bb[0] = 7.38101;
  p.gamma[0] = 512.0;
  p.scale[1] = 2.14389;
  p.ba[1] = 3.77701;
  p.bb[1] = 0.0;
  p.gamma[1] = 0.0;
  const double fcu[8] = {-112.945,    8510.04,     -261734.0,     4780090.0,
                         -52341900.0, 339124000.0, -1201500000.0, 1796190000.0};
  const double fh[5] = {-81.7537, 838.668, -3952.35, 8767.87, -6599.37};
  memcpy(p.f_cu, fcu, sizeof fcu);
  memcpy(p.f_h, fh, sizeof fh);
  p.rcut = 6.1;
  p.rc2 = p.rcut * p.rcut;
  const double r = p.rcut;
  for (int k = 0; k < 3; ++k) {
    volatile double t1 = p.da[k] * exp(-p.aa[k] * r);
    volatile double t2 = p.db[k] * exp(-p.ab[k] * r);
    volatile double s = t1 + t2;
    p.phicut[k] = s - 0.0;
  }
  {
    // Cu: r**6 by square-and-multiply (libgfortran pow_r8_i4): r^2 * r^4
    volatile double t1 = exp(-p.ba[0] * r);
    volatile double t2 = p.gamma[0] * exp(-p.bb[0] * r);
    volatile double r2 = r * r;
    volatile double r4 = r2 * r2;
    volatile double r6 = r2 * r4;
    volatile double t3 = p.scale[0] * r6;
    volatile double s = t1 + t2;
    p.rhocut[0] = t3 * s - 0.0;
  }
  {
    volatile double t1 = exp(-p.ba[1] * r);
    volatile double t2 = p.gamma[1] * exp(-p.bb[1] * r);
    volatile double t3 = p.scale[1] * 1.0;
    volatile double s = t1 + t2;
    p.rhocut[1] = t3 * s - 0.0;
  }
  finish_eam_params(p, p.rcut);  // sqrt(d2) < rcut, rgpot_cuh2.f90:522
  return p;
}

// Fe-He coefficients, rgpot_fehe.f90:47-156
FeHeParams make_fehe_params() {
  FeHeParams p{};
  p.rcut_fe_pair = 5.3;
  p.rcut_fe_rho = 4.2;
  p.rcut_fehe_pair = 3.9028;
  p.rcut_fehe_rho = 4.10;
  p.rcut_he_pair = 5.4;
  p.fe_zbl_scale = 9.7342365892908e+03;
  const double w[4] = {0.18180, 0.50990, 0.28020, 0.02817};
  const double d[4] = {2.8616724320005e+01, 8.4267310396064e+00, 3.6030244464156e+00, 1.8028536321603e+00};
  const double b[4] = {7.4122709384068e+00, -6.4180690713367e-01, -2.6043547961722e+00, 6.2625393931230e-01};
  memcpy(p.fe_zbl_weight, w, sizeof w);
  memcpy(p.fe_zbl_decay, d, sizeof d);
  memcpy(p.fe_bridge, b, sizeof b);
  p.fe_bridge_lo = 1.0;
  p.fe_bridge_hi = 2.05;
  const double pk[13] = {2.2, 2.3, 2.4, 2.5, 2.6, 2.7, 2.8, 3.0, 3.3, 3.7, 4.2, 4.7, 5.3};
  const double pc[13] = {-2.7444805994228e+01, 1.5738054058489e+01,  2.2077118733936e+00,  -2.4989799053251e+00,
                         4.2099676494795e+00,  -7.7361294129713e-01, 8.0656414937789e-01,  -2.3194358924605e+00,
                         2.6577406128280e+00,  -1.0260416933564e+00, 3.5018615891957e-01,  -5.8531821042271e-02,
                         -3.0458824556234e-03};
  memcpy(p.fe_pair_knot, pk, sizeof pk);
  memcpy(p.fe_pair_coeff, pc, sizeof pc);
  const double rk[3] = {2.4, 3.2, 4.2};
  const double rcf[3] = {1.1686859407970e+01, -1.4710740098830e-02, 4.7193527075943e-01};
  memcpy(p.fe_rho_knot, rk, sizeof rk);
  memcpy(p.fe_rho_coeff, rcf, sizeof rcf);
  p.fe_emb_quadratic = -6.7314115586063e-04;
  p.fe_emb_quartic = 7.6514905604792e-08;
  const double hk[8] = {1.5440, 1.6155, 1.6896, 1.8017, 2.0482, 2.3816, 3.5067, 3.9028};
  const double hc[8] = {559.804426025391, -45.916354995728, 35.550312671661, 164.319865173340,
                        -1.727464405060,  0.106771826237,   0.073715269849,  0.038235287677};
  memcpy(p.fehe_knot, hk, sizeof hk);
  memcpy(p.fehe_coeff, hc, sizeof hc);
  p.sband_amp = 20.0;
  p.sband_decay = 1.5312426703208 / 0.529177210818181818;
  p.sband_emb_sqrt = 0.220813420485;
  p.sband_emb_quadratic = 1.367508764130;
  p.sband_emb_quartic = 3.382256025271;
  p.he_rm = 2.9683;
  p.he_c6 = 1.35186623;
  p.he_c8 = 0.41495143;
  p.he_c10 = 0.17151143;
  p.he_aa = 186924.404;
  p.he_alpha = 10.5717543;
  p.he_beta = -2.07758779;
  p.he_damp = 1.438;
  p.he_eps = 10.956 * 1.380658e-23 / 1.602177e-19;
  p.rho_floor = 1.0e-35;
  return p;
}

// EAM-Al (rgpot_eam_al.f90:44-76) mapped onto slot 0 of the same tables; shifts evaluated on the
// host with libm like the reference does (pair_shape(rcut), density_shape(rcut), :103-122).
EamParams make_eam_al_params() {
  EamParams p{};
  p.gauge = -0.60647886749203;
  p.da[0] = 2294.3609145535;
  p.aa[0] = 3.0205380362464;
  p.db[0] = -192.06894637533;
  p.ab[0] = 1.5102690181232;
  p.scale[0] = 0.87928364657088;
  p.ba[0] = 3.3068828537934;
  p.bb[0] = 6.6137657075868;
  p.gamma[0] = 512.0;
  const double c[8] = {4.4572157051836,  193.21775368064, -1173.6502684704, 4203.3500116196,
                       -8785.1680280827, 10632.994102532, -6921.0410455328, 1875.2698365752};
  memcpy(p.f_cu, c, sizeof c);
  p.rcut = 7.1;
  p.rc2 = p.rcut * p.rcut;
  const double r = p.rcut;
  {
    volatile double t1 = p.da[0] * exp(-p.aa[0] * r);
    volatile double t2 = p.db[0] * exp(-p.ab[0] * r);
    p.phicut[0] = t1 + t2;
  }
  {
    volatile double r2 = r * r;
    volatile double r4 = r2 * r2;
    volatile double r6 = r2 * r4;  // r**6 by square-and-multiply
    volatile double s = exp(-p.ba[0] * r) + p.gamma[0] * exp(-p.bb[0] * r);
    volatile double shape = r6 * s;
    p.rhocut[0] = p.scale[0] * shape;
  }
  // every pair term vanishes from sqrt(0.9999) rcut on (rgpot_eam_al.f90:39-42, 94-99)
  finish_eam_params(p, sqrt(0.9999) * p.rcut);
  return p;
}

}  // namespace

struct RgpotB200Pot {
  RgpotB200Config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};  // copy/compute pipeline of the batch path
  cudaEvent_t ev_ready = nullptr;
  int sm_count = 0;
  size_t smem_optin = 0;
  size_t smem_per_sm = 0;
  double lj_shiftU = 0.0;
  double morse_ecut = 0.0;  // MorsePot.hpp:56-61
  std::vector<int> zbl_species;  // sorted unique atomic numbers of the cached table (ZBLPot.cc:150-166)
  DevBuf zbl_table, zbl_bitmap;
  double lj_rc = 0.0;  // effective cutoff (cluster: <= 0 -> 1e6)
  int forced_path = 0;
  // tuning overrides of the brick kernels (0 = automatic), rgpot_b200_set_option
  int opt_brick[3] = {0, 0, 0};
  int opt_tpa = 0;
  int opt_threads = 0;
  int opt_list = 0;
  int opt_slack = 0;  // percent of extra staged capacity over the mean halo population (0 = automatic)
  int opt_fine = 0;          // 1: never use half-cutoff cells
  int opt_bricks = 1;         // 0: gather kernels only (test hook)
  int opt_handoff = 1;        // EAM brick kernels: the force pass reuses the density pass's survivor lists
  int opt_tile = 0;           // 1: large systems on the tile kernels (tile.cu, opt-in); 0 = the brick kernels
  int opt_tile_group = 0;     // cells per group as decimal xyz (0 = automatic)
  int opt_tile_brick = 0;     // cells per brick as decimal xyz (0 = automatic)
  int opt_tile_cap = 0;       // staged capacity override (test hook: small values exercise cut boxes / the slow path)
  int opt_tile_wcap = 0;      // window chunk override
  int opt_batch_chunk = 0;    // systems per pipeline chunk of a host batch (0 = automatic)
  int opt_half = 1;           // batch kernel: every pair once (half lists + fixed-point scatter); 0 = full lists
  int opt_one = 1;            // single EAM systems of 64 .. 512 atoms on the many-block kernels (single.cuh)
  long opt_fine_min = 20000;  // atoms from which half-cutoff cells are used
  long launches = 0;
  long pool_used = 0;   // survivor-list entries the density pass handed to the force pass
  long slow_atoms = 0;  // diagnostics of the last single-system call (brick slow path)
  long prof[4] = {0, 0, 0, 0};  // RB_BRICK_TIMING builds only
  // feedback from the brick kernels: the largest halo population seen for a given (atoms, grid,
  // brick) signature sizes the staged capacity of the next call with that signature
  struct BrickMemo {
    long n = -1;
    int nc[3] = {0, 0, 0};
    int bd[3] = {0, 0, 0};  // memo_pending: the brick of the plan in flight; memo: of the last record
    int max_halo = 0;
    // one record per brick shape met with this (n, nc): a shape that no longer fits shared memory
    // with its own record makes the planner shrink the brick, and the smaller shape keeps its own
    struct Shape {
      int bd[3];
      int max_halo;
    } shapes[8];
    int nshapes = 0;
    bool matches(long n_, const int* nc_) const { return n == n_ && nc[0] == nc_[0] && nc[1] == nc_[1] && nc[2] == nc_[2]; }
    // largest halo recorded for this shape; failing that, for the smallest recorded shape that
    // contains it (a bound in all but pathological alignments -- and a short capacity only cuts)
    int lookup(const int* bd_) const {
      int best = 0;
      long best_vol = 0;
      for (int s = 0; s < nshapes; ++s) {
        const int* b = shapes[s].bd;
        if (b[0] == bd_[0] && b[1] == bd_[1] && b[2] == bd_[2]) return shapes[s].max_halo;
        if (b[0] >= bd_[0] && b[1] >= bd_[1] && b[2] >= bd_[2]) {
          const long vol = (long)b[0] * b[1] * b[2];
          if (best == 0 || vol < best_vol) {
            best = shapes[s].max_halo;
            best_vol = vol;
          }
        }
      }
      return best;
    }
    void record(const int* bd_, int halo) {
      for (int s = 0; s < nshapes; ++s) {
        int* b = shapes[s].bd;
        if (b[0] == bd_[0] && b[1] == bd_[1] && b[2] == bd_[2]) {
          shapes[s].max_halo = std::max(shapes[s].max_halo, halo);
          return;
        }
      }
      if (nshapes == 8) nshapes = 0;
      Shape& sh = shapes[nshapes++];
      for (int k = 0; k < 3; ++k) sh.bd[k] = bd_[k];
      sh.max_halo = halo;
    }
  } memo, memo_pending;
  int last_cap_cand = 0, last_cap_list = 0;  // what the last brick plan chose (rgpot_b200_stat)
  int last_used_tile = 0;                    // the last cell-list call ran on the tile kernels (tile.cu)
  std::string last_error;

  // single-system scratch
  DevBuf pos, Z, F, cell_of, rank, ashift, counts, starts, blocksums, order, order2, spos, scell, dfd, eemb,
      dfs, partials, partials2, energy, status, mm, listpool, listref, emit_pairs, emit_shifts, emit_count, emit_vecs, emit_dists;
  // batch scratch
  DevBuf b_pos, b_Z, b_F, b_off, b_boxes, b_E, b_flags, b_bad, b_map;
  HostBuf h_pos, h_Z, h_F, h_off, h_boxes, h_E, h_flags, h_bad, h_small;
  // ---- host side of the batch entry points (hostpipe.hpp) ----
  // helper threads that pack / unpack ForceBatch buffers into pinned staging (created on first use,
  // placed on the CPUs next to this device when the topology can be read)
  std::unique_ptr<rb::Team> team;
  int host_threads = 0;       // 0 = automatic
  int small_attr_set = 0;     // batch kernels whose dynamic shared-memory limit has been raised on this device
  size_t msg_sys_base = 0;    // index of this range's first system in the caller's batch (messages)
  DevBuf b_Z0;                // atomic numbers of one system, shared by a homogeneous batch
  // single small systems: one page-locked block the kernel reads and writes in place (zero copy)
  HostBuf h_fast;
  void* h_fast_dev = nullptr;  // its address as the device sees it
  unsigned one_call_id = 0;    // single.cuh: calls so far (the kernels' fallback word compares against it)
  DevBuf d_fast;              // device copy of the inputs + block partials + ticket (one system over many blocks)
  // pageable single systems: pinned ring for the staged H2D / D2H of large arrays
  HostBuf h_ring;
  cudaEvent_t ring_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ring_busy[8] = {false, false, false, false, false, false, false, false};  // a copy out of / into the slot may still be in flight
  unsigned ring_next = 0;                             // slots are handed out round robin across calls
  // batches cut across several GPUs: one single-device handle + one persistent host thread per device
  struct Shard {
    RgpotB200Pot* pot = nullptr;
    std::unique_ptr<rb::Worker> worker;
    int status = 0;
  };
  std::vector<Shard> shards;
  // outstanding async (device API) call, resolved by rgpot_b200_sync_status
  struct Pending {
    bool active = false;
    bool batch = false;        // a packed batch (else one system)
    bool small_single = false; // single system that ran as a batch of one
    long n = 0;
    size_t nsys = 0;
    const double* pos = nullptr;
    const int* Z = nullptr;
    double* F = nullptr;
    double* E = nullptr;
    const double* d_boxes = nullptr;
    double box[9] = {0};
    bool has_box = false;
    cudaStream_t st = nullptr;
    std::vector<size_t> offsets;
  } pend;

  ~RgpotB200Pot() {
    for (auto& sh : shards) {
      sh.worker.reset();  // joins the device's host thread
      delete sh.pot;
    }
    shards.clear();
    team.reset();
    cudaSetDevice(device);
    b_Z0.release();
    d_fast.release();
    h_fast.release();
    h_ring.release();
    for (auto& e : ring_ev)
      if (e) cudaEventDestroy(e);
    DevBuf* d[] = {&pos, &Z, &F, &cell_of, &rank, &ashift, &counts, &starts, &blocksums, &order, &order2,
                   &spos, &scell, &dfd, &dfs, &eemb, &partials, &partials2, &energy, &status, &mm, &listpool, &listref, &emit_pairs,
                   &emit_shifts, &emit_count, &emit_vecs, &emit_dists, &b_pos, &b_Z, &b_F, &b_off, &b_boxes, &b_E, &b_flags,
                   &b_bad, &b_map, &zbl_table, &zbl_bitmap};
    for (auto* b : d) b->release();
    HostBuf* h[] = {&h_pos, &h_Z, &h_F, &h_off, &h_boxes, &h_E, &h_flags, &h_bad, &h_small};
    for (auto* b : h) b->release();
    for (auto& s : pipe)
      if (s) cudaStreamDestroy(s);
    if (ev_ready) cudaEventDestroy(ev_ready);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

#define CU_TRY(pot, expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      (pot)->last_error = std::string("CUDA failure: ") + cudaGetErrorString(_e) + " at " + \
                          #expr;                                                            \
      return kStatusCuda;                                                                   \
    }                                                                                       \
  } while (0)

// RAII: make the handle's device current for the duration of an entry point, then put the caller's back
struct DevGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev == dev)
      prev = -1;  // already current: nothing to select, nothing to put back (a single small system's
                  // whole call is ~20 microseconds; two cudaSetDevice calls would be a tenth of it)
    else
      err = cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// helper threads of a handle (created on first use; on the CPUs next to the device when sysfs says
// which those are, so the pinned staging they fill is local to the GPU's PCIe root)
rb::Team* get_team(RgpotB200Pot* pot) {
  if (!pot->team) {
    int want = pot->host_threads > 0 ? pot->host_threads : std::max(1, std::min(14, rb::usable_cpus()));
    cpu_set_t allowed, local;
    CPU_ZERO(&allowed);
    const bool place = sched_getaffinity(0, sizeof allowed, &allowed) == 0 &&
                       rb::gpu_local_cpus(pot->device, allowed, local);
    if (place) want = std::min(want, std::max(1, CPU_COUNT(&local)));
    pot->team.reset(new rb::Team(want - 1, place ? &local : nullptr));
  }
  return pot->team.get();
}

int fail(RgpotB200Pot* pot, int status, const std::string& msg) {
  pot->last_error = msg;
  return status;
}

// ---- box / grid preparation (host, reference operation order) ---------------------------------
bool invert_box(const double* m, double* v) {
  // vesin Matrix::determinant / inverse, vesin-single-build.cpp:100-128
  const double det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                     m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (!(fabs(det) >= 1e-30)) return false;
  v[0] = (m[4] * m[8] - m[7] * m[5]) / det;
  v[1] = (m[2] * m[7] - m[1] * m[8]) / det;
  v[2] = (m[1] * m[5] - m[2] * m[4]) / det;
  v[3] = (m[5] * m[6] - m[3] * m[8]) / det;
  v[4] = (m[0] * m[8] - m[2] * m[6]) / det;
  v[5] = (m[3] * m[2] - m[0] * m[5]) / det;
  v[6] = (m[3] * m[7] - m[6] * m[4]) / det;
  v[7] = (m[6] * m[1] - m[0] * m[7]) / det;
  v[8] = (m[0] * m[4] - m[3] * m[1]) / det;
  return true;
}

void face_distances(const double* H, double* face) {
  // vesin-single-build.cpp:258-275
  auto cross = [](const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
  };
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  const double *a = H, *b = H + 3, *c = H + 6;
  double n[3];
  const double* vec[3] = {a, b, c};
  const double* o1[3] = {b, c, a};
  const double* o2[3] = {c, a, b};
  for (int k = 0; k < 3; ++k) {
    cross(o1[k], o2[k], n);
    const double nn = sqrt(dot(n, n));
    face[k] = fabs(dot(n, vec[k])) / nn;
  }
}

struct Plan {
  Box box;
  Grid grid;
  int geo;
  LJParams lj;
  double rc2;
};

// margin that keeps cell adjacency strictly conservative against rounding in the binning
constexpr double kCellMargin = 1.0 + 1.0e-6;

// fine = 1: cells of at least the cutoff (27-cell stencil); fine = 2: half-cutoff cells (125-cell
// stencil, 42 % fewer candidates per atom) for the brick kernels of large systems
void finish_grid(Plan& pl, double rc, long n, bool images, int fine) {
  const double rcm = rc * kCellMargin;
  double ncd[3];
  for (int k = 0; k < 3; ++k) {
    const double width = pl.box.periodic[k] ? pl.box.face[k] : pl.box.ext[k];
    double c = floor((double)fine * width / rcm);
    if (!(c >= 1.0)) c = 1.0;
    if (c > 1024.0) c = 1024.0;
    ncd[k] = c;
  }
  // bound the cell count (memory and empty-cell overhead): about one cell per atom, but never so
  // few that a mostly empty box (a droplet, a particle in vacuum) is left with oversized cells
  const double cap = std::max(262144.0, 2.0 * (double)n);
  double total = ncd[0] * ncd[1] * ncd[2];
  while (total > cap) {
    int kmax = 0;
    for (int k = 1; k < 3; ++k)
      if (ncd[k] > ncd[kmax]) kmax = k;
    if (ncd[kmax] <= 1.0) break;
    ncd[kmax] = floor(ncd[kmax] * 0.9);
    if (ncd[kmax] < 1.0) ncd[kmax] = 1.0;
    total = ncd[0] * ncd[1] * ncd[2];
  }
  for (int k = 0; k < 3; ++k) {
    const int nc = (int)ncd[k];
    pl.grid.nc[k] = nc;
    if (images) {
      // vesin CellList n_search (:1380-1392): cells to look through so no image is missed
      const double width = pl.box.face[k];
      int ns = (int)ceil(rcm * (double)nc / width);
      if (ns < 1) ns = 1;
      pl.grid.ns_lo[k] = pl.grid.ns_hi[k] = ns;
    } else if (pl.box.periodic[k] && fine == 1) {
      // nearest-image LJ: every DISTINCT neighbouring cell once
      pl.grid.ns_lo[k] = nc >= 3 ? 1 : 0;
      pl.grid.ns_hi[k] = nc >= 2 ? 1 : 0;
    } else {
      // fine periodic grids are only kept by the caller when nc >= 2 * fine + 1 (distinct cells);
      // a direction whose cells ended up wider than the cutoff (cell-count bound) needs one layer
      const double width = pl.box.periodic[k] ? pl.box.face[k] : pl.box.ext[k];
      const int ns = (int)ceil(rcm * (double)nc / width);
      pl.grid.ns_lo[k] = pl.grid.ns_hi[k] = std::max(1, std::min(ns, fine));
    }
  }
  pl.grid.ncells = pl.grid.nc[0] * pl.grid.nc[1] * pl.grid.nc[2];
}

}  // namespace

// ================================================================================================
//  launches
// ================================================================================================
namespace {

inline int div_up(long a, int b) { return (int)((a + b - 1) / b); }

inline bool is_eam(int kind) { return kind == RGPOT_B200_CUH2 || kind == RGPOT_B200_EAM_AL; }
// potentials that read atomic numbers and refuse foreign species (CuH2: 29 / 1, FeHe: 26 / 2)
inline bool reads_species(int kind) { return kind == RGPOT_B200_CUH2 || kind == RGPOT_B200_FEHE; }
inline double eam_rcut(int kind) { return kind == RGPOT_B200_EAM_AL ? 7.1 : 6.1; }
// potentials whose small systems run on k_eam_small (list-based block-per-system kernel)
inline bool uses_eam_small(int kind) { return is_eam(kind) || kind == RGPOT_B200_FEHE; }

int check_launch(RgpotB200Pot* pot, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pot->last_error = std::string("CUDA launch failure (") + what + "): " + cudaGetErrorString(e);
    return kStatusCuda;
  }
  pot->launches++;
  return 0;
}

#define LAUNCH_CHECK(pot, what)                 \
  do {                                          \
    int _s = check_launch(pot, what);           \
    if (_s) return _s;                          \
  } while (0)

// ---- brick kernels (brick.cuh): plan and launch ------------------------------------------------
inline int round_up(int v, int m) { return ((v + m - 1) / m) * m; }

// FP32 prefilter threshold: a superset of the exact accept test.  A = bound on the staged FP32
// offsets (brick halo relative to the brick origin), so rounding the two positions and their
// difference moves a component by at most delta, and d2 by at most the last factor.
float brick_threshold(const Plan& pl, const int bd[3], const int ns[3], double rc) {
  // staged FP32 offsets are relative to the brick centre: |component| <= A
  double A = 0.0;
  for (int k = 0; k < 3; ++k) {
    double ext = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double comp = pl.box.periodic[d] ? fabs(pl.box.H[3 * d + k]) : (d == k ? pl.box.ext[d] : 0.0);
      ext += (0.5 * bd[d] + ns[d] + 0.01) * comp / pl.grid.nc[d];
    }
    A = std::max(A, ext);
  }
  // The search works in FP16 (unit roundoff u = 2^-11): rounding the two offsets moves a component
  // by <= 2 A u, the subtraction rounds once more (|d| <= rc + ...), so a component of the computed
  // separation is off by <= delta; the three squares and two sums round partial results of
  // magnitude <= ~1.2 rc^2.  The bound on the computed d2 of an accepted pair follows (half again
  // for safety); every listed candidate still takes the exact FP64 test.
  const double u = ldexp(1.0, -11);
  const double delta = 2.0 * A * u * 1.001 + (rc + 4.0 * A * u) * u;
  const double err = 2.0 * sqrt(3.0) * rc * delta + 3.0 * delta * delta + 6.0 * u * 1.2 * rc * rc;
  const double t = (rc * rc + 1.5 * err) * (1.0 + 1.0e-6);
  return (float)t;  // rounded up to FP16 in the kernel
}

bool plan_bricks(const RgpotB200Pot* pot, const Plan& pl, long n, double rc, int words, BrickPlan& bp) {
  const Grid& g = pl.grid;
  int ns[3];
  for (int k = 0; k < 3; ++k) {
    if (g.ns_lo[k] != g.ns_hi[k] || g.ns_lo[k] < 1 || g.ns_lo[k] > 2) return false;
    // a halo cell is at most one image away (the kernels fold an atom's own wrap count of +-1 into
    // the 5 x 5 x 5 image codes)
    if (pl.box.periodic[k] && g.nc[k] < g.ns_lo[k]) return false;
    ns[k] = g.ns_lo[k];
  }
  const double mean = (double)n / (double)g.ncells;
  // expected neighbours per atom from the mean number density (cells tile the occupied volume)
  double cellvol = 1.0;
  {
    double cv[3][3];
    for (int d = 0; d < 3; ++d)
      for (int k = 0; k < 3; ++k)
        cv[d][k] = (pl.box.periodic[d] ? pl.box.H[3 * d + k] : (d == k ? pl.box.ext[d] : 0.0)) / g.nc[d];
    cellvol = fabs(cv[0][0] * (cv[1][1] * cv[2][2] - cv[1][2] * cv[2][1]) -
                   cv[0][1] * (cv[1][0] * cv[2][2] - cv[1][2] * cv[2][0]) +
                   cv[0][2] * (cv[1][0] * cv[2][1] - cv[1][1] * cv[2][0]));
  }
  const double nnb = mean / cellvol * (4.0 / 3.0) * M_PI * rc * rc * rc + 1.0;
  int tpa = pot->opt_tpa ? pot->opt_tpa : (nnb > 64.0 ? 4 : 2);  // long lists: more threads per atom
  if (tpa != 1 && tpa != 2 && tpa != 4) tpa = 2;
  int bd[3] = {1, 1, 1};
  if (pot->opt_brick[0] > 0) {
    for (int k = 0; k < 3; ++k) bd[k] = std::max(1, std::min({pot->opt_brick[k], g.nc[k], 4}));
  } else {
    // the largest brick (up to 4 cells a side) shared memory allows: staging and table work per
    // atom shrink with the brick; a brick with more atoms than threads / tpa takes several passes
    for (int k = 0; k < 3; ++k) bd[k] = g.nc[k] >= 4 ? 4 : (g.nc[k] >= 2 ? 2 : 1);
  }
  for (;;) {
    const int ncell = bd[0] * bd[1] * bd[2];
    const int nhalo = (bd[0] + 2 * ns[0]) * (bd[1] + 2 * ns[1]) * (bd[2] + 2 * ns[2]);
    if (ncell <= kBrickMaxCells && nhalo <= kBrickMaxHalo) {
      const double pop = mean * ncell;
      int threads = pot->opt_threads ? pot->opt_threads : round_up((int)ceil(pop * 1.15 * tpa), 32);
      threads = std::max(64, std::min(kBrickMaxThreads, round_up(threads, 32)));
      // capacities: the staged halo from the mean density (or from what the last call with this
      // signature met); the private lists as long as they can be without costing a resident block
      // (a full list only means one more search / physics round, a full halo a cut brick)
      {
        const double fcc = pot->opt_slack ? 1.0 + 0.01 * pot->opt_slack : 1.2;
        bp.cap_cand = std::min(65504, round_up((int)(nhalo * mean * fcc) + 32, 32));
        const auto& m = pot->memo;
        const int seen = (!pot->opt_slack && m.matches(n, g.nc)) ? m.lookup(bd) : 0;
        const bool known = seen > 0;
        if (known) bp.cap_cand = std::min(65504, round_up((int)(seen * 1.04) + 24, 32));
        const int want = (int)(nnb / tpa * 2.2) + 8, least = std::max(24, (int)(nnb / tpa) + 14);
        auto blocks_per_sm = [&](int cap_list) {
          const size_t per_block = brick_smem_bytes(bp.cap_cand, cap_list, threads, words) + 8192 + 1024;
          return (int)(pot->smem_per_sm / per_block);
        };
        bp.cap_list = least;
        if (pot->opt_list) {
          bp.cap_list = pot->opt_list;
        } else {
          const int by_regs = std::max(1, 768 / threads);  // 85 registers per thread
          const int blocks = std::min(by_regs, blocks_per_sm(least));
          for (int c = want; c > least; c -= 2) {
            if (std::min(by_regs, blocks_per_sm(c)) >= blocks) {
              bp.cap_list = c;
              break;
            }
          }
        }
        bp.smem = brick_smem_bytes(bp.cap_cand, bp.cap_list, threads, words);
      }
      if ((size_t)bp.smem + 10240 <= pot->smem_optin) {
        for (int k = 0; k < 3; ++k) {
          bp.bd[k] = bd[k];
          bp.ns[k] = ns[k];
          bp.nb[k] = (g.nc[k] + bd[k] - 1) / bd[k];
        }
        bp.nbricks = bp.nb[0] * bp.nb[1] * bp.nb[2];
        {
          auto& mp = const_cast<RgpotB200Pot*>(pot)->memo_pending;
          mp.n = n;
          for (int k = 0; k < 3; ++k) {
            mp.nc[k] = g.nc[k];
            mp.bd[k] = bd[k];
          }
          mp.max_halo = 0;
          const_cast<RgpotB200Pot*>(pot)->last_cap_cand = bp.cap_cand;
          const_cast<RgpotB200Pot*>(pot)->last_cap_list = bp.cap_list;
        }
        bp.tpa = tpa;
        bp.threads = threads;
        bp.thr2f = brick_threshold(pl, bd, ns, rc);
        return true;
      }
    }
    // too large for shared memory: shrink the brick, largest direction first
    int kmax = 2;
    for (int k = 1; k >= 0; --k)
      if (bd[k] > bd[kmax]) kmax = k;
    if (bd[kmax] == 1) return false;
    bd[kmax] = bd[kmax] == 4 ? 2 : 1;
  }
}

template <int POT>
int launch_brick(RgpotB200Pot* pot, const GatherArgs& ga, const BrickPlan& bp, const TileOut& to,
                 cudaStream_t st, const char* what) {
  const unsigned smem = brick_smem_bytes(bp.cap_cand, bp.cap_list, bp.threads, tile_words(POT));
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<bp.nbricks, bp.threads, smem, st>>>(ga, bp, to);
  };
  if (bp.tpa == 1)
    go(k_brick<POT, 1>);
  else if (bp.tpa == 4)
    go(k_brick<POT, 4>);
  else
    go(k_brick<POT, 2>);
  return check_launch(pot, what);
}

// ---- tile kernels (tile.cu): plan and launch ---------------------------------------------------
// Threshold of the tensor-core prefilter: a superset of the exact accept test.  Staged positions are
// FP16 offsets from the box centre, |component| <= A, each rounded once (error <= A 2^-11), so a
// component of the separation is off by <= delta = 2 A 2^-11 and |d~|^2 <= (|d| + sqrt(3) delta)^2;
// the products are exact in FP32, the seven-term accumulation rounds partial sums of magnitude
// <= ~8 A^2 (slack term).  Every kept candidate still takes the exact FP64 test.
float tile_threshold(const Plan& pl, const int bd[3], const int ns[3], double rc) {
  double A = 0.0;
  for (int k = 0; k < 3; ++k) {
    double ext = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double comp = pl.box.periodic[d] ? fabs(pl.box.H[3 * d + k]) : (d == k ? pl.box.ext[d] : 0.0);
      ext += (0.5 * bd[d] + ns[d] + 0.01) * comp / pl.grid.nc[d];
    }
    A = std::max(A, ext);
  }
  const double delta = 2.0 * A * ldexp(1.0, -11) * 1.001;
  const double slack = 16.0 * (8.0 * A * A + rc * rc) * ldexp(1.0, -24) + 1.0e-4 * rc * rc * 1.0e-2;
  const double t = rc * rc + 2.0 * sqrt(3.0) * rc * delta + 3.0 * delta * delta + slack;
  return (float)(t * (1.0 + 1.0e-6));
}

bool plan_tiles(RgpotB200Pot* pot, const Plan& pl, long n, double rc, int words, TilePlan& tp) {
  const Grid& g = pl.grid;
  for (int k = 0; k < 3; ++k) {
    if (g.ns_lo[k] != g.ns_hi[k] || g.ns_lo[k] < 1 || g.ns_lo[k] > 2) return false;
    // a halo cell is at most one image away (an atom's own wrap count of +-1 is folded into the
    // 5 x 5 x 5 image codes)
    if (pl.box.periodic[k] && g.nc[k] < g.ns_lo[k]) return false;
    tp.ns[k] = g.ns_lo[k];
  }
  const double mean = (double)n / (double)g.ncells;
  // cells per group: the shape that tests the fewest candidates per atom with passes of sixteen
  int best[3] = {1, 1, 1};
  if (pot->opt_tile_group > 0) {
    best[0] = std::max(1, std::min({(pot->opt_tile_group / 100) % 10, g.nc[0], 4}));
    best[1] = std::max(1, std::min({(pot->opt_tile_group / 10) % 10, g.nc[1], 4}));
    best[2] = std::max(1, std::min({pot->opt_tile_group % 10, g.nc[2], 4}));
    while (best[0] * best[1] * best[2] > kTileMaxGroupCells) best[2] = std::max(1, best[2] - 1);
  } else {
    double best_cost = 1e300;
    for (int gz = 1; gz <= std::min(4, g.nc[2]); ++gz)
      for (int gy = 1; gy <= std::min(4, g.nc[1]); ++gy)
        for (int gx = 1; gx <= std::min(4, g.nc[0]); ++gx) {
          if (gx * gy * gz > kTileMaxGroupCells) continue;
          const double m = mean * gx * gy * gz;
          const double passes = ceil(std::max(m * 1.1, 1.0) / 16.0);
          const double window = mean * (gx + 2 * tp.ns[0]) * (gy + 2 * tp.ns[1]) * (gz + 2 * tp.ns[2]);
          // per atom: tensor-core tiles of the search plus the physics' idle quads when a pass is short
          const double fill = std::min(1.0, m / (16.0 * passes));
          const double cost = passes * window / std::max(m, 1e-9) + 120.0 / std::max(fill, 0.05);
          if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best[0] = gx;
            best[1] = gy;
            best[2] = gz;
          }
        }
  }
  for (int k = 0; k < 3; ++k) tp.gd[k] = best[k];
  // cells per brick: two groups a side (eight groups: one per warp), fewer where the grid is narrow
  int bd[3];
  for (int k = 0; k < 3; ++k) {
    int want = 2 * tp.gd[k];
    if (pot->opt_tile_brick > 0) {
      const int v = k == 0 ? (pot->opt_tile_brick / 100) % 10 : (k == 1 ? (pot->opt_tile_brick / 10) % 10 : pot->opt_tile_brick % 10);
      want = std::max(tp.gd[k], (v / tp.gd[k]) * tp.gd[k]);
    }
    bd[k] = std::min(want, ((g.nc[k] + tp.gd[k] - 1) / tp.gd[k]) * tp.gd[k]);
  }
  for (;;) {
    const int nhalo = (bd[0] + 2 * tp.ns[0]) * (bd[1] + 2 * tp.ns[1]) * (bd[2] + 2 * tp.ns[2]);
    if (nhalo <= kTileMaxHalo) {
      const double fcc = 1.2;
      int cap = std::min(65000, round_up((int)(nhalo * mean * fcc) + 32, 8));
      const auto& m = pot->memo;
      const int seen = m.matches(n, g.nc) ? m.lookup(bd) : 0;
      if (seen > 0) cap = std::min(65000, round_up((int)(seen * 1.04) + 24, 8));
      if (pot->opt_tile_cap > 0) cap = round_up(pot->opt_tile_cap, 8);
      const double window = mean * (tp.gd[0] + 2 * tp.ns[0]) * (tp.gd[1] + 2 * tp.ns[1]) * (tp.gd[2] + 2 * tp.ns[2]);
      int wcap = std::max(128, std::min(1024, round_up((int)(window * 1.25) + 16, 64)));
      if (seen > 0) {  // the densest halo met last time, scaled to a window: dense regions of an uneven system
        const double wfrac = (double)((tp.gd[0] + 2 * tp.ns[0]) * (tp.gd[1] + 2 * tp.ns[1]) * (tp.gd[2] + 2 * tp.ns[2])) / (double)nhalo;
        wcap = std::max(wcap, std::min(1024, round_up((int)(seen * wfrac * 1.1) + 16, 64)));
      }
      if (pot->opt_tile_wcap > 0) wcap = std::max(64, round_up(pot->opt_tile_wcap, 64));
      const unsigned smem = tile_smem_bytes(cap, wcap, kTileThreads, words);
      if ((size_t)smem + 12 * 1024 <= pot->smem_optin) {
        for (int k = 0; k < 3; ++k) {
          tp.bd[k] = bd[k];
          tp.nb[k] = (g.nc[k] + bd[k] - 1) / bd[k];
        }
        tp.nbricks = tp.nb[0] * tp.nb[1] * tp.nb[2];
        tp.cap_cand = cap;
        tp.wcap = wcap;
        tp.threads = kTileThreads;
        tp.smem = smem;
        tp.thr2 = tile_threshold(pl, bd, tp.ns, rc);
        auto& mp = pot->memo_pending;
        mp.n = n;
        for (int k = 0; k < 3; ++k) {
          mp.nc[k] = g.nc[k];
          mp.bd[k] = bd[k];
        }
        mp.max_halo = 0;
        pot->last_cap_cand = cap;
        pot->last_cap_list = wcap;
        return true;
      }
    }
    // too large for shared memory: shrink the brick by one group, largest direction first
    int kmax = -1;
    for (int k = 0; k < 3; ++k)
      if (bd[k] > tp.gd[k] && (kmax < 0 || bd[k] > bd[kmax])) kmax = k;
    if (kmax < 0) return false;
    bd[kmax] -= tp.gd[kmax];
  }
}

// lower bound on the 16-atom passes of a grid: one per non-empty group at least
unsigned long long g_min_passes(const Grid& g, const TilePlan& tp) {
  unsigned long long ng = 1;
  for (int k = 0; k < 3; ++k) ng *= (unsigned long long)((g.nc[k] + tp.gd[k] - 1) / tp.gd[k]);
  return ng;
}

int launch_tile(RgpotB200Pot* pot, int tile_pot, const GatherArgs& ga, const TilePlan& tp, const TileOut& to,
                cudaStream_t st, const char* what) {
  pot->last_used_tile = 1;
  const cudaError_t e = rb::tile_launch(tile_pot, ga, tp, to, st);
  if (e != cudaSuccess) {
    pot->last_error = std::string("CUDA launch failure (") + what + "): " + cudaGetErrorString(e);
    return kStatusCuda;
  }
  pot->launches++;
  return 0;
}

// fixed-order fold of the energy partials: long arrays in two stages
int reduce_energy(RgpotB200Pot* pot, double* d_part, int n, double* d_E, cudaStream_t st) {
  if (n > 4096) {
    const int nb = n < 262144 ? 128 : 1024, per = (n + nb - 1) / nb;  // (a function of n only: one fixed order)
    CU_TRY(pot, pot->partials2.reserve(sizeof(double) * nb));
    k_reduce_slices<<<nb, 256, 0, st>>>(d_part, n, per, pot->partials2.as<double>(), pot->status.as<Status>(), d_E);
    LAUNCH_CHECK(pot, "k_reduce_slices");
    return 0;
  } else {
    k_reduce_partials<<<1, 1024, 0, st>>>(d_part, n, d_E);
  }
  LAUNCH_CHECK(pot, "k_reduce_partials");
  return 0;
}

// Cell-list path for ONE system whose positions / Z already sit on the device.
// Leaves forces in d_F (original order), energy in d_E[0], flags in pot->status.
int run_cell_path(RgpotB200Pot* pot, long n, const double* d_pos, const int* d_Z, double* d_F,
                  double* d_E, const double* box, cudaStream_t st, const EmitArgs* emit,
                  bool allow_tile = true, double nlist_rc = 0.0) {
  // nlist_rc > 0: a plain vesin-shaped neighbor list of that cutoff (every image, self images kept),
  // whatever potential the handle was created for (rgpot_b200_neighbor_list)
  const bool nlist = nlist_rc > 0.0 && emit != nullptr;
  pot->last_used_tile = 0;
  if (!pot->opt_bricks) allow_tile = false;
  const int kind = nlist ? RGPOT_B200_LJ : pot->cfg.kind;
  const bool eam = is_eam(kind);
  const bool al = kind == RGPOT_B200_EAM_AL;
  const bool fehe = kind == RGPOT_B200_FEHE;
  const bool images = eam || fehe || pot->cfg.lj_all_images || nlist;
  Plan pl{};
  const double rc = nlist ? nlist_rc : (fehe ? 5.4 : (eam ? eam_rcut(kind) : pot->lj_rc));

  CU_TRY(pot, pot->status.reserve(sizeof(Status)));
  Status init{};
  init.first_bad_atom = INT_MAX;
  CU_TRY(pot, cudaMemcpyAsync(pot->status.p, &init, sizeof init, cudaMemcpyHostToDevice, st));

  if (images) {
    memcpy(pl.box.H, box, 9 * sizeof(double));
    if (!invert_box(pl.box.H, pl.box.Hinv))
      return fail(pot, 1, "rgpot_neighbors: vesin build failed: the box matrix is not invertible");
    face_distances(pl.box.H, pl.box.face);
    for (int k = 0; k < 3; ++k) {
      pl.box.periodic[k] = 1;
      pl.box.lo[k] = 0.0;
      pl.box.ext[k] = 1.0;
    }
  } else {
    // LJ reference semantics: only the box diagonal is read (LJPot.cc:35, PairListCache.hpp:176-184)
    const bool cluster = kind == RGPOT_B200_LJ_CLUSTER || kind == RGPOT_B200_ZBL;
    memset(pl.box.H, 0, sizeof pl.box.H);
    memset(pl.box.Hinv, 0, sizeof pl.box.Hinv);
    bool need_bbox = false;
    for (int k = 0; k < 3; ++k) {
      const double w = box ? box[4 * k] : 0.0;
      const bool per = !cluster && w != 0.0 && std::isfinite(w);
      pl.box.periodic[k] = per ? 1 : 0;
      pl.box.H[4 * k] = per ? w : 1.0;
      pl.box.Hinv[4 * k] = per ? 1.0 / w : 1.0;
      pl.box.face[k] = per ? fabs(w) : 1.0;
      pl.lj.w[k] = cluster ? 0.0 : w;
      pl.lj.inv[k] = per ? 1.0 / w : 0.0;
      if (!per) need_bbox = true;
    }
    for (int k = 0; k < 3; ++k) {
      pl.box.lo[k] = 0.0;
      pl.box.ext[k] = 1.0;
    }
    if (need_bbox) {
      CU_TRY(pot, pot->mm.reserve(6 * sizeof(unsigned long long)));
      unsigned long long initmm[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
      CU_TRY(pot, cudaMemcpyAsync(pot->mm.p, initmm, sizeof initmm, cudaMemcpyHostToDevice, st));
      k_bbox<<<std::min(div_up(n, 256), 4 * pot->sm_count), 256, 0, st>>>(
          d_pos, (int)n, pot->mm.as<unsigned long long>());
      LAUNCH_CHECK(pot, "k_bbox");
      unsigned long long got[6];
      CU_TRY(pot, cudaMemcpyAsync(got, pot->mm.p, sizeof got, cudaMemcpyDeviceToHost, st));
      CU_TRY(pot, cudaStreamSynchronize(st));
      for (int k = 0; k < 3; ++k) {
        if (pl.box.periodic[k]) continue;
        double lo = dkey_inv(got[k]), hi = dkey_inv(got[3 + k]);
        if (!std::isfinite(lo) || !std::isfinite(hi)) {
          lo = 0.0;
          hi = 1.0;
        }
        double ext = (hi - lo) * (1.0 + 1e-9);
        if (!(ext > 1e-6)) ext = 1.0;
        pl.box.lo[k] = lo;
        pl.box.ext[k] = ext;
      }
    }
  }
  pl.lj.u0 = pot->cfg.u0;
  pl.lj.psi2 = pot->cfg.psi * pot->cfg.psi;
  pl.lj.shiftU = pot->lj_shiftU;
  pl.lj.rc2 = rc * rc;
  pl.lj.morse = pot->cfg.kind == RGPOT_B200_MORSE ? 1 : (pot->cfg.kind == RGPOT_B200_ZBL ? 2 : 0);
  pl.lj.zbl_table = pot->zbl_table.as<double>();
  pl.lj.zbl_nt = (int)pot->zbl_species.size();
  pl.lj.zbl_rin = pot->cfg.zbl_cut_inner;
  pl.lj.zbl_rin2 = pot->cfg.zbl_cut_inner * pot->cfg.zbl_cut_inner;
  pl.lj.m_de = pot->cfg.morse_De;
  pl.lj.m_a = pot->cfg.morse_a;
  pl.lj.m_re = pot->cfg.morse_re;
  pl.lj.m_ecut = pot->morse_ecut;
  pl.lj.m_two_de_a = 2.0 * pot->cfg.morse_De * pot->cfg.morse_a;
  pl.rc2 = rc * rc;
  // half-cutoff cells for systems large enough to fill the machine with bricks
  int fine = (allow_tile && !emit && n >= pot->opt_fine_min && pot->opt_fine != 1) ? 2 : 1;
  for (;;) {
    finish_grid(pl, rc, n, images, fine);
    bool ok = true;
    if (fine == 2) {
      for (int k = 0; k < 3; ++k) {
        if (!images && pl.box.periodic[k] && pl.grid.nc[k] < 5) ok = false;
        if (pl.grid.ns_lo[k] > 2 || pl.grid.ns_hi[k] > 2) ok = false;
      }
    }
    if (ok) break;
    fine = 1;
  }
  if (images) {
    pl.geo = GEO_IMAGES;
  } else {
    bool shift_ok = true;
    for (int k = 0; k < 3; ++k)
      if (pl.box.periodic[k] && pl.grid.nc[k] < 2 * fine + 1) shift_ok = false;
    pl.geo = shift_ok ? GEO_LJ_SHIFT : GEO_LJ_FOLD;
  }

  const int ncells = pl.grid.ncells;
  const int nscan = div_up(ncells, kScanBlock * kScanItems);
  CU_TRY(pot, pot->cell_of.reserve(sizeof(int) * n));
  CU_TRY(pot, pot->rank.reserve(sizeof(int) * n));
  CU_TRY(pot, pot->ashift.reserve(sizeof(int) * n));
  CU_TRY(pot, pot->counts.reserve(sizeof(int) * ((size_t)ncells + 1)));
  CU_TRY(pot, pot->starts.reserve(sizeof(int) * ((size_t)ncells + 1)));
  CU_TRY(pot, pot->blocksums.reserve(sizeof(int) * ((size_t)nscan + 1)));
  CU_TRY(pot, pot->order.reserve(sizeof(int) * n));
  CU_TRY(pot, pot->spos.reserve(sizeof(double4) * n));
  CU_TRY(pot, pot->scell.reserve(sizeof(int) * n));
  int nblk = div_up(n, 128);
  // the gather kernels as a stand-by behind the brick kernels (they return at once unless an atom lies
  // far outside the cell): a few blocks per SM, grid-stride
  const int nblk_standby = std::min(nblk, 8 * pot->sm_count);
  CU_TRY(pot, pot->partials.reserve(sizeof(double) * n));  // per-atom energies, sorted order
  if (eam || fehe) {
    CU_TRY(pot, pot->dfd.reserve(sizeof(double) * n));
    CU_TRY(pot, pot->eemb.reserve(sizeof(double) * n));
  }
  if (fehe) CU_TRY(pot, pot->dfs.reserve(sizeof(double) * n));

  Status* d_st = pot->status.as<Status>();
  CU_TRY(pot, cudaMemsetAsync(pot->counts.p, 0, sizeof(int) * ((size_t)ncells + 1), st));
  k_bin<<<div_up(n, 256), 256, 0, st>>>(d_pos, (int)n, pl.box, pl.grid, pot->cell_of.as<int>(),
                                        pot->rank.as<int>(), pot->ashift.as<int>(),
                                        pot->counts.as<int>(), d_st);
  LAUNCH_CHECK(pot, "k_bin");
  k_scan_partial<<<nscan, kScanBlock, 0, st>>>(pot->counts.as<int>(), ncells, pot->blocksums.as<int>());
  LAUNCH_CHECK(pot, "k_scan_partial");
  k_scan_final<<<nscan, kScanBlock, 0, st>>>(pot->counts.as<int>(), ncells, pot->blocksums.as<int>(),
                                             pot->starts.as<int>());
  LAUNCH_CHECK(pot, "k_scan_final");
  k_scatter<<<div_up(n, 256), 256, 0, st>>>(pot->cell_of.as<int>(), pot->rank.as<int>(),
                                            pot->starts.as<int>(), (int)n, pot->order.as<int>());
  LAUNCH_CHECK(pot, "k_scatter");
  k_order_gather<<<div_up(n, 256), 256, 0, st>>>(d_pos, d_Z, pot->starts.as<int>(), pot->cell_of.as<int>(),
                                                 pot->order.as<int>(), pot->ashift.as<int>(), (int)n,
                                                 kind == RGPOT_B200_CUH2 ? 1 : (kind == RGPOT_B200_ZBL ? 2 : (fehe ? 3 : 0)),
                                                 pot->spos.as<double4>(), pot->scell.as<int>(), d_st);
  LAUNCH_CHECK(pot, "k_order_gather");

  GatherArgs ga{};
  ga.spos = pot->spos.as<double4>();
  ga.scell = pot->scell.as<int>();
  ga.starts = pot->starts.as<int>();
  ga.n = (int)n;
  ga.box = pl.box;
  ga.grid = pl.grid;
  ga.lj = pl.lj;
  ga.rc2 = pl.rc2;
  ga.st_in = d_st;
  ga.keep_self_images = nlist ? 1 : 0;

  if (emit) {
    if (pl.geo == GEO_IMAGES)
      k_emit_pairs<GEO_IMAGES><<<nblk, 128, 0, st>>>(ga, *emit);
    else if (pl.geo == GEO_LJ_SHIFT)
      k_emit_pairs<GEO_LJ_SHIFT><<<nblk, 128, 0, st>>>(ga, *emit);
    else
      k_emit_pairs<GEO_LJ_FOLD><<<nblk, 128, 0, st>>>(ga, *emit);
    LAUNCH_CHECK(pot, "k_emit_pairs");
    return 0;
  }

  double* d_part = pot->partials.as<double>();  // every atom's energy is written by exactly one kernel
  if (eam || fehe) {
    BrickPlan bp{};
    TilePlan tpl{};
    if (allow_tile && pot->opt_tile && plan_tiles(pot, pl, n, rc, fehe ? 5 : 4, tpl)) {
      // tile kernels (tile.cu): density + embedding, then forces; the gather kernels stand by for atoms
      // more than one cell vector outside the cell
      TileOut to{d_F, d_part, pot->dfd.as<double>(), pot->eemb.as<double>(), fehe ? pot->dfs.as<double>() : nullptr,
                 d_st, nullptr, nullptr, 0, nullptr};
      if (pot->opt_handoff) {
        // survivor masks handed from the density to the force pass: 32 lanes x (window / 64) words per
        // 16-atom pass; a shortfall only makes some passes search again
        const unsigned long long per_pass = 32ull * (unsigned long long)(tpl.wcap / 64);
        const unsigned long long want = ((unsigned long long)n / 6ull + (unsigned long long)g_min_passes(pl.grid, tpl)) * per_pass + 4096ull;
        const unsigned long long cap = std::min(want, 1ull << 31);
        if (pot->listpool.reserve(sizeof(unsigned) * cap) == cudaSuccess &&
            pot->listref.reserve(sizeof(unsigned long long) * (size_t)n) == cudaSuccess) {
          to.poolw = pot->listpool.as<unsigned>();
          to.lref = pot->listref.as<unsigned long long>();
          to.pool_cap = cap;
        } else {
          cudaGetLastError();
        }
      }
      int ls = launch_tile(pot, fehe ? TILE_FEHE_DENSITY : (al ? TILE_AL_DENSITY : TILE_EAM_DENSITY), ga, tpl, to, st, "k_tile<density>");
      if (ls) return ls;
      ls = launch_tile(pot, fehe ? TILE_FEHE_FORCE : (al ? TILE_AL_FORCE : TILE_EAM_FORCE), ga, tpl, to, st, "k_tile<force>");
      if (ls) return ls;
      ga.only_if_big_shift = 1;
      nblk = nblk_standby;
    } else if (allow_tile && plan_bricks(pot, pl, n, rc, fehe ? 5 : 4, bp)) {
      // brick kernels; the gather kernels cover atoms outside the cell (per-pair image shifts) --
      // each side returns at once when the binning flag is not its own
      TileOut to{d_F, d_part, pot->dfd.as<double>(), pot->eemb.as<double>(), fehe ? pot->dfs.as<double>() : nullptr,
                 d_st, nullptr, nullptr, 0, nullptr};
      if (pot->opt_handoff) {
        // survivor lists handed from the density to the force pass: room for every thread's list at
        // its capacity in half-filled warps (a shortfall only makes some threads search again)
        const unsigned long long want = (unsigned long long)n * bp.tpa * bp.cap_list * 3ull / 2ull + 4096ull;
        const unsigned long long cap = std::min(want, 1ull << 32);
        if (pot->listpool.reserve(sizeof(unsigned short) * cap) == cudaSuccess &&
            pot->listref.reserve(sizeof(unsigned long long) * (size_t)n * bp.tpa) == cudaSuccess) {
          to.pool = pot->listpool.as<unsigned short>();
          to.lref = pot->listref.as<unsigned long long>();
          to.pool_cap = cap;
        } else {
          cudaGetLastError();
        }
      }
      int ls = fehe ? launch_brick<TILE_FEHE_DENSITY>(pot, ga, bp, to, st, "k_brick<fehe density>")
               : al ? launch_brick<TILE_AL_DENSITY>(pot, ga, bp, to, st, "k_brick<al density>")
                    : launch_brick<TILE_EAM_DENSITY>(pot, ga, bp, to, st, "k_brick<eam density>");
      if (ls) return ls;
      ls = fehe ? launch_brick<TILE_FEHE_FORCE>(pot, ga, bp, to, st, "k_brick<fehe force>")
           : al ? launch_brick<TILE_AL_FORCE>(pot, ga, bp, to, st, "k_brick<al force>")
                : launch_brick<TILE_EAM_FORCE>(pot, ga, bp, to, st, "k_brick<eam force>");
      if (ls) return ls;
      ga.only_if_big_shift = 1;
      nblk = nblk_standby;
    }
    if (fehe) {
      // Fe-He on the gather walk (rgpot_fehe.f90:573-620): the general path
      k_fehe_density<<<nblk, 128, 0, st>>>(ga, pot->dfd.as<double>(), pot->dfs.as<double>(), pot->eemb.as<double>());
      LAUNCH_CHECK(pot, "k_fehe_density");
      k_fehe_force<<<nblk, 128, 0, st>>>(ga, pot->dfd.as<double>(), pot->dfs.as<double>(), pot->eemb.as<double>(), d_F,
                                         d_part);
    } else if (al) {
      k_eam_density<EAM_AL><<<nblk, 128, 0, st>>>(ga, pot->dfd.as<double>(), pot->eemb.as<double>(), d_st);
      LAUNCH_CHECK(pot, "k_eam_density");
      k_eam_force<EAM_AL><<<nblk, 128, 0, st>>>(ga, pot->dfd.as<double>(), pot->eemb.as<double>(), d_F, d_part);
    } else {
      k_eam_density<EAM_CUH2><<<nblk, 128, 0, st>>>(ga, pot->dfd.as<double>(), pot->eemb.as<double>(), d_st);
      LAUNCH_CHECK(pot, "k_eam_density");
      k_eam_force<EAM_CUH2><<<nblk, 128, 0, st>>>(ga, pot->dfd.as<double>(), pot->eemb.as<double>(), d_F, d_part);
    }
    LAUNCH_CHECK(pot, "k_eam_force");
    return reduce_energy(pot, d_part, (int)n, d_E, st);
  } else if (pl.geo == GEO_IMAGES) {
    k_lj_gather<GEO_IMAGES><<<nblk, 128, 0, st>>>(ga, d_F, d_part);
    LAUNCH_CHECK(pot, "k_lj_gather<images>");
  } else if (TilePlan tpl{}; pl.geo == GEO_LJ_SHIFT && allow_tile && pot->opt_tile && plan_tiles(pot, pl, n, rc, 3, tpl)) {
    // tile kernel (tile.cu); the literal-fold gather kernel covers atoms far outside the cell
    TileOut to{d_F, d_part, nullptr, nullptr, nullptr, d_st, nullptr, nullptr, 0, nullptr};
    int ls = launch_tile(pot, pl.lj.morse ? TILE_PAIR : TILE_LJ, ga, tpl, to, st, "k_tile<lj>");
    if (ls) return ls;
    ga.only_if_big_shift = 1;
    k_lj_gather<GEO_LJ_SHIFT><<<nblk_standby, 128, 0, st>>>(ga, d_F, d_part);
    LAUNCH_CHECK(pot, "k_lj_gather<shift>");
    return reduce_energy(pot, d_part, (int)n, d_E, st);
  } else if (BrickPlan bp{}; pl.geo == GEO_LJ_SHIFT && allow_tile && plan_bricks(pot, pl, n, rc, 3, bp)) {
    // brick kernel; the literal-fold gather kernel covers the (rare) case of atoms outside the
    // cell -- each of the two returns at once when the binning flag is not its own
    TileOut to{d_F, d_part, nullptr, nullptr, nullptr, d_st, nullptr, nullptr, 0, nullptr};
    int ls = pl.lj.morse ? launch_brick<TILE_PAIR>(pot, ga, bp, to, st, "k_brick<pair>")
                         : launch_brick<TILE_LJ>(pot, ga, bp, to, st, "k_brick<lj>");
    if (ls) return ls;
    ga.only_if_big_shift = 1;
    k_lj_gather<GEO_LJ_SHIFT><<<nblk_standby, 128, 0, st>>>(ga, d_F, d_part);
    LAUNCH_CHECK(pot, "k_lj_gather<shift>");
    return reduce_energy(pot, d_part, (int)n, d_E, st);
  } else if (pl.geo == GEO_LJ_SHIFT) {
    k_lj_gather<GEO_LJ_SHIFT><<<nblk, 128, 0, st>>>(ga, d_F, d_part);
    LAUNCH_CHECK(pot, "k_lj_gather<shift>");
  } else {
    k_lj_gather<GEO_LJ_FOLD><<<nblk, 128, 0, st>>>(ga, d_F, d_part);
    LAUNCH_CHECK(pot, "k_lj_gather<fold>");
  }
  return reduce_energy(pot, d_part, (int)n, d_E, st);
}

// shared memory a k_eam_small block needs
int eam_small_tpa(long nmax) {  // lanes per atom: fill about 448 threads
  int tpa = 1;
  while (tpa < 8 && nmax * (tpa * 2) <= 448) tpa *= 2;
  return tpa;
}

size_t eam_small_smem(long n, int maxnb) {  // must match the layout in k_eam_small
  const int tpa = eam_small_tpa(n);
  const size_t ntl = (size_t)n * tpa, lcap = (size_t)(maxnb + tpa - 1) / tpa + 4;
  size_t bytes = sizeof(double) * (10 * (size_t)n + 81);
  bytes = (bytes + 15) & ~(size_t)15;
  return bytes + sizeof(float) * 4 * (size_t)n + sizeof(int) * 4 * (size_t)n +
         sizeof(unsigned short) * (((ntl + 1) & ~(size_t)1) + lcap * ntl) + 16;
}

// pick the list capacity for a batch whose largest system has nmax atoms (0 = does not fit)
int eam_small_maxnb(const RgpotB200Pot* pot, long nmax) {
  if (nmax < 1) return 96;
  if (nmax > 65535) return 0;
  // two resident blocks per SM when possible, else the roomiest list that still fits one block
  const size_t per_sm = 227 * 1024;
  for (int maxnb : {160, 128, 96}) {
    if (eam_small_smem(nmax, maxnb) + 10 * 1024 <= per_sm / 2 - 1024) return maxnb;  // (+ the kernel's static tables)
  }
  for (int maxnb : {160, 128, 96}) {
    if (eam_small_smem(nmax, maxnb) + 10 * 1024 <= pot->smem_optin) return maxnb;
  }
  return 0;
}

// Launch the batched all-pairs kernel over nsys systems (device-resident, packed).
int launch_small(RgpotB200Pot* pot, size_t nsys, long nmax, const long long* d_off,
                 const double* d_boxes, const double* d_pos, const int* d_Z, double* d_F,
                 double* d_E, unsigned* d_flags, int* d_bad, int maxnb, cudaStream_t st,
                 const EmitArgs* emit, bool z_shared = false, int single_n = 0, int split = 0,
                 double* e_parts = nullptr, int* ticket = nullptr, int lanes = 0, const double* host_box = nullptr,
                 double* f_stage = nullptr) {
  SmallArgs a{};
  a.z_shared = z_shared ? 1 : 0;
  a.single_n = single_n;
  a.split = split;
  a.lanes = lanes;
  a.e_parts = e_parts;
  a.ticket = ticket;
  a.f_stage = f_stage;
  a.half = single_n > 0 ? 0 : pot->opt_half;  // (one system on one block is a latency case: the extra barriers of the half-list mode cost it a microsecond)
  if (host_box && uses_eam_small(pot->cfg.kind)) {  // one system: invert the cell here, not on one GPU thread
    const double rc = pot->cfg.kind == RGPOT_B200_FEHE ? 5.4 : eam_rcut(pot->cfg.kind);
    small_box_init(host_box, rc, a.sb);
    if (!a.sb.bad) small_boxf_init(a.sb, rc, a.sbf);
    a.box_ready = 1;
  }
  a.pos = d_pos;
  a.Z = d_Z;
  a.off = d_off;
  a.boxes = d_boxes;
  a.F = d_F;
  a.E = d_E;
  a.sys_flags = d_flags;
  a.sys_bad = d_bad;
  a.nsys = (int)nsys;
  a.maxnb = maxnb;
  a.cluster = pot->cfg.kind == RGPOT_B200_LJ_CLUSTER || pot->cfg.kind == RGPOT_B200_ZBL;
  a.u0 = pot->cfg.u0;
  a.psi2 = pot->cfg.psi * pot->cfg.psi;
  a.shiftU = pot->lj_shiftU;
  a.cutoff = pot->lj_rc;
  a.morse = pot->cfg.kind == RGPOT_B200_MORSE ? 1 : (pot->cfg.kind == RGPOT_B200_ZBL ? 2 : 0);
  a.zbl_table = pot->zbl_table.as<double>();
  a.zbl_nt = (int)pot->zbl_species.size();
  a.zbl_rin = pot->cfg.zbl_cut_inner;
  a.m_de = pot->cfg.morse_De;
  a.m_a = pot->cfg.morse_a;
  a.m_re = pot->cfg.morse_re;
  a.m_ecut = pot->morse_ecut;
  if (emit) a.emit = *emit;
  if (uses_eam_small(pot->cfg.kind)) {
    const bool al = pot->cfg.kind == RGPOT_B200_EAM_AL;
    const bool fehe = pot->cfg.kind == RGPOT_B200_FEHE;
    const size_t smem = eam_small_smem(nmax, maxnb);
    const int tpa = eam_small_tpa(nmax);
    int threads = (int)std::min<long>(448, ((nmax * tpa + 31) / 32) * 32);
    if (threads < 32) threads = 32;
    const int attr_bit = 1 << ((tpa == 1 ? 0 : tpa == 2 ? 1 : tpa == 4 ? 2 : 3) * 3 + (al ? 1 : (fehe ? 2 : 0)));
    auto launch = [&](auto kern) -> int {
      if (!(pot->small_attr_set & attr_bit)) {  // once per handle and kernel: the calls cost microseconds
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        if (e == cudaSuccess)
          e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(pot->smem_optin - fa.sharedSizeBytes));
        if (e != cudaSuccess) {
          pot->last_error = std::string("CUDA failure: ") + cudaGetErrorString(e);
          return kStatusCuda;
        }
        pot->small_attr_set |= attr_bit;
      }
      kern<<<(unsigned)nsys, threads, smem, st>>>(a);
      return check_launch(pot, "k_eam_small");
    };
    switch (tpa) {
      case 1: return fehe ? launch(k_eam_small<1, EAM_FEHE>) : al ? launch(k_eam_small<1, EAM_AL>) : launch(k_eam_small<1, EAM_CUH2>);
      case 2: return fehe ? launch(k_eam_small<2, EAM_FEHE>) : al ? launch(k_eam_small<2, EAM_AL>) : launch(k_eam_small<2, EAM_CUH2>);
      case 4: return fehe ? launch(k_eam_small<4, EAM_FEHE>) : al ? launch(k_eam_small<4, EAM_AL>) : launch(k_eam_small<4, EAM_CUH2>);
      default: return fehe ? launch(k_eam_small<8, EAM_FEHE>) : al ? launch(k_eam_small<8, EAM_AL>) : launch(k_eam_small<8, EAM_CUH2>);
    }
  }
  const size_t smem = sizeof(double) * 3 * (size_t)nmax + sizeof(int) * (size_t)nmax + 16;
  if (!(pot->small_attr_set & (1 << 30))) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, k_lj_small);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k_lj_small, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(pot->smem_optin - fa.sharedSizeBytes));
    if (e != cudaSuccess) return fail(pot, kStatusCuda, std::string("CUDA failure: ") + cudaGetErrorString(e));
    pot->small_attr_set |= 1 << 30;
  }
  int threads = (int)std::min<long>(256, ((nmax + 31) / 32) * 32);
  if (threads < 32) threads = 32;
  if (split > 0) {  // one system: `lanes` lanes per atom, spread over `split` blocks
    const int t1 = (int)std::min<long>(256, (((long)single_n * lanes + 31) / 32) * 32);
    k_lj_small<<<(unsigned)split, split > 1 ? 256 : t1, smem, st>>>(a);
  } else {
    k_lj_small<<<(unsigned)nsys, threads, smem, st>>>(a);
  }
  return check_launch(pot, "k_lj_small");
}

// can a system of n atoms run on the all-pairs kernels?
bool small_eligible(const RgpotB200Pot* pot, long n) {
  if (uses_eam_small(pot->cfg.kind)) return n <= 512 && eam_small_maxnb(pot, n) > 0;
  if (pot->cfg.lj_all_images) return false;  // opt-in mode lives on the cell path only
  return n <= 2048;
}

// status word -> ABI status + message.  Precedence follows rgpot_cuh2.f90:412-429:
// classify (2) before the table build (1) before the coincidence guard (3).
int map_flags(RgpotB200Pot* pot, unsigned flags, int bad_atom, int bad_z, const double* host_pos) {
  const bool eam = is_eam(pot->cfg.kind);
  if (pot->cfg.kind == RGPOT_B200_FEHE && (flags & kFlagForeignZ)) {
    char buf[256];  // rgpot_fehe.f90:558-563
    snprintf(buf, sizeof buf,
             "rgpot_fehe: unsupported atomic number %d on atom %d; this potential covers iron (26) and helium (2) only",
             bad_z, bad_atom + 1);
    return fail(pot, 2, buf);
  }
  if (pot->cfg.kind == RGPOT_B200_CUH2 && (flags & kFlagForeignZ)) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "rgpot_cuh2: atom %d has atomic number %d; this potential covers Cu (29) and H (1) only",
             bad_atom + 1, bad_z);
    return fail(pot, 2, buf);
  }
  const bool images = eam || pot->cfg.kind == RGPOT_B200_FEHE || pot->cfg.lj_all_images;
  if (images && (flags & kFlagSingularBox))
    return fail(pot, 1, "rgpot_neighbors: vesin build failed: the box matrix is not invertible");
  if (images && (flags & kFlagNonFinite)) {
    char buf[256];
    if (host_pos && bad_atom != INT_MAX && bad_atom >= 0) {
      int axis = 0;
      for (int k = 0; k < 3; ++k)
        if (!std::isfinite(host_pos[3 * (size_t)bad_atom + k])) {
          axis = k;
          break;
        }
      snprintf(buf, sizeof buf,
               "rgpot_neighbors: vesin build failed: point %d has non-finite coordinate along axis %d: %f",
               bad_atom, axis, host_pos[3 * (size_t)bad_atom + axis]);
    } else {
      snprintf(buf, sizeof buf,
               "rgpot_neighbors: vesin build failed: non-finite coordinate or singular box");
    }
    return fail(pot, 1, buf);
  }
  if (flags & kFlagShiftRange)
    return fail(pot, kStatusRange, "rgpot_b200: an atom lies more than 500 cell widths outside the box");
  // rgpot_cuh2.f90:422-428.  The aluminium kernel has no such guard upstream (it would return NaN
  // forces): refusing the input there is a documented deviation, under its own name.
  if (eam && (flags & kFlagCoincident))
    return fail(pot, 3, pot->cfg.kind == RGPOT_B200_EAM_AL ? "rgpot_eam_al: two atoms occupy the same position"
                                                          : "rgpot_cuh2: two atoms occupy the same position");
  return 0;
}

// ---- ZBL species tables (ZBLPot.cc:38-72 radial functions, :88-129 buildTables) ------------------
constexpr int kStatusSpecies = 103;
struct ZblCoeffs {
  double d1, d2, d3, d4, zze, sw1, sw2, sw3, sw4, sw5;
};
double zbl_e(const ZblCoeffs& p, double r) {
  const double r_inv = 1.0 / r;
  const double sum = 0.18175 * exp(-p.d1 * r) + 0.50986 * exp(-p.d2 * r) + 0.28022 * exp(-p.d3 * r) +
                     0.02817 * exp(-p.d4 * r);
  return p.zze * sum * r_inv;
}
double zbl_de(const ZblCoeffs& p, double r) {
  const double r_inv = 1.0 / r;
  const double e1 = exp(-p.d1 * r), e2 = exp(-p.d2 * r), e3 = exp(-p.d3 * r), e4 = exp(-p.d4 * r);
  const double sum = 0.18175 * e1 + 0.50986 * e2 + 0.28022 * e3 + 0.02817 * e4;
  const double sum_p = -0.18175 * p.d1 * e1 - 0.50986 * p.d2 * e2 - 0.28022 * p.d3 * e3 - 0.02817 * p.d4 * e4;
  return p.zze * (sum_p - sum * r_inv) * r_inv;
}
double zbl_d2e(const ZblCoeffs& p, double r) {
  const double r_inv = 1.0 / r;
  const double e1 = exp(-p.d1 * r), e2 = exp(-p.d2 * r), e3 = exp(-p.d3 * r), e4 = exp(-p.d4 * r);
  const double sum = 0.18175 * e1 + 0.50986 * e2 + 0.28022 * e3 + 0.02817 * e4;
  const double sum_p = 0.18175 * e1 * p.d1 + 0.50986 * e2 * p.d2 + 0.28022 * e3 * p.d3 + 0.02817 * e4 * p.d4;
  const double sum_pp = 0.18175 * e1 * p.d1 * p.d1 + 0.50986 * e2 * p.d2 * p.d2 + 0.28022 * e3 * p.d3 * p.d3 +
                        0.02817 * e4 * p.d4 * p.d4;
  return p.zze * (sum_pp + 2.0 * sum_p * r_inv + 2.0 * sum * r_inv * r_inv) * r_inv;
}

// Map atomic numbers to type indices (sorted unique, ZBLPot.cc:186-196) and make sure the device
// table matches the species set.  `types` receives one index per atom.
int zbl_table_for(RgpotB200Pot* pot, const std::vector<int>& uniq, cudaStream_t st);

int zbl_prepare(RgpotB200Pot* pot, size_t n, const int* Z, int* types, cudaStream_t st) {
  std::vector<int> uniq(Z, Z + n);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  if (uniq.size() > 16)
    return fail(pot, kStatusSpecies, "rgpot_b200: ZBL handles at most 16 distinct atomic numbers per call");
  for (size_t k = 0; k < n; ++k)
    types[k] = (int)(std::lower_bound(uniq.begin(), uniq.end(), Z[k]) - uniq.begin());
  return zbl_table_for(pot, uniq, st);
}

// Device-resident atomic numbers: the species set is read off a 288-bit presence bitmap (one small
// D2H instead of a round trip of the whole array) and the atoms are mapped to type indices on the
// device.  Leaves the type indices in pot->Z.
int zbl_prepare_device(RgpotB200Pot* pot, size_t n, const int* d_Z, cudaStream_t st) {
  CU_TRY(pot, pot->zbl_bitmap.reserve(sizeof(unsigned) * 9 + sizeof(int) * 256));
  unsigned* d_bits = pot->zbl_bitmap.as<unsigned>();
  int* d_lut = reinterpret_cast<int*>(d_bits + 9);
  CU_TRY(pot, cudaMemsetAsync(d_bits, 0, sizeof(unsigned) * 9, st));
  k_species_present<<<std::min(div_up((long)n, 256), 8 * pot->sm_count), 256, 0, st>>>(d_Z, (int)n, d_bits);
  LAUNCH_CHECK(pot, "k_species_present");
  unsigned bits[9];
  CU_TRY(pot, cudaMemcpyAsync(bits, d_bits, sizeof bits, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaStreamSynchronize(st));
  std::vector<int> uniq;
  for (int z = 0; z < 256; ++z)
    if ((bits[z >> 5] >> (z & 31)) & 1u) uniq.push_back(z);
  if (bits[8] || uniq.size() > 16)
    return fail(pot, kStatusSpecies, bits[8] ? "rgpot_b200: ZBL atomic numbers must lie in 0..255"
                                            : "rgpot_b200: ZBL handles at most 16 distinct atomic numbers per call");
  int s = zbl_table_for(pot, uniq, st);
  if (s) return s;
  int lut[256] = {0};
  for (size_t k = 0; k < uniq.size(); ++k) lut[uniq[k]] = (int)k;
  CU_TRY(pot, cudaMemcpyAsync(d_lut, lut, sizeof lut, cudaMemcpyHostToDevice, st));
  CU_TRY(pot, pot->Z.reserve(sizeof(int) * n));
  k_map_types<<<div_up((long)n, 256), 256, 0, st>>>(d_Z, (int)n, d_lut, pot->Z.as<int>());
  LAUNCH_CHECK(pot, "k_map_types");
  CU_TRY(pot, cudaStreamSynchronize(st));  // (lut lives on this stack frame)
  return 0;
}

int zbl_table_for(RgpotB200Pot* pot, const std::vector<int>& uniq, cudaStream_t st) {
  if (uniq == pot->zbl_species && pot->zbl_table.p) return 0;
  const size_t nt = uniq.size();
  const double cut_inner = pot->cfg.zbl_cut_inner, cut_global = pot->cfg.cutoff;
  const double tc = cut_global - cut_inner;
  std::vector<double> table(10 * nt * nt);
  for (size_t i = 0; i < nt; ++i)
    for (size_t j = i; j < nt; ++j) {
      const double zi = (double)uniq[i], zj = (double)uniq[j];
      const double a_inv = (pow(zi, 0.23) + pow(zj, 0.23)) / 0.46850;
      ZblCoeffs p{};
      p.d1 = 3.19980 * a_inv;
      p.d2 = 0.94229 * a_inv;
      p.d3 = 0.40290 * a_inv;
      p.d4 = 0.20162 * a_inv;
      p.zze = zi * zj * 14.399645;
      const double fc = zbl_e(p, cut_global), fcp = zbl_de(p, cut_global), fcpp = zbl_d2e(p, cut_global);
      const double swa = (-3.0 * fcp + tc * fcpp) / (tc * tc);
      const double swb = (2.0 * fcp - tc * fcpp) / (tc * tc * tc);
      const double swc = -fc + (tc / 2.0) * fcp - (tc * tc / 12.0) * fcpp;
      p.sw1 = swa;
      p.sw2 = swb;
      p.sw3 = swa / 3.0;
      p.sw4 = swb / 4.0;
      p.sw5 = swc;
      memcpy(&table[10 * (i * nt + j)], &p, sizeof p);
      memcpy(&table[10 * (j * nt + i)], &p, sizeof p);
    }
  CU_TRY(pot, cudaStreamSynchronize(st));  // the previous table may still be in use
  CU_TRY(pot, pot->zbl_table.reserve(sizeof(double) * table.size()));
  CU_TRY(pot, cudaMemcpy(pot->zbl_table.p, table.data(), sizeof(double) * table.size(), cudaMemcpyHostToDevice));
  pot->zbl_species = uniq;
  return 0;
}

void lj_setup(RgpotB200Pot* pot) {
  // LJPot.hpp:52-53 ; LJClusterPot.cc:46 (cutoff <= 0 means "no cutoff" for clusters)
  const RgpotB200Config& c = pot->cfg;
  const double a = pow(c.psi / c.cutoff, 6.0);
  pot->lj_shiftU = 4.0 * c.u0 * a * (a - 1.0);
  pot->lj_rc = c.cutoff;
  if (c.kind == RGPOT_B200_LJ_CLUSTER && !(c.cutoff > 0.0)) pot->lj_rc = 1.0e6;
  // MorsePot.hpp:56-61: shift so that U(cutoff) = 0, the same closed form as the pair kernel
  const double d = 1.0 - exp(-c.morse_a * (c.cutoff - c.morse_re));
  pot->morse_ecut = c.morse_De * d * d - c.morse_De;
}

// ---- FP64 peak microbenchmark ------------------------------------------------------------------
// Eight independent chains per thread, sixteen rounds per trip: 128 DFMA per loop branch, so the loop's
// own integer instructions are ~2 % of the issue slots (an FP64 instruction holds the dispatch port for
// two cycles: every other instruction comes out of the FP64 pipe's share).  ncu: FP64 pipe >= 95 % busy.
__global__ void k_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i += 16) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace

// ================================================================================================
//  C ABI
// ================================================================================================
extern "C" {

int rgpot_b200_abi_version(void) { return RGPOT_B200_ABI_VERSION; }

int rgpot_b200_available(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major >= 10) ++ok;
  }
  return ok;
}

void rgpot_b200_default_config(int kind, RgpotB200Config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof *cfg);
  cfg->kind = kind;
  cfg->device = -1;
  cfg->u0 = 1.0;
  cfg->cutoff = kind == RGPOT_B200_FEHE ? 5.4 : is_eam(kind) ? eam_rcut(kind) : (kind == RGPOT_B200_MORSE ? 9.5 : (kind == RGPOT_B200_ZBL ? 2.5 : 15.0));
  cfg->zbl_cut_inner = 2.0;  // ZBLPot.hpp:29-32
  cfg->psi = 1.0;
  cfg->morse_De = 0.7102;  // MorsePot.hpp:31-36 (platinum)
  cfg->morse_a = 1.6047;
  cfg->morse_re = 2.8970;
}

// one single-device handle on `dev` (its streams, its copy of the coefficient tables)
static RgpotB200Pot* create_on_device(const RgpotB200Config* cfg, int dev, std::string& msg) {
  cudaDeviceProp prop;
  if (cudaSetDevice(dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    msg = "rgpot_b200_create: cannot select the device";
    return nullptr;
  }
  if (prop.major < 10) {
    msg = "rgpot_b200_create: the engine is built for sm_100a (B200) only";
    return nullptr;
  }
  auto* pot = new RgpotB200Pot();
  pot->cfg = *cfg;
  pot->cfg.device = dev;
  pot->device = dev;
  pot->host_threads = cfg->host_threads > 0 ? cfg->host_threads : 0;
  pot->sm_count = prop.multiProcessorCount;
  pot->smem_optin = prop.sharedMemPerBlockOptin;
  pot->smem_per_sm = prop.sharedMemPerMultiprocessor;
  bool ok = cudaStreamCreateWithFlags(&pot->stream, cudaStreamNonBlocking) == cudaSuccess;
  for (auto& s : pot->pipe) ok = ok && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&pot->ev_ready, cudaEventDisableTiming) == cudaSuccess;
  for (auto& e : pot->ring_ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    delete pot;
    msg = "rgpot_b200_create: cannot create streams";
    return nullptr;
  }
  if (is_eam(cfg->kind)) {
    const EamParams p = cfg->kind == RGPOT_B200_EAM_AL ? make_eam_al_params() : make_eam_params();
    double tab[64];
    for (int j = 0; j < 64; ++j) tab[j] = exp2((double)j / 64.0);
    if (cudaMemcpyToSymbol(c_exp2tab, tab, sizeof tab) != cudaSuccess ||
        cudaMemcpyToSymbol(g_exp2tab, tab, sizeof tab) != cudaSuccess ||
        (cfg->kind == RGPOT_B200_EAM_AL ? cudaMemcpyToSymbol(c_eam_al, &p, sizeof p)
                                        : cudaMemcpyToSymbol(c_eam, &p, sizeof p)) != cudaSuccess ||
        rb::tile_upload_constants(cfg->kind == RGPOT_B200_EAM_AL ? 1 : 0, &p, sizeof p, tab) != cudaSuccess) {
      delete pot;
      msg = "rgpot_b200_create: cannot upload the EAM coefficients";
      return nullptr;
    }
  } else if (cfg->kind == RGPOT_B200_FEHE) {
    const FeHeParams p = make_fehe_params();
    if (cudaMemcpyToSymbol(c_fehe, &p, sizeof p) != cudaSuccess ||
        rb::tile_upload_constants(2, &p, sizeof p, nullptr) != cudaSuccess) {
      delete pot;
      msg = "rgpot_b200_create: cannot upload the Fe-He coefficients";
      return nullptr;
    }
  } else {
    lj_setup(pot);
  }
  return pot;
}

RgpotB200Pot* rgpot_b200_create(const RgpotB200Config* cfg, char* errbuf, size_t errlen) {
  auto err = [&](const std::string& m) -> RgpotB200Pot* {
    if (errbuf && errlen) snprintf(errbuf, errlen, "%s", m.c_str());
    return nullptr;
  };
  if (!cfg) return err("rgpot_b200_create: null config");
  if (cfg->kind < RGPOT_B200_LJ || cfg->kind > RGPOT_B200_FEHE)
    return err("rgpot_b200_create: unknown potential kind");
  if (cfg->kind == RGPOT_B200_ZBL && (!(cfg->zbl_cut_inner > 0.0) || !(cfg->zbl_cut_inner < cfg->cutoff)))
    return err("ZBLPot: invalid cutoffs, require 0.0 < cut_inner < cut_global.");  // ZBLPot.cc:76-79
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return err(std::string("rgpot_b200_create: no CUDA device (") +
               (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
               "); this engine has no CPU fallback");
  }
  int prev_dev = -1;
  if (cudaGetDevice(&prev_dev) != cudaSuccess) prev_dev = -1;
  struct Restore {  // the caller's current device is left as it was
    int d;
    ~Restore() {
      if (d >= 0) cudaSetDevice(d);
    }
  } restore{prev_dev};
  // the batch devices: an explicit list, every usable device (device = -2), or just one
  std::vector<int> devs;
  if (cfg->n_devices > 1) {
    if (cfg->n_devices > 16) return err("rgpot_b200_create: at most 16 devices");
    for (int k = 0; k < cfg->n_devices; ++k) {
      const int d = cfg->devices[k];
      if (d < 0 || d >= ndev) return err("rgpot_b200_create: device ordinal out of range");
      if (std::find(devs.begin(), devs.end(), d) != devs.end())
        return err("rgpot_b200_create: a device is listed twice");
      devs.push_back(d);
    }
  } else if (cfg->device == -2) {
    for (int d = 0; d < ndev && (int)devs.size() < 16; ++d) {
      cudaDeviceProp p;
      if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major >= 10) devs.push_back(d);
    }
    if (devs.empty()) return err("rgpot_b200_create: the engine is built for sm_100a (B200) only");
  } else {
    int dev = cfg->device;
    if (dev < 0) dev = prev_dev >= 0 ? prev_dev : 0;
    if (dev >= ndev) return err("rgpot_b200_create: device ordinal out of range");
    devs.push_back(dev);
  }
  std::string msg;
  RgpotB200Pot* pot = create_on_device(cfg, devs[0], msg);
  if (!pot) return err(msg);
  if (devs.size() > 1) {
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    const bool have_mask = sched_getaffinity(0, sizeof allowed, &allowed) == 0;
    int per_dev = cfg->host_threads;
    if (per_dev <= 0) per_dev = std::max(1, std::min(14, rb::usable_cpus() / (int)devs.size()));
    pot->host_threads = per_dev;
    for (int d : devs) {
      RgpotB200Config c1 = *cfg;
      c1.n_devices = 0;
      c1.device = d;
      c1.host_threads = per_dev;
      RgpotB200Pot* sp = create_on_device(&c1, d, msg);
      if (!sp) {
        delete pot;
        return err(msg);
      }
      RgpotB200Pot::Shard sh;
      sh.pot = sp;
      cpu_set_t local;
      if (have_mask && rb::gpu_local_cpus(d, allowed, local))
        sh.worker.reset(new rb::Worker(&local));
      else
        sh.worker.reset(new rb::Worker(nullptr));
      pot->shards.push_back(std::move(sh));
    }
  }
  return pot;
}

void rgpot_b200_destroy(RgpotB200Pot* pot) { delete pot; }

int rgpot_b200_last_error(RgpotB200Pot* pot, char* buffer, int buffer_len) {
  if (!pot || !buffer || buffer_len <= 0) return 0;
  const int n = (int)std::min<size_t>(pot->last_error.size(), (size_t)buffer_len - 1);
  memcpy(buffer, pot->last_error.data(), n);
  buffer[n] = 0;
  return n;
}

long rgpot_b200_launch_count(RgpotB200Pot* pot) { return pot ? pot->launches : 0; }

long rgpot_b200_stat(RgpotB200Pot* pot, const char* key) {
  if (!pot || !key) return -1;
  if (strcmp(key, "launches") == 0) return pot->launches;
  if (strcmp(key, "slow_atoms") == 0) return pot->slow_atoms;
  if (strcmp(key, "pool_used") == 0) return pot->pool_used;
  if (strcmp(key, "max_halo") == 0) return pot->memo.max_halo;
  if (strcmp(key, "brick") == 0) return pot->memo.bd[0] * 100 + pot->memo.bd[1] * 10 + pot->memo.bd[2];
  if (strcmp(key, "cap_cand") == 0) return pot->last_cap_cand;
  if (strcmp(key, "cap_list") == 0) return pot->last_cap_list;
  if (strcmp(key, "tile_used") == 0) return pot->last_used_tile;
  if (strncmp(key, "prof", 4) == 0 && key[4] >= '0' && key[4] <= '3') return pot->prof[key[4] - '0'];
  return -1;
}

int rgpot_b200_set_option(RgpotB200Pot* pot, const char* key, long value) {
  if (!pot || !key) return kStatusArgs;
  if (strcmp(key, "path") == 0) {
    pot->forced_path = (int)value;
    return 0;
  }
  if (strcmp(key, "brick") == 0) {  // decimal xyz, e.g. 221; 0 = automatic
    pot->opt_brick[0] = (int)(value / 100) % 10;
    pot->opt_brick[1] = (int)(value / 10) % 10;
    pot->opt_brick[2] = (int)value % 10;
    return 0;
  }
  if (strcmp(key, "brick_tpa") == 0) {
    pot->opt_tpa = (int)value;
    return 0;
  }
  if (strcmp(key, "brick_list") == 0) {
    pot->opt_list = (int)value;
    return 0;
  }
  if (strcmp(key, "brick_slack") == 0) {
    pot->opt_slack = (int)value;
    return 0;
  }
  if (strcmp(key, "fine") == 0) {
    pot->opt_fine = (int)value;
    return 0;
  }
  if (strcmp(key, "fine_min") == 0) {
    pot->opt_fine_min = value;
    return 0;
  }
  if (strcmp(key, "bricks") == 0) {
    pot->opt_bricks = (int)value;
    return 0;
  }
  if (strcmp(key, "handoff") == 0) {
    pot->opt_handoff = (int)value;
    return 0;
  }
  if (strcmp(key, "brick_threads") == 0) {
    pot->opt_threads = (int)value;
    return 0;
  }
  if (strcmp(key, "tile") == 0) {
    pot->opt_tile = (int)value;
    return 0;
  }
  if (strcmp(key, "tile_group") == 0) {
    pot->opt_tile_group = (int)value;
    return 0;
  }
  if (strcmp(key, "tile_brick") == 0) {
    pot->opt_tile_brick = (int)value;
    return 0;
  }
  if (strcmp(key, "tile_cap") == 0) {
    pot->opt_tile_cap = (int)value;
    return 0;
  }
  if (strcmp(key, "tile_wcap") == 0) {
    pot->opt_tile_wcap = (int)value;
    return 0;
  }
  if (strcmp(key, "one") == 0) {
    pot->opt_one = (int)value;
    return 0;
  }
  if (strcmp(key, "half") == 0) {
    pot->opt_half = (int)value;
    for (auto& sh : pot->shards) sh.pot->opt_half = (int)value;
    return 0;
  }
  if (strcmp(key, "batch_chunk") == 0) {
    pot->opt_batch_chunk = (int)value;
    for (auto& sh : pot->shards) sh.pot->opt_batch_chunk = (int)value;
    return 0;
  }
  return kStatusArgs;
}

void* rgpot_b200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void rgpot_b200_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

double rgpot_b200_measure_fp64_tflops(int device, int iters) {
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return 0.0;
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0.0;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return 0.0;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  if (iters < 1000) iters = 1000;
  iters = (iters + 15) / 16 * 16;  // (the kernel runs sixteen rounds per trip)
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(t0);
    k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(out);
  return best;
}

// What the host's memory system gives a pack / unpack: `threads` threads copy `bytes` from pageable
// memory into page-locked memory and back with the engine's own streaming copy, in 5 KB pieces like the
// systems of a batch; returns GB/s of payload (bytes moved one way + bytes moved back) / time, best of
// `reps`.  The ceiling of the pageable ForceBatch entry when the GPUs are not the limit.
double rgpot_b200_measure_host_copy_gbs(int threads, size_t bytes, int reps) {
  if (threads < 1 || bytes < (1u << 20) || reps < 1) return 0.0;
  void* pinned = nullptr;
  if (cudaMallocHost(&pinned, bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0.0;
  }
  std::vector<char> src(bytes, 1), dst(bytes, 0);
  rb::Team team(threads - 1, nullptr);
  const size_t piece = 5232, npieces = bytes / piece;
  double best = 0.0;
  for (int r = 0; r <= reps; ++r) {
    const auto t0 = std::chrono::steady_clock::now();
    team.run(npieces, 32, [&](size_t b, size_t e) {
      for (size_t k = b; k < e; ++k) rb::copy_stream(static_cast<char*>(pinned) + k * piece, src.data() + k * piece, piece);
      rb::copy_stream_fence();
    });
    team.run(npieces, 32, [&](size_t b, size_t e) {
      for (size_t k = b; k < e; ++k) rb::copy_stream(dst.data() + k * piece, static_cast<char*>(pinned) + k * piece, piece);
      rb::copy_stream_fence();
    });
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (r > 0) best = std::max(best, 2.0 * (double)(npieces * piece) / dt / 1e9);
  }
  cudaFreeHost(pinned);
  return best;
}

// What the host side of the box can move: every listed device copies h2d_bytes from and d2h_bytes into
// its own page-locked buffers (allocated by a thread on the CPUs next to that device), both directions
// at once, all devices at once; returns the aggregate GB/s of the slowest-finishing device set
// (bytes of all devices / wall time between a common start and the last finish).  This is the ceiling
// of any end-to-end number of a batch: bench.py prints it beside `e2e`.
double rgpot_b200_measure_copy_gbs(const int* devices, int n_devices, size_t h2d_bytes, size_t d2h_bytes,
                                   int reps) {
  if (n_devices < 1 || !devices || reps < 1) return 0.0;
  struct Lane {
    int dev;
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    bool ok = true;
  };
  std::vector<Lane> lanes(n_devices);
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  const bool have_mask = sched_getaffinity(0, sizeof allowed, &allowed) == 0;
  std::vector<std::unique_ptr<rb::Worker>> workers;
  for (int k = 0; k < n_devices; ++k) {
    lanes[k].dev = devices[k];
    cpu_set_t local;
    const bool place = have_mask && rb::gpu_local_cpus(devices[k], allowed, local);
    workers.emplace_back(new rb::Worker(place ? &local : nullptr));
  }
  auto all = [&](const std::function<void(Lane&)>& f) {
    for (int k = 0; k < n_devices; ++k) workers[k]->submit([&, k] { f(lanes[k]); });
    for (int k = 0; k < n_devices; ++k) workers[k]->wait();
  };
  all([&](Lane& l) {
    l.ok = cudaSetDevice(l.dev) == cudaSuccess && cudaMallocHost(&l.h_in, std::max<size_t>(h2d_bytes, 1)) == cudaSuccess &&
           cudaMallocHost(&l.h_out, std::max<size_t>(d2h_bytes, 1)) == cudaSuccess &&
           cudaMalloc(&l.d_in, std::max<size_t>(h2d_bytes, 1)) == cudaSuccess &&
           cudaMalloc(&l.d_out, std::max<size_t>(d2h_bytes, 1)) == cudaSuccess &&
           cudaStreamCreateWithFlags(&l.s_in, cudaStreamNonBlocking) == cudaSuccess &&
           cudaStreamCreateWithFlags(&l.s_out, cudaStreamNonBlocking) == cudaSuccess;
    if (l.ok) {
      memset(l.h_in, 1, std::max<size_t>(h2d_bytes, 1));
      memset(l.h_out, 1, std::max<size_t>(d2h_bytes, 1));
    }
  });
  bool ok = true;
  for (auto& l : lanes) ok = ok && l.ok;
  double best = 0.0;
  if (ok) {
    auto pass = [&](Lane& l) {
      cudaSetDevice(l.dev);
      // sixteen slices each way, like a pipelined batch
      const int slices = 16;
      for (int c = 0; c < slices; ++c) {
        const size_t i0 = h2d_bytes * c / slices, i1 = h2d_bytes * (c + 1) / slices;
        const size_t o0 = d2h_bytes * c / slices, o1 = d2h_bytes * (c + 1) / slices;
        if (i1 > i0) cudaMemcpyAsync((char*)l.d_in + i0, (char*)l.h_in + i0, i1 - i0, cudaMemcpyHostToDevice, l.s_in);
        if (o1 > o0) cudaMemcpyAsync((char*)l.h_out + o0, (char*)l.d_out + o0, o1 - o0, cudaMemcpyDeviceToHost, l.s_out);
      }
      cudaStreamSynchronize(l.s_in);
      cudaStreamSynchronize(l.s_out);
    };
    all(pass);  // warm-up
    for (int r = 0; r < reps; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      all(pass);
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const double gbs = (double)(h2d_bytes + d2h_bytes) * n_devices / dt / 1e9;
      best = std::max(best, gbs);
    }
  }
  all([&](Lane& l) {
    cudaSetDevice(l.dev);
    if (l.s_in) cudaStreamDestroy(l.s_in);
    if (l.s_out) cudaStreamDestroy(l.s_out);
    if (l.h_in) cudaFreeHost(l.h_in);
    if (l.h_out) cudaFreeHost(l.h_out);
    if (l.d_in) cudaFree(l.d_in);
    if (l.d_out) cudaFree(l.d_out);
  });
  return ok ? best : 0.0;
}

// ------------------------------------------------------------------------------------------------
// device-resident entry points
// ------------------------------------------------------------------------------------------------
static int force_device_impl(RgpotB200Pot* pot, long n, const double* d_pos, const int* d_Z,
                             double* d_F, double* d_E, const double* box, cudaStream_t st,
                             const EmitArgs* emit, int path) {
  const int kind = pot->cfg.kind;
  if (path == 0) path = pot->forced_path;
  bool small = small_eligible(pot, n);
  if (path == 1) small = false;
  if (path == 2 && !small_eligible(pot, n))
    return fail(pot, kStatusArgs, "rgpot_b200: system too large for the all-pairs kernels");
  if (small) {
    // a batch of one
    CU_TRY(pot, pot->b_off.reserve(2 * sizeof(long long)));
    CU_TRY(pot, pot->b_boxes.reserve(9 * sizeof(double)));
    CU_TRY(pot, pot->b_flags.reserve(sizeof(unsigned)));
    CU_TRY(pot, pot->b_bad.reserve(2 * sizeof(int)));
    const long long off[2] = {0, n};
    double hb[9] = {0};
    if (box) memcpy(hb, box, sizeof hb);
    CU_TRY(pot, cudaMemcpyAsync(pot->b_off.p, off, sizeof off, cudaMemcpyHostToDevice, st));
    CU_TRY(pot, cudaMemcpyAsync(pot->b_boxes.p, hb, sizeof hb, cudaMemcpyHostToDevice, st));
    CU_TRY(pot, cudaMemsetAsync(pot->b_flags.p, 0, sizeof(unsigned), st));
    const int maxnb = uses_eam_small(kind) ? eam_small_maxnb(pot, n) : 0;
    int s = launch_small(pot, 1, n, pot->b_off.as<long long>(), pot->b_boxes.as<double>(), d_pos, d_Z,
                         d_F, d_E, pot->b_flags.as<unsigned>(), pot->b_bad.as<int>(), maxnb, st, emit);
    if (s) return s;
    pot->pend.small_single = true;
    return 0;
  }
  pot->pend.small_single = false;
  return run_cell_path(pot, n, d_pos, d_Z, d_F, d_E, box, st, emit);
}

// Read back the flags of the last single-system call and turn them into a status.
// May rerun a list-overflow system on the cell path (needs the device inputs again).
static int finish_single(RgpotB200Pot* pot, long n, const double* d_pos, const int* d_Z, double* d_F,
                         double* d_E, const double* box, cudaStream_t st, const double* host_pos,
                         const int* host_Z) {
  unsigned flags = 0;
  int bad = INT_MAX, badz = 0;
  if (pot->pend.small_single) {
    int hb[2] = {INT_MAX, 0};
    CU_TRY(pot, cudaMemcpyAsync(&flags, pot->b_flags.p, sizeof flags, cudaMemcpyDeviceToHost, st));
    CU_TRY(pot, cudaMemcpyAsync(hb, pot->b_bad.p, sizeof hb, cudaMemcpyDeviceToHost, st));
    CU_TRY(pot, cudaStreamSynchronize(st));
    if (flags & kFlagListOverflow) {
      pot->pend.small_single = false;
      int s = run_cell_path(pot, n, d_pos, d_Z, d_F, d_E, box, st, nullptr, false);
      if (s) return s;
      return finish_single(pot, n, d_pos, d_Z, d_F, d_E, box, st, host_pos, host_Z);
    }
    bad = hb[0];
    badz = hb[1];
  } else {
    Status stt{};
    CU_TRY(pot, cudaMemcpyAsync(&stt, pot->status.p, sizeof stt, cudaMemcpyDeviceToHost, st));
    CU_TRY(pot, cudaStreamSynchronize(st));
    flags = stt.flags;
    pot->slow_atoms = stt.slow_atoms;
    pot->pool_used = (long)stt.pool_used;
    for (int k = 0; k < 4; ++k) pot->prof[k] = (long)stt.prof[k];
    if (stt.max_halo > 0 && pot->memo_pending.n >= 0) {  // brick kernels ran: remember what they met
      const auto& mp = pot->memo_pending;
      if (!pot->memo.matches(mp.n, mp.nc)) {
        pot->memo = mp;
        pot->memo.nshapes = 0;
      }
      pot->memo.record(mp.bd, stt.max_halo);
      for (int k = 0; k < 3; ++k) pot->memo.bd[k] = mp.bd[k];
      pot->memo.max_halo = pot->memo.lookup(mp.bd);
    }
    bad = stt.first_bad_atom;
    if ((flags & kFlagForeignZ) && bad != INT_MAX) {
      if (host_Z) {
        badz = host_Z[bad];
      } else {
        CU_TRY(pot, cudaMemcpy(&badz, d_Z + bad, sizeof(int), cudaMemcpyDeviceToHost));
      }
    }
  }
  return map_flags(pot, flags, bad, badz, host_pos);
}

int rgpot_b200_force_device(RgpotB200Pot* pot, long nAtoms, const double* d_positions,
                            const int* d_atomicNrs, double* d_forces, double* d_energy,
                            const double* box, void* cuda_stream) {
  if (!pot) return kStatusArgs;
  pot->last_error.clear();
  DevGuard guard(pot->device);
  CU_TRY(pot, guard.err);
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : pot->stream;
  pot->pend.active = false;
  if (nAtoms < 0 || nAtoms > kMaxAtoms) return fail(pot, kStatusArgs, "rgpot_b200: bad atom count");
  const bool eam = reads_species(pot->cfg.kind);  // (reads atomic numbers)
  // reference early-outs: rgpot_cuh2_capi.f90:50 / rgpot_eam_al_capi.f90:50 (natoms < 1), LJPot.cc:50-52,
  // LJClusterPot.cc:43-45
  const bool trivial = (is_eam(pot->cfg.kind) || pot->cfg.kind == RGPOT_B200_FEHE) ? nAtoms < 1
                           : (nAtoms < 2 || (pot->cfg.kind != RGPOT_B200_LJ_CLUSTER && pot->cfg.kind != RGPOT_B200_ZBL && !(pot->cfg.cutoff > 0.0)));
  if (trivial) {
    if (nAtoms > 0) CU_TRY(pot, cudaMemsetAsync(d_forces, 0, sizeof(double) * 3 * nAtoms, st));
    CU_TRY(pot, cudaMemsetAsync(d_energy, 0, sizeof(double), st));
    return 0;
  }
  const bool zbl = pot->cfg.kind == RGPOT_B200_ZBL;
  if ((eam || zbl) && !d_atomicNrs)
    return fail(pot, kStatusArgs, "rgpot_b200: this potential needs atomic numbers");
  if (pot->cfg.kind != RGPOT_B200_LJ_CLUSTER && !zbl && !box)
    return fail(pot, kStatusArgs, "rgpot_b200_force_device: null box");
  if (zbl) {
    // the species set decides the coefficient table (ZBLPot.cc:150-196): presence bitmap + device mapping
    int zs = zbl_prepare_device(pot, (size_t)nAtoms, d_atomicNrs, st);
    if (zs) return zs;
    d_atomicNrs = pot->Z.as<int>();
  }
  int s = force_device_impl(pot, nAtoms, d_positions, d_atomicNrs, d_forces, d_energy, box, st,
                            nullptr, 0);
  if (s) {
    cudaMemsetAsync(d_forces, 0, sizeof(double) * 3 * nAtoms, st);
    cudaMemsetAsync(d_energy, 0, sizeof(double), st);
    return s;
  }
  RgpotB200Pot::Pending& pd = pot->pend;
  pd.active = true;
  pd.batch = false;
  pd.n = nAtoms;
  pd.pos = d_positions;
  pd.Z = d_atomicNrs;
  pd.F = d_forces;
  pd.E = d_energy;
  pd.st = st;
  pd.has_box = box != nullptr;
  if (box) memcpy(pd.box, box, sizeof pd.box);
  return 0;
}

// forward declaration (batch device)
static int batch_sync_status(RgpotB200Pot* pot);

int rgpot_b200_sync_status(RgpotB200Pot* pot) {
  if (!pot) return kStatusArgs;
  DevGuard guard(pot->device);
  CU_TRY(pot, guard.err);
  RgpotB200Pot::Pending& pd = pot->pend;
  if (!pd.active) {
    CU_TRY(pot, cudaStreamSynchronize(pot->stream));
    return 0;
  }
  pd.active = false;
  if (pd.batch) return batch_sync_status(pot);
  int s = finish_single(pot, pd.n, pd.pos, pd.Z, pd.F, pd.E, pd.has_box ? pd.box : nullptr, pd.st,
                        nullptr, nullptr);
  if (s) {
    cudaMemsetAsync(pd.F, 0, sizeof(double) * 3 * pd.n, pd.st);
    cudaMemsetAsync(pd.E, 0, sizeof(double), pd.st);
    cudaStreamSynchronize(pd.st);
  }
  return s;
}

// ------------------------------------------------------------------------------------------------
// host-buffer single system
// ------------------------------------------------------------------------------------------------
// Large host arrays: page-locked memory is copied by the DMA engine straight from / into the caller's
// array; pageable memory (what Potential::operator() hands over, CppCore/rgpot/Potential.hpp:165-211)
// goes through a ring of pinned slices filled / drained by the helper threads, slice k + 1 on the CPU
// while slice k is on the bus.
constexpr size_t kRingSlice = 1u << 20;
constexpr int kRingSlots = 8;

// take the next ring slot, waiting for whatever copy still uses it
static int ring_acquire(RgpotB200Pot* pot, int& slot) {
  slot = (int)(pot->ring_next++ % kRingSlots);
  if (pot->ring_busy[slot]) {
    CU_TRY(pot, cudaEventSynchronize(pot->ring_ev[slot]));
    pot->ring_busy[slot] = false;
  }
  return 0;
}

static int upload_array(RgpotB200Pot* pot, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
  if (bytes < (1u << 20) || rb::host_ptr_is_pinned(h_src)) {
    CU_TRY(pot, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  CU_TRY(pot, pot->h_ring.reserve(kRingSlice * kRingSlots));
  rb::Team* team = get_team(pot);
  char* ring = pot->h_ring.as<char>();
  const size_t nslices = (bytes + kRingSlice - 1) / kRingSlice;
  for (size_t i = 0; i < nslices; ++i) {
    int slot = 0;
    int rs = ring_acquire(pot, slot);
    if (rs) return rs;
    const size_t o = i * kRingSlice, len = std::min(kRingSlice, bytes - o);
    char* dst = ring + (size_t)slot * kRingSlice;
    const char* src = static_cast<const char*>(h_src) + o;
    team->run(len, 32u << 10, [&](size_t b0, size_t b1) {
      rb::copy_stream(dst + b0, src + b0, b1 - b0);
      rb::copy_stream_fence();
    });
    CU_TRY(pot, cudaMemcpyAsync(static_cast<char*>(d_dst) + o, dst, len, cudaMemcpyHostToDevice, st));
    CU_TRY(pot, cudaEventRecord(pot->ring_ev[slot], st));
    pot->ring_busy[slot] = true;
  }
  return 0;
}

static int download_array(RgpotB200Pot* pot, void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
  if (bytes < (1u << 20) || rb::host_ptr_is_pinned(h_dst)) {
    CU_TRY(pot, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    return 0;
  }
  CU_TRY(pot, pot->h_ring.reserve(kRingSlice * kRingSlots));
  rb::Team* team = get_team(pot);
  char* ring = pot->h_ring.as<char>();
  const size_t nslices = (bytes + kRingSlice - 1) / kRingSlice;
  int slots[kRingSlots];
  auto issue = [&](size_t i) -> int {
    int slot = 0;
    int rs = ring_acquire(pot, slot);
    if (rs) return rs;
    slots[i % kRingSlots] = slot;
    const size_t o = i * kRingSlice, len = std::min(kRingSlice, bytes - o);
    CU_TRY(pot, cudaMemcpyAsync(ring + (size_t)slot * kRingSlice, static_cast<const char*>(d_src) + o, len,
                                cudaMemcpyDeviceToHost, st));
    CU_TRY(pot, cudaEventRecord(pot->ring_ev[slot], st));
    pot->ring_busy[slot] = true;
    return 0;
  };
  for (size_t i = 0; i < std::min<size_t>(nslices, kRingSlots); ++i) {
    int rs = issue(i);
    if (rs) return rs;
  }
  for (size_t i = 0; i < nslices; ++i) {
    const int slot = slots[i % kRingSlots];
    const size_t o = i * kRingSlice, len = std::min(kRingSlice, bytes - o);
    CU_TRY(pot, cudaEventSynchronize(pot->ring_ev[slot]));
    pot->ring_busy[slot] = false;
    const char* src = ring + (size_t)slot * kRingSlice;
    char* dst = static_cast<char*>(h_dst) + o;
    team->run(len, 32u << 10, [&](size_t b0, size_t b1) {
      rb::copy_stream(dst + b0, src + b0, b1 - b0);
      rb::copy_stream_fence();
    });
    if (i + kRingSlots < nslices) {
      int rs = issue(i + kRingSlots);
      if (rs) return rs;
    }
  }
  return 0;
}

// ONE EAM system of 64 .. 512 atoms on many blocks (single.cuh): one upload, two launches, results in the
// mapped block.  Returns 0, a status, or -1 when the system is not the clean case (the batch kernel then
// takes it and does the error reporting).
static int force_one_eam(RgpotB200Pot* pot, long nAtoms, const double* positions, const int* atomicNrs,
                         double* forces, double* energy, const double* box) {
  const size_t n = (size_t)nAtoms;
  const bool al = pot->cfg.kind == RGPOT_B200_EAM_AL;
  const int nb = (int)((nAtoms + kOneAtomsPerBlock - 1) / kOneAtomsPerBlock);
  // mapped block (8-byte words): box[9] | pos[3n] | Z[(n + 1) / 2] || F[3n] | E | flags
  const size_t in_words = 9 + 3 * n + (n + 1) / 2, words = in_words + 3 * n + 2;
  if (pot->h_fast.cap < words * 8) {
    if (pot->h_fast.p) cudaFreeHost(pot->h_fast.p);
    pot->h_fast.p = nullptr;
    pot->h_fast.cap = 0;
    const size_t want = std::max<size_t>(words * 8 * 2, 64u << 10);
    CU_TRY(pot, cudaHostAlloc(&pot->h_fast.p, want, cudaHostAllocMapped | cudaHostAllocPortable));
    pot->h_fast.cap = want;
    pot->h_fast_dev = nullptr;
  }
  if (!pot->h_fast_dev) CU_TRY(pot, cudaHostGetDevicePointer(&pot->h_fast_dev, pot->h_fast.p, 0));
  cudaStream_t st = pot->stream;
  // device scratch (8-byte words): ticket | fallback | inputs | dfd[n] | eemb[n] | staged forces[3n] | block energies
  const size_t dev_words = 2 + in_words + 5 * n + (size_t)nb;
  if (pot->d_fast.cap < dev_words * 8) {
    CU_TRY(pot, pot->d_fast.reserve(dev_words * 8 * 2));
    CU_TRY(pot, cudaMemsetAsync(pot->d_fast.p, 0, pot->d_fast.cap, st));
    pot->one_call_id = 0;
  }
  double* h = pot->h_fast.as<double>();
  double* hd = static_cast<double*>(pot->h_fast_dev);
  memcpy(h, box, 9 * sizeof(double));
  memcpy(h + 9, positions, sizeof(double) * 3 * n);
  if (!al) memcpy(h + 9 + 3 * n, atomicNrs, sizeof(int) * n);
  volatile unsigned* h_flags = reinterpret_cast<unsigned*>(h + in_words + 3 * n + 1);
  *h_flags = 0xffffffffu;
  double* dd = pot->d_fast.as<double>();
  CU_TRY(pot, cudaMemcpyAsync(dd + 2, h, sizeof(double) * in_words, cudaMemcpyHostToDevice, st));
  OneArgs a{};
  a.box = dd + 2;
  a.pos = dd + 2 + 9;
  a.Z = reinterpret_cast<const int*>(dd + 2 + 9 + 3 * n);
  a.n = (int)nAtoms;
  a.dfd = dd + 2 + in_words;
  a.eemb = a.dfd + n;
  a.f_stage = a.eemb + n;
  a.e_parts = a.f_stage + 3 * n;
  a.ticket = reinterpret_cast<int*>(dd);
  a.fallback = reinterpret_cast<unsigned*>(dd + 1);
  a.call_id = ++pot->one_call_id;
  a.F = hd + in_words;
  a.E = hd + in_words + 3 * n;
  a.flags = reinterpret_cast<unsigned*>(hd + in_words + 3 * n + 1);
  const size_t smem = one_smem_bytes((int)nAtoms);
  const int threads = kOneAtomsPerBlock * kOneLanes;
  if (!one_prepare_box(box, eam_rcut(pot->cfg.kind), a.sb, a.sbf)) return -1;  // singular / narrow cell: the batch kernel's case
  {
    // a cooperative launch: all blocks resident, so the grid barrier between the passes cannot hang
    void* kargs[] = {(void*)&a};
    const void* fn = al ? (const void*)k_eam_one<EAM_AL> : (const void*)k_eam_one<EAM_CUH2>;
    CU_TRY(pot, cudaLaunchCooperativeKernel(fn, dim3(nb), dim3(threads), kargs, smem, st));
  }
  LAUNCH_CHECK(pot, "k_eam_one");
  {
    const auto t_start = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (*h_flags == 0xffffffffu) {
      if ((++spins & 1023u) == 0 &&
          std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 2.0e-3) {
        CU_TRY(pot, cudaStreamSynchronize(st));
        break;
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  const unsigned flags = *h_flags;
  if (flags == 0xffffffffu) return fail(pot, kStatusCuda, "rgpot_b200: the single-system kernels left no status");
  if (flags != 0u) return -1;
  memcpy(forces, h + in_words, sizeof(double) * 3 * n);
  *energy = h[in_words + 3 * n];
  return 0;
}

// Single small systems (BASELINE configs[0] / [1], the reference's PotBench sizes): the whole call is
// ONE kernel launch.  Inputs are written into a page-locked block that the kernel reads in place over
// PCIe (a few kilobytes, read once), forces / energy / status come back the same way: no copy
// commands, no flag reset, one synchronisation.  Returns -1 when the system has to take the general
// path instead (list overflow: far images, crowded atoms).
static int force_small_fast(RgpotB200Pot* pot, long nAtoms, const double* positions, const int* atomicNrs,
                            double* forces, double* energy, const double* box) {
  const size_t n = (size_t)nAtoms;
  const bool eam = reads_species(pot->cfg.kind);
  // host block (8-byte words): off[2] | box[9] | - | pos[3n] | Z[(n + 1) / 2] || F[3n] | E | flags, - | bad[2]
  // (the pad word puts the positions on a 16-byte boundary: vector loads in the kernels)
  const size_t z_words = eam ? (n + 1) / 2 : 0, in_words = 12 + 3 * n + z_words;
  const size_t words = in_words + 3 * n + 3;
  if (pot->h_fast.cap < words * 8) {
    if (pot->h_fast.p) cudaFreeHost(pot->h_fast.p);
    pot->h_fast.p = nullptr;
    pot->h_fast.cap = 0;
    const size_t want = std::max<size_t>(words * 8 * 2, 64u << 10);
    CU_TRY(pot, cudaHostAlloc(&pot->h_fast.p, want, cudaHostAllocMapped | cudaHostAllocPortable));
    pot->h_fast.cap = want;
    pot->h_fast_dev = nullptr;
  }
  if (!pot->h_fast_dev) CU_TRY(pot, cudaHostGetDevicePointer(&pot->h_fast_dev, pot->h_fast.p, 0));
  double* h = pot->h_fast.as<double>();
  double* d = static_cast<double*>(pot->h_fast_dev);
  long long* h_off = reinterpret_cast<long long*>(h);
  double* h_box = h + 2;
  double* h_pos = h + 12;
  int* h_Z = reinterpret_cast<int*>(h + 12 + 3 * n);
  double* h_F = h + in_words;
  double* h_E = h_F + 3 * n;
  unsigned* h_flags = reinterpret_cast<unsigned*>(h_E + 1);
  int* h_bad = reinterpret_cast<int*>(h_E + 2);
  h_off[0] = 0;
  h_off[1] = (long long)n;
  if (box)
    memcpy(h_box, box, 9 * sizeof(double));
  else
    memset(h_box, 0, 9 * sizeof(double));
  memcpy(h_pos, positions, sizeof(double) * 3 * n);
  if (eam) memcpy(h_Z, atomicNrs, sizeof(int) * n);
  *h_flags = 0xffffffffu;
  h_bad[0] = INT_MAX;
  h_bad[1] = 0;
  const size_t iE = in_words + 3 * n;
  const int maxnb = uses_eam_small(pot->cfg.kind) ? eam_small_maxnb(pot, nAtoms) : 0;
  cudaStream_t st = pot->stream;
  // The inputs go to the device in ONE copy command (a kernel that reads them in place pays a PCIe round
  // trip of ~1.5 us per dependent fetch, and one per 128 bytes and warp beyond the first); the results
  // land in the mapped block behind a system-scope fence.
  const bool pairpot = !uses_eam_small(pot->cfg.kind);
  int lanes = 0, split = 0;
  if (pairpot) {
    // lanes per atom: a warp for anything but the smallest clusters (every lane then has a partner or two)
    lanes = 32;
    while (lanes > 1 && lanes / 2 >= nAtoms) lanes >>= 1;
    if (nAtoms > 1024) lanes = 16;
    const int apb = 256 / lanes;
    split = (long)nAtoms * lanes <= 256 ? 1 : (int)((nAtoms + apb - 1) / apb);
  }
  // device block (8-byte words): ticket | - | inputs | block partials | staged forces
  const size_t need = sizeof(double) * (2 + in_words + (size_t)split + 3 * n) + 64;
  if (pot->d_fast.cap < need) {
    CU_TRY(pot, pot->d_fast.reserve(need * 2));
    CU_TRY(pot, cudaMemsetAsync(pot->d_fast.p, 0, pot->d_fast.cap, st));  // (the ticket starts at zero)
    pot->one_call_id = 0;  // (force_one_eam shares the block)
  }
  double* dd = pot->d_fast.as<double>() + 2;
  CU_TRY(pot, cudaMemcpyAsync(dd, h, sizeof(double) * in_words, cudaMemcpyHostToDevice, st));
  int s = launch_small(pot, 1, nAtoms, nullptr, dd + 2, dd + 12, eam ? reinterpret_cast<const int*>(dd + 12 + 3 * n) : nullptr,
                       d + in_words, d + iE, reinterpret_cast<unsigned*>(d + iE + 1), reinterpret_cast<int*>(d + iE + 2),
                       maxnb, st, nullptr, false, (int)nAtoms, split, dd + in_words, reinterpret_cast<int*>(dd - 2), lanes,
                       box ? h_box : nullptr, dd + in_words + split);
  if (s) return s;
  {
    // the kernel writes the status word last, behind a system-scope fence: poll it in place (a stream
    // synchronisation costs several microseconds more), with the driver as the fallback
    volatile unsigned* vf = h_flags;
    const auto t_start = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (*vf == 0xffffffffu) {
      if ((++spins & 1023u) == 0 &&
          std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 2.0e-3) {
        CU_TRY(pot, cudaStreamSynchronize(st));
        break;
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  const unsigned flags = *h_flags;
  if (flags == 0xffffffffu) return fail(pot, kStatusCuda, "rgpot_b200: the single-system kernel left no status");
  if (flags & kFlagListOverflow) return -1;
  s = map_flags(pot, flags, h_bad[0], h_bad[1], positions);
  if (s) return s;
  memcpy(forces, h_F, sizeof(double) * 3 * n);
  *energy = *h_E;
  return 0;
}

int rgpot_b200_force(RgpotB200Pot* pot, long nAtoms, const double* positions, const int* atomicNrs,
                     double* forces, double* energy, double* variance, const double* box) {
  if (!pot) return kStatusArgs;
  pot->last_error.clear();
  if (variance) *variance = 0.0;
  if (energy) *energy = 0.0;
  // outputs are zeroed on every early return and on failure (the reference zeroes first,
  // LJPot.cc:44-49 / rgpot_cuh2_capi.f90:47-48); on success they are overwritten wholesale
  struct ZeroOnExit {
    double* f;
    size_t n;
    bool armed = true;
    ~ZeroOnExit() {
      if (armed && f && n) memset(f, 0, sizeof(double) * 3 * n);
    }
  } zero{forces, nAtoms > 0 ? (size_t)nAtoms : 0};
  if (!energy || (nAtoms > 0 && (!positions || !forces)))
    return fail(pot, kStatusArgs, "rgpot_b200_force: null buffer");
  DevGuard guard(pot->device);
  CU_TRY(pot, guard.err);
  if (nAtoms < 0 || nAtoms > kMaxAtoms) return fail(pot, kStatusArgs, "rgpot_b200: bad atom count");
  const bool eam = reads_species(pot->cfg.kind);  // (reads atomic numbers)
  const bool trivial = (is_eam(pot->cfg.kind) || pot->cfg.kind == RGPOT_B200_FEHE) ? nAtoms < 1
                           : (nAtoms < 2 || (pot->cfg.kind != RGPOT_B200_LJ_CLUSTER && pot->cfg.kind != RGPOT_B200_ZBL && !(pot->cfg.cutoff > 0.0)));
  if (trivial) return 0;
  const bool zbl = pot->cfg.kind == RGPOT_B200_ZBL;
  if ((eam || zbl) && !atomicNrs) return fail(pot, kStatusArgs, "rgpot_b200: this potential needs atomic numbers");
  if (pot->cfg.kind != RGPOT_B200_LJ_CLUSTER && pot->cfg.kind != RGPOT_B200_ZBL && !box)
    return fail(pot, kStatusArgs, "rgpot_b200_force: null box");
  if (is_eam(pot->cfg.kind) && pot->forced_path == 0 && pot->opt_one && nAtoms >= 64 && small_eligible(pot, nAtoms)) {
    // one EAM system of a few hundred atoms: many blocks, two launches (single.cuh)
    const int fo = force_one_eam(pot, nAtoms, positions, atomicNrs, forces, energy, box);
    if (fo == 0) {
      zero.armed = false;
      return 0;
    }
    if (fo > 0) {
      *energy = 0.0;
      return fo;
    }
    // fo < 0: not the clean case; the batch kernel below takes the system
  }
  if (!zbl && pot->forced_path != 1 && small_eligible(pot, nAtoms)) {
    const int fs = force_small_fast(pot, nAtoms, positions, atomicNrs, forces, energy, box);
    if (fs == 0) {
      zero.armed = false;
      return 0;
    }
    if (fs > 0) {
      *energy = 0.0;
      return fs;
    }
    // fs < 0: the all-pairs kernel bowed out; the cell-list kernels below take the system
  }
  cudaStream_t st = pot->stream;
  const size_t n = (size_t)nAtoms;
  CU_TRY(pot, pot->pos.reserve(sizeof(double) * 3 * n));
  CU_TRY(pot, pot->F.reserve(sizeof(double) * 3 * n));
  CU_TRY(pot, pot->energy.reserve(sizeof(double)));
  {
    int us = upload_array(pot, pot->pos.p, positions, sizeof(double) * 3 * n, st);
    if (us) return us;
  }
  const int* d_Z = nullptr;
  if (eam) {
    CU_TRY(pot, pot->Z.reserve(sizeof(int) * n));
    int us = upload_array(pot, pot->Z.p, atomicNrs, sizeof(int) * n, st);
    if (us) return us;
    d_Z = pot->Z.as<int>();
  } else if (zbl) {
    // species-pair tables are keyed on the set of atomic numbers (ZBLPot.cc:150-196)
    std::vector<int> types(n);
    int zs = zbl_prepare(pot, n, atomicNrs, types.data(), st);
    if (zs) return zs;
    CU_TRY(pot, pot->Z.reserve(sizeof(int) * n));
    CU_TRY(pot, cudaMemcpy(pot->Z.p, types.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    d_Z = pot->Z.as<int>();
  }
  const int saved_path = pot->forced_path;
  if (!zbl && small_eligible(pot, nAtoms) && pot->forced_path != 2) pot->forced_path = 1;  // (the fast path bowed out)
  int s = force_device_impl(pot, nAtoms, pot->pos.as<double>(), d_Z, pot->F.as<double>(),
                            pot->energy.as<double>(), box, st, nullptr, 0);
  pot->forced_path = saved_path;
  if (s) return s;
  s = finish_single(pot, nAtoms, pot->pos.as<double>(), d_Z, pot->F.as<double>(),
                    pot->energy.as<double>(), box, st, positions, atomicNrs);
  if (s) return s;  // outputs stay zeroed, like the reference
  s = download_array(pot, forces, pot->F.p, sizeof(double) * 3 * n, st);
  if (s) return s;
  CU_TRY(pot, cudaMemcpyAsync(energy, pot->energy.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaStreamSynchronize(st));
  zero.armed = false;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// neighbor table inspection
// ------------------------------------------------------------------------------------------------
int rgpot_b200_neighbors(RgpotB200Pot* pot, long nAtoms, const double* positions, const int* atomicNrs,
                         const double* box, int path, long cap, int32_t* pairs, int32_t* shifts,
                         long* n_out) {
  if (!pot || !n_out) return kStatusArgs;
  pot->last_error.clear();
  *n_out = 0;
  DevGuard guard(pot->device);
  CU_TRY(pot, guard.err);
  if (nAtoms < 1) return 0;
  cudaStream_t st = pot->stream;
  const size_t n = (size_t)nAtoms;
  const bool eam = reads_species(pot->cfg.kind);
  CU_TRY(pot, pot->pos.reserve(sizeof(double) * 3 * n));
  CU_TRY(pot, pot->F.reserve(sizeof(double) * 3 * n));
  CU_TRY(pot, pot->energy.reserve(sizeof(double)));
  CU_TRY(pot, cudaMemcpyAsync(pot->pos.p, positions, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  const int* d_Z = nullptr;
  std::vector<int> fakeZ;
  if (eam) {
    CU_TRY(pot, pot->Z.reserve(sizeof(int) * n));
    if (!atomicNrs) {
      fakeZ.assign(n, 29);
      atomicNrs = fakeZ.data();
    }
    CU_TRY(pot, cudaMemcpyAsync(pot->Z.p, atomicNrs, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    d_Z = pot->Z.as<int>();
  } else if (pot->cfg.kind == RGPOT_B200_ZBL) {
    std::vector<int> z(n, 1), types(n);
    int zs = zbl_prepare(pot, n, atomicNrs ? atomicNrs : z.data(), types.data(), st);
    if (zs) return zs;
    CU_TRY(pot, pot->Z.reserve(sizeof(int) * n));
    CU_TRY(pot, cudaMemcpy(pot->Z.p, types.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    d_Z = pot->Z.as<int>();
  }
  const size_t capz = (size_t)std::max<long>(cap, 1);
  CU_TRY(pot, pot->emit_pairs.reserve(sizeof(int32_t) * 2 * capz));
  CU_TRY(pot, pot->emit_shifts.reserve(sizeof(int32_t) * 3 * capz));
  CU_TRY(pot, pot->emit_count.reserve(sizeof(unsigned long long)));
  CU_TRY(pot, cudaMemsetAsync(pot->emit_count.p, 0, sizeof(unsigned long long), st));
  EmitArgs em{pot->emit_pairs.as<int32_t>(), pot->emit_shifts.as<int32_t>(),
              pot->emit_count.as<unsigned long long>(), cap};
  int s = force_device_impl(pot, nAtoms, pot->pos.as<double>(), d_Z, pot->F.as<double>(),
                            pot->energy.as<double>(), box, st, &em, path);
  if (s) return s;
  if (pot->pend.small_single) {
    // the all-pairs kernel bows out of systems whose lists it cannot hold (far images, crowded
    // atoms): the table then comes from the cell-list kernels, like the forces would
    unsigned flags = 0;
    CU_TRY(pot, cudaMemcpyAsync(&flags, pot->b_flags.p, sizeof flags, cudaMemcpyDeviceToHost, st));
    CU_TRY(pot, cudaStreamSynchronize(st));
    if (flags & kFlagListOverflow) {
      CU_TRY(pot, cudaMemsetAsync(pot->emit_count.p, 0, sizeof(unsigned long long), st));
      s = run_cell_path(pot, nAtoms, pot->pos.as<double>(), d_Z, pot->F.as<double>(),
                        pot->energy.as<double>(), box, st, &em);
      if (s) return s;
    }
  }
  unsigned long long cnt = 0;
  CU_TRY(pot, cudaMemcpyAsync(&cnt, pot->emit_count.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaStreamSynchronize(st));
  *n_out = (long)cnt;
  const size_t m = (size_t)std::min<unsigned long long>(cnt, (unsigned long long)std::max<long>(cap, 0));
  if (m > 0 && pairs && shifts) {
    CU_TRY(pot, cudaMemcpy(pairs, pot->emit_pairs.p, sizeof(int32_t) * 2 * m, cudaMemcpyDeviceToHost));
    CU_TRY(pot, cudaMemcpy(shifts, pot->emit_shifts.p, sizeof(int32_t) * 3 * m, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// vesin-shaped neighbor lists
// ------------------------------------------------------------------------------------------------
void rgpot_b200_neighbor_list_free(RgpotB200NeighborList* nl) {
  if (!nl) return;
  free(nl->pairs);
  free(nl->shifts);
  free(nl->distances);
  free(nl->vectors);
  memset(nl, 0, sizeof *nl);
}

int rgpot_b200_neighbor_list(RgpotB200Pot* pot, size_t n_points, const double* points, const double* box,
                             const int* periodic, double cutoff, int full, int sorted,
                             RgpotB200NeighborList* out) {
  if (!pot || !out) return kStatusArgs;
  pot->last_error.clear();
  rgpot_b200_neighbor_list_free(out);
  if (!(cutoff > 0.0) || !std::isfinite(cutoff)) return fail(pot, kStatusArgs, "rgpot_b200_neighbor_list: cutoff must be positive and finite");
  if (n_points > (size_t)kMaxAtoms) return fail(pot, kStatusArgs, "rgpot_b200: bad atom count");
  if (n_points == 0) return 0;
  if (!points) return fail(pot, kStatusArgs, "rgpot_b200_neighbor_list: null points");
  const bool per = periodic ? (periodic[0] && periodic[1] && periodic[2]) : true;
  const bool none = periodic && !periodic[0] && !periodic[1] && !periodic[2];
  if (!per && !none)
    return fail(pot, kStatusArgs, "rgpot_b200_neighbor_list: mixed periodicity is not supported (all or none)");
  if (per && !box) return fail(pot, kStatusArgs, "rgpot_b200_neighbor_list: null box");
  DevGuard guard(pot->device);
  CU_TRY(pot, guard.err);
  cudaStream_t st = pot->stream;
  const size_t n = n_points;
  // A free system is searched inside a fictitious cubic cell so wide that no image comes within the
  // cutoff: the pairs are the free ones.  The points are NOT moved: atoms outside the cell are
  // handled by the per-pair shift bookkeeping, which for two atoms of the same image gives S = 0 and
  // therefore exactly (p_j - p_i), vesin's expression for a free system.
  double cell[9];
  if (per) {
    memcpy(cell, box, sizeof cell);
  } else {
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, maxabs = 0.0;
    for (size_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        const double v = points[3 * i + k];
        if (!std::isfinite(v)) return fail(pot, 1, "rgpot_b200_neighbor_list: a point has a non-finite coordinate");
        lo[k] = std::min(lo[k], v);
        hi[k] = std::max(hi[k], v);
        maxabs = std::max(maxabs, fabs(v));
      }
    double ext = 0.0;
    for (int k = 0; k < 3; ++k) ext = std::max(ext, hi[k] - lo[k]);
    const double L = std::max(ext + 2.5 * cutoff + 1.0, maxabs / 400.0);  // (wrap counts stay below 500)
    memset(cell, 0, sizeof cell);
    cell[0] = cell[4] = cell[8] = L;
  }
  const double* search_points = points;
  CU_TRY(pot, pot->pos.reserve(sizeof(double) * 3 * n));
  CU_TRY(pot, pot->F.reserve(sizeof(double) * 3 * n));
  CU_TRY(pot, pot->energy.reserve(sizeof(double)));
  CU_TRY(pot, pot->emit_count.reserve(sizeof(unsigned long long)));
  CU_TRY(pot, cudaMemcpyAsync(pot->pos.p, search_points, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  // two passes: count, then fill (the list length is not known beforehand)
  unsigned long long cnt = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const size_t cap = pass == 0 ? 0 : (size_t)cnt;
    if (pass == 1) {
      if (cnt == 0) break;
      CU_TRY(pot, pot->emit_pairs.reserve(sizeof(int32_t) * 2 * cap));
      CU_TRY(pot, pot->emit_shifts.reserve(sizeof(int32_t) * 3 * cap));
      CU_TRY(pot, pot->emit_vecs.reserve(sizeof(double) * 3 * cap));
      CU_TRY(pot, pot->emit_dists.reserve(sizeof(double) * cap));
    }
    CU_TRY(pot, cudaMemsetAsync(pot->emit_count.p, 0, sizeof(unsigned long long), st));
    EmitArgs em{};
    em.pairs = pot->emit_pairs.as<int32_t>();
    em.shifts = pot->emit_shifts.as<int32_t>();
    em.count = pot->emit_count.as<unsigned long long>();
    em.cap = (long)cap;
    em.vecs = pot->emit_vecs.as<double>();
    em.dists = pot->emit_dists.as<double>();
    em.half = full ? 0 : 1;
    int s = run_cell_path(pot, (long)n, pot->pos.as<double>(), nullptr, pot->F.as<double>(),
                          pot->energy.as<double>(), cell, st, &em, false, cutoff);
    if (s) return s;
    Status stt{};
    CU_TRY(pot, cudaMemcpyAsync(&stt, pot->status.p, sizeof stt, cudaMemcpyDeviceToHost, st));
    CU_TRY(pot, cudaMemcpyAsync(&cnt, pot->emit_count.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
    CU_TRY(pot, cudaStreamSynchronize(st));
    if (stt.flags & (kFlagNonFinite | kFlagShiftRange | kFlagSingularBox))
      return map_flags(pot, stt.flags & (kFlagNonFinite | kFlagShiftRange | kFlagSingularBox), stt.first_bad_atom, 0, points);
  }
  const size_t m = (size_t)cnt;
  out->length = m;
  if (m == 0) return 0;
  std::vector<int32_t> hp(2 * m), hs(3 * m);
  out->pairs = (size_t*)malloc(sizeof(size_t) * 2 * m);
  out->shifts = (int32_t*)malloc(sizeof(int32_t) * 3 * m);
  out->distances = (double*)malloc(sizeof(double) * m);
  out->vectors = (double*)malloc(sizeof(double) * 3 * m);
  if (!out->pairs || !out->shifts || !out->distances || !out->vectors) {
    rgpot_b200_neighbor_list_free(out);
    return fail(pot, kStatusArgs, "rgpot_b200_neighbor_list: out of host memory");
  }
  CU_TRY(pot, cudaMemcpy(hp.data(), pot->emit_pairs.p, sizeof(int32_t) * 2 * m, cudaMemcpyDeviceToHost));
  CU_TRY(pot, cudaMemcpy(hs.data(), pot->emit_shifts.p, sizeof(int32_t) * 3 * m, cudaMemcpyDeviceToHost));
  std::vector<size_t> order(m);
  for (size_t k = 0; k < m; ++k) order[k] = k;
  if (sorted)  // by the first index, like vesin (GrowableNeighborList::sort sorts on (i, j, shift))
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      if (hp[2 * a] != hp[2 * b]) return hp[2 * a] < hp[2 * b];
      if (hp[2 * a + 1] != hp[2 * b + 1]) return hp[2 * a + 1] < hp[2 * b + 1];
      for (int k = 0; k < 3; ++k)
        if (hs[3 * a + k] != hs[3 * b + k]) return hs[3 * a + k] < hs[3 * b + k];
      return false;
    });
  std::vector<double> hv(3 * m), hd(m);
  CU_TRY(pot, cudaMemcpy(hv.data(), pot->emit_vecs.p, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost));
  CU_TRY(pot, cudaMemcpy(hd.data(), pot->emit_dists.p, sizeof(double) * m, cudaMemcpyDeviceToHost));
  for (size_t k = 0; k < m; ++k) {
    const size_t src = order[k];
    out->pairs[2 * k] = (size_t)hp[2 * src];
    out->pairs[2 * k + 1] = (size_t)hp[2 * src + 1];
    for (int c = 0; c < 3; ++c) {
      out->shifts[3 * k + c] = hs[3 * src + c];  // (all zero for a free system)
      out->vectors[3 * k + c] = hv[3 * src + c];
    }
    out->distances[k] = hd[src];
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// batches
// ------------------------------------------------------------------------------------------------
// Device-resident packed batch.  offsets: host array (nSystems + 1); d_boxes: device, 9 per system.
// host copy of one system's box: from the caller's host array when there is one, else from the device
static int fetch_box(RgpotB200Pot* pot, const double* h_boxes, const double* d_boxes, size_t id,
                     double* out) {
  if (h_boxes) {
    memcpy(out, h_boxes + 9 * id, 9 * sizeof(double));
    return 0;
  }
  CU_TRY(pot, cudaMemcpy(out, d_boxes + 9 * id, 9 * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

static int batch_device_impl(RgpotB200Pot* pot, size_t nsys, const size_t* offsets, const double* d_pos,
                             const int* d_Z, const double* d_boxes, const double* h_boxes,
                             double* d_F, double* d_E, cudaStream_t st) {
  const int kind = pot->cfg.kind;
  const bool eam = is_eam(kind);
  // plan: which systems fit the all-pairs kernel
  long nmax_small = 0;
  std::vector<int> small_ids, big_ids;
  small_ids.reserve(nsys);
  for (size_t s = 0; s < nsys; ++s) {
    const long n = (long)(offsets[s + 1] - offsets[s]);
    if (pot->forced_path != 1 && small_eligible(pot, n)) {
      small_ids.push_back((int)s);
      nmax_small = std::max(nmax_small, n);
    } else {
      big_ids.push_back((int)s);
    }
  }
  CU_TRY(pot, pot->b_off.reserve(sizeof(long long) * (nsys + 1)));
  CU_TRY(pot, pot->h_off.reserve(sizeof(long long) * (nsys + 1)));
  CU_TRY(pot, pot->b_flags.reserve(sizeof(unsigned) * nsys));
  CU_TRY(pot, pot->b_bad.reserve(sizeof(int) * 2 * nsys));
  long long* ho = pot->h_off.as<long long>();
  for (size_t s = 0; s <= nsys; ++s) ho[s] = (long long)offsets[s];
  CU_TRY(pot, cudaMemcpyAsync(pot->b_off.p, ho, sizeof(long long) * (nsys + 1), cudaMemcpyHostToDevice, st));
  CU_TRY(pot, cudaMemsetAsync(pot->b_flags.p, 0, sizeof(unsigned) * nsys, st));

  if (!small_ids.empty()) {
    const int maxnb = uses_eam_small(kind) ? eam_small_maxnb(pot, nmax_small) : 0;
    if (small_ids.size() == nsys) {
      int s = launch_small(pot, nsys, nmax_small, pot->b_off.as<long long>(), d_boxes, d_pos, d_Z, d_F,
                           d_E, pot->b_flags.as<unsigned>(), pot->b_bad.as<int>(), maxnb, st, nullptr);
      if (s) return s;
    } else {
      // per-system launches for the small members of a mixed batch (rare)
      for (int id : small_ids) {
        const long n = (long)(offsets[id + 1] - offsets[id]);
        int s = launch_small(pot, 1, n, pot->b_off.as<long long>() + id, d_boxes + 9 * (size_t)id, d_pos,
                             d_Z, d_F, d_E + id, pot->b_flags.as<unsigned>() + id,
                             pot->b_bad.as<int>() + 2 * (size_t)id, maxnb, st, nullptr);
        if (s) return s;
      }
    }
  }
  // big systems: one cell-list pass each (status folded into the per-system flag array)
  for (int id : big_ids) {
    const long n = (long)(offsets[id + 1] - offsets[id]);
    const size_t o = offsets[id];
    const bool trivial = uses_eam_small(kind) ? n < 1 : (n < 2 || (kind != RGPOT_B200_LJ_CLUSTER && kind != RGPOT_B200_ZBL && !(pot->cfg.cutoff > 0.0)));
    if (trivial) {
      if (n > 0) CU_TRY(pot, cudaMemsetAsync(d_F + 3 * o, 0, sizeof(double) * 3 * n, st));
      CU_TRY(pot, cudaMemsetAsync(d_E + id, 0, sizeof(double), st));
      continue;
    }
    double hbox[9];
    int s = fetch_box(pot, h_boxes, d_boxes, (size_t)id, hbox);
    if (s) return s;
    s = run_cell_path(pot, n, d_pos + 3 * o, d_Z ? d_Z + o : nullptr, d_F + 3 * o, d_E + id, hbox, st,
                      nullptr);
    if (s == 1) {
      // singular box of one member: record, keep going
      unsigned f = kFlagSingularBox;
      CU_TRY(pot, cudaMemcpyAsync(pot->b_flags.as<unsigned>() + id, &f, sizeof f, cudaMemcpyHostToDevice, st));
      continue;
    }
    if (s) return s;
    // fold the single-system Status into the per-system arrays
    CU_TRY(pot, cudaMemcpyAsync(pot->b_flags.as<unsigned>() + id, pot->status.p, sizeof(unsigned),
                                cudaMemcpyDeviceToDevice, st));
    CU_TRY(pot, cudaMemcpyAsync(pot->b_bad.as<int>() + 2 * (size_t)id,
                                (char*)pot->status.p + offsetof(Status, first_bad_atom), sizeof(int),
                                cudaMemcpyDeviceToDevice, st));
    // the status buffer is reused by the next system: order is kept by the stream
  }
  return 0;
}

// After a batch ran: read flags, rerun list-overflow systems on the cell path, produce statuses.
// Returns the first non-zero status.  h_* describe where the data of the batch lives.
static int batch_finish(RgpotB200Pot* pot, size_t nsys, const size_t* offsets, const double* d_pos,
                        const int* d_Z, const double* h_boxes, const double* d_boxes, double* d_F,
                        double* d_E, cudaStream_t st, const double* host_pos,
                        const std::function<const int*(size_t)>& host_Z_of, int* statuses,
                        std::vector<size_t>* touched = nullptr) {
  CU_TRY(pot, pot->h_flags.reserve(sizeof(unsigned) * nsys));
  CU_TRY(pot, pot->h_bad.reserve(sizeof(int) * 2 * nsys));
  unsigned* hf = pot->h_flags.as<unsigned>();
  int* hb = pot->h_bad.as<int>();
  CU_TRY(pot, cudaMemcpyAsync(hf, pot->b_flags.p, sizeof(unsigned) * nsys, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaMemcpyAsync(hb, pot->b_bad.p, sizeof(int) * 2 * nsys, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaStreamSynchronize(st));
  int first = 0;
  std::string first_msg;
  for (size_t s = 0; s < nsys; ++s) {
    unsigned f = hf[s];
    const long n = (long)(offsets[s + 1] - offsets[s]);
    const size_t o = offsets[s];
    int bad = hb[2 * s], badz = hb[2 * s + 1];
    if (f & kFlagListOverflow) {
      if (touched) touched->push_back(s);
      double hbox[9];
      int rs = fetch_box(pot, h_boxes, d_boxes, s, hbox);
      if (rs) return rs;
      if (host_Z_of && d_Z && n > 0) {
        // a homogeneous batch shares one copy of the atomic numbers: this member's own were never uploaded
        const int* src = host_Z_of(s);
        if (src) CU_TRY(pot, cudaMemcpyAsync(const_cast<int*>(d_Z) + o, src, sizeof(int) * n, cudaMemcpyHostToDevice, st));
      }
      rs = run_cell_path(pot, n, d_pos + 3 * o, d_Z ? d_Z + o : nullptr, d_F + 3 * o, d_E + s, hbox, st,
                         nullptr, false);
      if (rs >= kStatusCuda) return rs;
      Status stt{};
      if (rs == 0) {
        CU_TRY(pot, cudaMemcpyAsync(&stt, pot->status.p, sizeof stt, cudaMemcpyDeviceToHost, st));
        CU_TRY(pot, cudaStreamSynchronize(st));
        f = stt.flags;
        bad = stt.first_bad_atom;
      } else {
        f = kFlagSingularBox;
      }
    }
    int status = 0;
    if (f & ~(kFlagAtomShift | kFlagBigShift)) {
      if ((f & kFlagForeignZ) && bad != INT_MAX && bad >= 0) {
        const int* zsrc = host_Z_of ? host_Z_of(s) : nullptr;
        if (zsrc) {
          badz = zsrc[bad];
        } else if (d_Z) {
          CU_TRY(pot, cudaMemcpy(&badz, d_Z + o + bad, sizeof(int), cudaMemcpyDeviceToHost));
        }
      }
      status = map_flags(pot, f, bad, badz, host_pos ? host_pos + 3 * o : nullptr);
      if (status) {
        if (touched) touched->push_back(s);
        if (n > 0) CU_TRY(pot, cudaMemsetAsync(d_F + 3 * o, 0, sizeof(double) * 3 * n, st));
        CU_TRY(pot, cudaMemsetAsync(d_E + s, 0, sizeof(double), st));
        if (!first) {
          first = status;
          char pre[64];
          snprintf(pre, sizeof pre, "system %zu: ", s + pot->msg_sys_base);
          first_msg = pre + pot->last_error;
        }
      }
    }
    if (statuses) statuses[s] = status;
  }
  if (first) pot->last_error = first_msg;
  return first;
}

int rgpot_b200_force_batch_device(RgpotB200Pot* pot, size_t nSystems, const size_t* offsets,
                                  const double* d_positions, const int* d_atomicNrs,
                                  const double* d_boxes, double* d_forces, double* d_energies,
                                  void* cuda_stream) {
  if (!pot) return kStatusArgs;
  pot->last_error.clear();
  pot->pend.active = false;
  if (nSystems == 0) return 0;
  if (!offsets || !d_boxes || !d_positions || !d_forces || !d_energies)
    return fail(pot, kStatusArgs, "rgpot_b200_force_batch_device: null buffer");
  if (reads_species(pot->cfg.kind) && !d_atomicNrs)
    return fail(pot, kStatusArgs, "rgpot_b200: CuH2 needs atomic numbers");
  if (pot->cfg.kind == RGPOT_B200_ZBL)
    return fail(pot, kStatusArgs, "rgpot_b200: ZBL batches go through the host entry points "
                                  "(the species table is built from the atomic numbers on the host)");
  DevGuard guard(pot->device);
  CU_TRY(pot, guard.err);
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : pot->stream;
  int s = batch_device_impl(pot, nSystems, offsets, d_positions, d_atomicNrs, d_boxes, nullptr, d_forces,
                            d_energies, st);
  if (s) return s;
  RgpotB200Pot::Pending& pd = pot->pend;
  pd.active = true;
  pd.batch = true;
  pd.nsys = nSystems;
  pd.offsets.assign(offsets, offsets + nSystems + 1);
  pd.pos = d_positions;
  pd.Z = d_atomicNrs;
  pd.F = d_forces;
  pd.E = d_energies;
  pd.d_boxes = d_boxes;
  pd.st = st;
  return 0;
}

uint64_t rgpot_b200_hash64(const void* data, size_t len) {
  static const unsigned char none = 0;
  return rb::xxh3_host(len ? static_cast<const unsigned char*>(data) : &none, data ? len : 0);
}

// ---- result-cache keys (hash.cuh) ----------------------------------------------------------------
static int cache_keys_enqueue(RgpotB200Pot* pot, size_t nSystems, const size_t* offsets, const double* d_pos,
                              const int* d_Z, const double* d_boxes, uint64_t pot_type, uint64_t params_key,
                              uint64_t* d_keys, cudaStream_t st) {
  CU_TRY(pot, pot->b_off.reserve(sizeof(size_t) * (nSystems + 1)));
  CU_TRY(pot, cudaMemcpyAsync(pot->b_off.p, offsets, sizeof(size_t) * (nSystems + 1), cudaMemcpyHostToDevice, st));
  // Potential.hpp:330-334: the type as size_t and the parameter fingerprint, eight bytes each
  const uint64_t tail = rb::xxh3_host(reinterpret_cast<const unsigned char*>(&pot_type), 8) ^
                        rb::xxh3_host(reinterpret_cast<const unsigned char*>(&params_key), 8);
  rb::KeyArgs ka{pot->b_off.as<size_t>(), d_pos, d_Z, d_boxes, tail, d_keys, nSystems};
  const int warps = 8, per_block = 4 * warps;  // eight lanes per system
  k_cache_keys<<<(unsigned)((nSystems + per_block - 1) / per_block), 32 * warps, 0, st>>>(ka);
  LAUNCH_CHECK(pot, "k_cache_keys");
  return 0;
}

int rgpot_b200_cache_keys_device(RgpotB200Pot* pot, size_t nSystems, const size_t* offsets, const double* d_positions,
                                 const int* d_atomicNrs, const double* d_boxes, uint64_t pot_type, uint64_t params_key,
                                 uint64_t* d_keys, void* cuda_stream) {
  if (!pot) return kStatusArgs;
  pot->last_error.clear();
  if (nSystems == 0) return 0;
  if (!offsets || !d_positions || !d_atomicNrs || !d_boxes || !d_keys)
    return fail(pot, kStatusArgs, "rgpot_b200_cache_keys_device: null buffer");
  for (size_t i = 0; i < nSystems; ++i)
    if (offsets[i + 1] < offsets[i]) return fail(pot, kStatusArgs, "rgpot_b200_cache_keys: offsets must not decrease");
  DevGuard guard(pot->device);
  CU_TRY(pot, guard.err);
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : pot->stream;
  return cache_keys_enqueue(pot, nSystems, offsets, d_positions, d_atomicNrs, d_boxes, pot_type, params_key, d_keys, st);
}

int rgpot_b200_cache_keys(RgpotB200Pot* pot, size_t nSystems, const size_t* offsets, const double* positions,
                          const int* atomicNrs, const double* boxes, uint64_t pot_type, uint64_t params_key,
                          uint64_t* keys) {
  if (!pot) return kStatusArgs;
  pot->last_error.clear();
  if (nSystems == 0) return 0;
  if (!offsets || !positions || !atomicNrs || !boxes || !keys)
    return fail(pot, kStatusArgs, "rgpot_b200_cache_keys: null buffer");
  for (size_t i = 0; i < nSystems; ++i)
    if (offsets[i + 1] < offsets[i]) return fail(pot, kStatusArgs, "rgpot_b200_cache_keys: offsets must not decrease");
  DevGuard guard(pot->device);
  CU_TRY(pot, guard.err);
  cudaStream_t st = pot->stream;
  const size_t o0 = offsets[0], total = offsets[nSystems] - o0;
  CU_TRY(pot, pot->b_pos.reserve(sizeof(double) * 3 * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->b_Z.reserve(sizeof(int) * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->b_boxes.reserve(sizeof(double) * 9 * nSystems));
  CU_TRY(pot, pot->b_E.reserve(sizeof(uint64_t) * nSystems));
  if (total) {
    CU_TRY(pot, cudaMemcpyAsync(pot->b_pos.p, positions + 3 * o0, sizeof(double) * 3 * total, cudaMemcpyHostToDevice, st));
    CU_TRY(pot, cudaMemcpyAsync(pot->b_Z.p, atomicNrs + o0, sizeof(int) * total, cudaMemcpyHostToDevice, st));
  }
  CU_TRY(pot, cudaMemcpyAsync(pot->b_boxes.p, boxes, sizeof(double) * 9 * nSystems, cudaMemcpyHostToDevice, st));
  std::vector<size_t> rel(nSystems + 1);
  for (size_t i = 0; i <= nSystems; ++i) rel[i] = offsets[i] - o0;
  int s = cache_keys_enqueue(pot, nSystems, rel.data(), pot->b_pos.as<double>(), pot->b_Z.as<int>(),
                             pot->b_boxes.as<double>(), pot_type, params_key, pot->b_E.as<uint64_t>(), st);
  if (s) return s;
  CU_TRY(pot, cudaMemcpyAsync(keys, pot->b_E.p, sizeof(uint64_t) * nSystems, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaStreamSynchronize(st));
  return 0;
}

static int batch_sync_status(RgpotB200Pot* pot) {
  RgpotB200Pot::Pending& pd = pot->pend;
  int s = batch_finish(pot, pd.nsys, pd.offsets.data(), pd.pos, pd.Z, nullptr, pd.d_boxes, pd.F, pd.E,
                       pd.st, nullptr, {}, nullptr);
  cudaStreamSynchronize(pd.st);
  return s;
}

// ------------------------------------------------------------------------------------------------
// Copy/compute pipeline of a homogeneous (all-pairs-eligible) host batch: the batch is cut into
// chunks that flow through three streams, so the H2D copy of chunk k+1, the kernel of chunk k and
// the D2H copy of chunk k-1 overlap.  `stage_in(c0, c1)` is called right before chunk [c0, c1) is
// copied (the pointer-array entry packs the chunk into pinned staging there); `stage_out(c0, c1)`
// is called once the chunk's results are on the host.
// ------------------------------------------------------------------------------------------------
}  // extern "C" (templates need C++ linkage)

// `stage_in(c0, c1)` is called right before chunk [c0, c1) is copied (the pointer-array entry packs
// the chunk into pinned staging there) and returns whether every system of the chunk carries the
// atomic numbers `z0` of the batch's first system (then the chunk's own are not uploaded: the kernel
// reads the shared copy); `stage_out(c0, c1)` is called once the chunk's results are on the host --
// for chunk k - 2 while the copies and kernels of chunks k - 1 and k are in flight.
template <class StageIn, class StageOut>
static int batch_pipeline(RgpotB200Pot* pot, size_t nsys, const size_t* rel, long nmax, const double* h_pos,
                          const int* h_Z, const double* h_boxes, double* h_F, double* h_E, int* statuses,
                          const int* z0, const std::function<const int*(size_t)>& host_Z_of,
                          StageIn&& stage_in, StageOut&& stage_out) {
  const bool eam = reads_species(pot->cfg.kind) || pot->cfg.kind == RGPOT_B200_ZBL;  // has a Z array
  const size_t total = rel[nsys];
  CU_TRY(pot, pot->b_pos.reserve(sizeof(double) * 3 * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->b_F.reserve(sizeof(double) * 3 * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->b_E.reserve(sizeof(double) * nsys));
  CU_TRY(pot, pot->b_boxes.reserve(sizeof(double) * 9 * nsys));
  if (eam) CU_TRY(pot, pot->b_Z.reserve(sizeof(int) * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->b_off.reserve(sizeof(long long) * (nsys + 1)));
  CU_TRY(pot, pot->h_off.reserve(sizeof(long long) * (nsys + 1)));
  CU_TRY(pot, pot->b_flags.reserve(sizeof(unsigned) * nsys));
  CU_TRY(pot, pot->b_bad.reserve(sizeof(int) * 2 * nsys));
  long long* ho = pot->h_off.as<long long>();
  for (size_t s = 0; s <= nsys; ++s) ho[s] = (long long)rel[s];
  cudaStream_t st0 = pot->stream;
  CU_TRY(pot, cudaMemcpyAsync(pot->b_off.p, ho, sizeof(long long) * (nsys + 1), cudaMemcpyHostToDevice, st0));
  const size_t n0 = rel[1] - rel[0];
  if (!eam || n0 == 0) z0 = nullptr;
  if (z0) {
    CU_TRY(pot, pot->b_Z0.reserve(sizeof(int) * n0));
    CU_TRY(pot, cudaMemcpyAsync(pot->b_Z0.p, z0, sizeof(int) * n0, cudaMemcpyHostToDevice, st0));
  }
  CU_TRY(pot, cudaEventRecord(pot->ev_ready, st0));
  for (auto& s : pot->pipe) CU_TRY(pot, cudaStreamWaitEvent(s, pot->ev_ready, 0));

  const int maxnb = uses_eam_small(pot->cfg.kind) ? eam_small_maxnb(pot, nmax) : 0;
  // chunks: about thirty-two per batch (the first upload and the last download are not hidden behind
  // anything: the shorter they are the better), but never so short that a kernel is a few partial waves
  size_t chunk = (nsys + 31) / 32;
  chunk = std::min<size_t>(std::max<size_t>(chunk, (size_t)8 * pot->sm_count), 8192);
  if (pot->opt_batch_chunk > 0) chunk = (size_t)pot->opt_batch_chunk;
  const size_t nchunks = (nsys + chunk - 1) / chunk;
  std::vector<cudaEvent_t> done(nchunks, nullptr);
  int rc = 0;
  double* d_pos = pot->b_pos.as<double>();
  int* d_Z = eam ? pot->b_Z.as<int>() : nullptr;
  double* d_boxes = pot->b_boxes.as<double>();
  double* d_F = pot->b_F.as<double>();
  double* d_E = pot->b_E.as<double>();
  size_t retired = 0;
  auto retire = [&](size_t upto) {  // unpack chunks [retired, upto) as their results arrive
    for (; retired < upto; ++retired) {
      if (!done[retired]) continue;
      if (rc == 0 && cudaEventSynchronize(done[retired]) == cudaSuccess) {
        const size_t c0 = retired * chunk, c1 = std::min(nsys, c0 + chunk);
        stage_out(c0, c1);
      }
      cudaEventDestroy(done[retired]);
      done[retired] = nullptr;
    }
  };
  for (size_t k = 0; k < nchunks && rc == 0; ++k) {
    const size_t c0 = k * chunk, c1 = std::min(nsys, c0 + chunk);
    const size_t a0 = rel[c0], a1 = rel[c1];
    cudaStream_t st = pot->pipe[k % 3];
    const bool shared = stage_in(c0, c1) && z0 != nullptr;
    auto cu = [&](cudaError_t err, const char* what) {
      if (err != cudaSuccess && rc == 0) {
        pot->last_error = std::string("CUDA failure: ") + cudaGetErrorString(err) + " at " + what;
        rc = kStatusCuda;
      }
    };
    if (a1 > a0) {
      cu(cudaMemcpyAsync(d_pos + 3 * a0, h_pos + 3 * a0, sizeof(double) * 3 * (a1 - a0), cudaMemcpyHostToDevice, st), "H2D positions");
      if (eam && !shared) cu(cudaMemcpyAsync(d_Z + a0, h_Z + a0, sizeof(int) * (a1 - a0), cudaMemcpyHostToDevice, st), "H2D atomic numbers");
    }
    cu(cudaMemcpyAsync(d_boxes + 9 * c0, h_boxes + 9 * c0, sizeof(double) * 9 * (c1 - c0), cudaMemcpyHostToDevice, st), "H2D boxes");
    if (rc) break;
    rc = launch_small(pot, c1 - c0, nmax, pot->b_off.as<long long>() + c0, d_boxes + 9 * c0, d_pos,
                      shared ? pot->b_Z0.as<int>() : d_Z, d_F, d_E + c0, pot->b_flags.as<unsigned>() + c0,
                      pot->b_bad.as<int>() + 2 * c0, maxnb, st, nullptr, shared);
    if (rc) break;
    if (a1 > a0)
      cu(cudaMemcpyAsync(h_F + 3 * a0, d_F + 3 * a0, sizeof(double) * 3 * (a1 - a0), cudaMemcpyDeviceToHost, st), "D2H forces");
    cu(cudaMemcpyAsync(h_E + c0, d_E + c0, sizeof(double) * (c1 - c0), cudaMemcpyDeviceToHost, st), "D2H energies");
    cu(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming), "event");
    if (done[k]) cu(cudaEventRecord(done[k], st), "event record");
    if (k >= 2) retire(k - 1);  // chunk k - 2 is unpacked while chunks k - 1 and k are in flight
  }
  retire(nchunks);
  for (auto& s : pot->pipe) cudaStreamSynchronize(s);
  if (rc) return rc;

  // statuses; list-overflow members rerun on the cell path; failed members are zeroed
  std::vector<size_t> touched;
  int s = batch_finish(pot, nsys, rel, d_pos, d_Z, h_boxes, d_boxes, d_F, d_E, st0, h_pos, host_Z_of, statuses, &touched);
  if (s >= kStatusCuda) return s;
  for (size_t id : touched) {
    const size_t a0 = rel[id], a1 = rel[id + 1];
    if (a1 > a0)
      CU_TRY(pot, cudaMemcpyAsync(h_F + 3 * a0, d_F + 3 * a0, sizeof(double) * 3 * (a1 - a0), cudaMemcpyDeviceToHost, st0));
    CU_TRY(pot, cudaMemcpyAsync(h_E + id, d_E + id, sizeof(double), cudaMemcpyDeviceToHost, st0));
  }
  CU_TRY(pot, cudaStreamSynchronize(st0));
  if (!touched.empty()) {
    std::sort(touched.begin(), touched.end());
    for (size_t id : touched) stage_out(id, id + 1);
  }
  return s;
}

extern "C" {

// largest system of the batch if EVERY member can run on the all-pairs kernels, else -1
static long batch_all_small(const RgpotB200Pot* pot, size_t nsys, const size_t* rel) {
  if (pot->forced_path == 1) return -1;
  long nmax = 0;
  for (size_t s = 0; s < nsys; ++s) {
    const long n = (long)(rel[s + 1] - rel[s]);
    if (!small_eligible(pot, n)) return -1;
    nmax = std::max(nmax, n);
  }
  return nmax;
}

static const double kFreeBox[9] = {1e6, 0, 0, 0, 1e6, 0, 0, 0, 1e6};

}  // extern "C" (templates need C++ linkage)

// One batch on ONE device whose buffers are (or may be) pageable: every system is copied into pinned
// staging by the handle's helper threads, chunk by chunk, overlapped with the copies and kernels of the
// chunks before it, and the results are copied back the same way (ForceBatch semantics: the caller
// owns nSystems separate allocations, CppCore/rgpot/ForceStructs.hpp:40-54).
// pos_of / Z_of / box_of / F_of return the host pointers of system s (box_of may return NULL = free).
template <class PosOf, class ZOf, class BoxOf, class FOf>
static int batch_staged(RgpotB200Pot* pot, size_t nsys, const size_t* off, PosOf&& pos_of, ZOf&& Z_of,
                        BoxOf&& box_of, FOf&& F_of, double* energies, int* statuses) {
  const bool zbl = pot->cfg.kind == RGPOT_B200_ZBL;
  const bool eam = reads_species(pot->cfg.kind) || zbl;  // carries a per-atom integer array
  const size_t total = off[nsys];
  CU_TRY(pot, pot->h_pos.reserve(sizeof(double) * 3 * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->h_F.reserve(sizeof(double) * 3 * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->h_boxes.reserve(sizeof(double) * 9 * nsys));
  CU_TRY(pot, pot->h_E.reserve(sizeof(double) * nsys));
  if (eam) CU_TRY(pot, pot->h_Z.reserve(sizeof(int) * std::max<size_t>(total, 1)));
  double* hp = pot->h_pos.as<double>();
  double* hbx = pot->h_boxes.as<double>();
  int* hz = pot->h_Z.as<int>();
  double* hF = pot->h_F.as<double>();
  double* hE = pot->h_E.as<double>();
  rb::Team* team = get_team(pot);
  const size_t n0 = nsys ? off[1] - off[0] : 0;
  const int* z0 = (eam && !zbl && n0 > 0) ? Z_of(0) : nullptr;  // candidate for the shared copy
  auto pack = [&](size_t c0, size_t c1) -> bool {
    std::atomic<int> same{z0 ? 1 : 0};
    team->run(c1 - c0, 32, [&](size_t b, size_t e) {
      bool eq = true;
      for (size_t s = c0 + b; s < c0 + e; ++s) {
        const size_t n = off[s + 1] - off[s];
        if (n > 0) {
          rb::copy_stream(hp + 3 * off[s], pos_of(s), sizeof(double) * 3 * n);
          if (eam) {
            const int* z = Z_of(s);
            if (z0 && n == n0 && (z == z0 || memcmp(z, z0, sizeof(int) * n) == 0)) {
              // same species as the first system: nothing to copy if the whole chunk is like this
            } else {
              eq = false;
              memcpy(hz + off[s], z, sizeof(int) * n);
            }
          }
        } else if (eam) {
          eq = false;
        }
        const double* bx = box_of(s);
        memcpy(hbx + 9 * s, bx ? bx : kFreeBox, sizeof(double) * 9);
      }
      rb::copy_stream_fence();  // the DMA engine reads the staging next
      if (!eq) same.store(0, std::memory_order_relaxed);
    });
    if (eam && z0 && !same.load()) {  // a mixed chunk uploads its own atomic numbers: fill in the skipped ones
      team->run(c1 - c0, 64, [&](size_t b, size_t e) {
        for (size_t s = c0 + b; s < c0 + e; ++s) {
          const size_t n = off[s + 1] - off[s];
          if (n > 0) memcpy(hz + off[s], Z_of(s), sizeof(int) * n);
        }
      });
    }
    return same.load() != 0;
  };
  auto unpack = [&](size_t c0, size_t c1) {
    team->run(c1 - c0, 32, [&](size_t b, size_t e) {
      for (size_t s = c0 + b; s < c0 + e; ++s) {
        const size_t n = off[s + 1] - off[s];
        if (n > 0) rb::copy_stream(F_of(s), hF + 3 * off[s], sizeof(double) * 3 * n);
        energies[s] = hE[s];
      }
      rb::copy_stream_fence();
    });
  };
  std::function<const int*(size_t)> host_Z_of;
  if (eam) host_Z_of = [&](size_t s) -> const int* { return Z_of(s); };
  const long nmax = batch_all_small(pot, nsys, off);
  if (nmax >= 0 && !zbl)
    return batch_pipeline(pot, nsys, off, nmax, hp, eam ? hz : nullptr, hbx, hF, hE, statuses, z0, host_Z_of,
                          pack, unpack);
  // mixed or large members: pack everything, run, unpack everything
  z0 = nullptr;
  pack(0, nsys);
  int st = rgpot_b200_force_batch_packed(pot, nsys, off, hp, eam ? hz : nullptr, hbx, hF, hE, statuses);
  if (st >= kStatusCuda) return st;
  unpack(0, nsys);
  return st;
}

// cut [0, nsys) into `parts` contiguous ranges balanced by the atom count
static std::vector<size_t> cut_by_atoms(size_t nsys, const std::function<size_t(size_t)>& atoms_before, size_t parts) {
  std::vector<size_t> cut(parts + 1, nsys);
  cut[0] = 0;
  const size_t total = atoms_before(nsys);
  size_t s = 0;
  for (size_t p = 1; p < parts; ++p) {
    const size_t target = (size_t)((double)total * (double)p / (double)parts);
    while (s < nsys && atoms_before(s) < target) ++s;
    // nearest boundary to the target
    if (s > cut[p - 1] && s <= nsys && target - atoms_before(s - 1) < atoms_before(s) - target && s - 1 >= cut[p - 1]) --s;
    cut[p] = std::max(s, cut[p - 1]);
  }
  return cut;
}

// run `job(shard index, first system, one past the last)` on every device's host thread and join;
// returns the first non-zero status in batch order (its message becomes the handle's)
template <class Job>
static int run_on_shards(RgpotB200Pot* pot, const std::vector<size_t>& cut, Job&& job) {
  const size_t parts = cut.size() - 1;
  for (size_t p = 0; p < parts; ++p) {
    auto& sh = pot->shards[p];
    sh.status = 0;
    if (cut[p + 1] == cut[p]) continue;
    sh.pot->msg_sys_base = cut[p];
    sh.worker->submit([&, p] {
      auto& me = pot->shards[p];
      DevGuard g(me.pot->device);
      me.status = job(me.pot, cut[p], cut[p + 1]);
    });
  }
  for (size_t p = 0; p < parts; ++p)
    if (cut[p + 1] > cut[p]) pot->shards[p].worker->wait();
  int first = 0;
  for (size_t p = 0; p < parts; ++p) {
    const auto& sh = pot->shards[p];
    if (sh.status >= kStatusCuda) {  // an engine failure outranks per-system statuses
      pot->last_error = sh.pot->last_error;
      return sh.status;
    }
    if (sh.status && !first) {
      first = sh.status;
      pot->last_error = sh.pot->last_error;
    }
  }
  return first;
}

// how many of the handle's devices a batch of this size is worth cutting across
static size_t shards_for(const RgpotB200Pot* pot, size_t nsys, size_t total_atoms) {
  if (pot->shards.size() < 2) return 1;
  size_t parts = std::min(pot->shards.size(), nsys);
  parts = std::min(parts, std::max<size_t>(1, total_atoms / 16384));
  return std::max<size_t>(parts, 1);
}

extern "C" {

static int force_batch_packed_one(RgpotB200Pot* pot, size_t nSystems, const size_t* offsets,
                                  const double* positions, const int* atomicNrs, const double* boxes,
                                  double* forces, double* energies, int* statuses);

int rgpot_b200_force_batch_packed(RgpotB200Pot* pot, size_t nSystems, const size_t* offsets,
                                  const double* positions, const int* atomicNrs, const double* boxes,
                                  double* forces, double* energies, int* statuses) {
  if (!pot) return kStatusArgs;
  pot->last_error.clear();
  if (nSystems == 0) return 0;
  if (!offsets || !positions || !boxes || !forces || !energies)
    return fail(pot, kStatusArgs, "rgpot_b200_force_batch_packed: null buffer");
  const bool zbl = pot->cfg.kind == RGPOT_B200_ZBL;
  const bool eam = reads_species(pot->cfg.kind) || zbl;  // "carries a per-atom integer array"
  if (eam && !atomicNrs) return fail(pot, kStatusArgs, "rgpot_b200: this potential needs atomic numbers");
  for (size_t i = 0; i < nSystems; ++i)
    if (offsets[i + 1] < offsets[i]) return fail(pot, kStatusArgs, "rgpot_b200_force_batch_packed: offsets must not decrease");
  const size_t parts = shards_for(pot, nSystems, offsets[nSystems] - offsets[0]);
  if (parts > 1 && !zbl) {
    // one range per device, each driven by its own host thread; offsets / positions / forces are
    // indexed absolutely, so a range is the same arrays with a later first system
    const size_t base = offsets[0];
    const std::vector<size_t> cut =
        cut_by_atoms(nSystems, [&](size_t s) { return offsets[s] - base; }, parts);
    return run_on_shards(pot, cut, [&](RgpotB200Pot* sp, size_t s0, size_t s1) {
      return force_batch_packed_one(sp, s1 - s0, offsets + s0, positions, atomicNrs, boxes + 9 * s0, forces,
                                    energies + s0, statuses ? statuses + s0 : nullptr);
    });
  }
  RgpotB200Pot* target = pot->shards.empty() ? pot : pot->shards[0].pot;
  DevGuard g(target->device);
  if (g.err != cudaSuccess) return fail(pot, kStatusCuda, std::string("CUDA failure: ") + cudaGetErrorString(g.err));
  target->msg_sys_base = 0;
  const int st = force_batch_packed_one(target, nSystems, offsets, positions, atomicNrs, boxes, forces, energies, statuses);
  if (target != pot) pot->last_error = target->last_error;
  return st;
}

static int force_batch_packed_one(RgpotB200Pot* pot, size_t nSystems, const size_t* offsets,
                                  const double* positions, const int* atomicNrs, const double* boxes,
                                  double* forces, double* energies, int* statuses) {
  pot->last_error.clear();
  const bool zbl = pot->cfg.kind == RGPOT_B200_ZBL;
  const bool eam = reads_species(pot->cfg.kind) || zbl;
  cudaStream_t st = pot->stream;
  const size_t base = offsets[0];
  const size_t total = offsets[nSystems] - base;
  std::vector<size_t> rel(nSystems + 1);
  for (size_t s = 0; s <= nSystems; ++s) rel[s] = offsets[s] - base;
  std::vector<int> zbl_types;
  if (zbl) {
    // one table over the union of the batch's species: coefficients only depend on the pair
    zbl_types.resize(std::max<size_t>(total, 1));
    int zs = zbl_prepare(pot, total, atomicNrs + base, zbl_types.data(), st);
    if (zs) return zs;
    atomicNrs = zbl_types.data() - base;
  }

  const long nmax = batch_all_small(pot, nSystems, rel.data());
  // page-locked arrays are copied by the DMA engines straight from / into the caller's memory;
  // pageable ones go through the handle's pinned staging like a ForceBatch does
  const bool pinned = rb::host_ptr_is_pinned(positions + 3 * base) && rb::host_ptr_is_pinned(forces + 3 * base) &&
                      (!eam || zbl || rb::host_ptr_is_pinned(atomicNrs + base)) && rb::host_ptr_is_pinned(boxes) &&
                      rb::host_ptr_is_pinned(energies);
  if (nmax >= 0 && !zbl && !pinned && total >= 4096) {
    return batch_staged(pot, nSystems, rel.data(),
                        [&](size_t s) { return positions + 3 * offsets[s]; },
                        [&](size_t s) { return atomicNrs + offsets[s]; },
                        [&](size_t s) { return boxes + 9 * s; },
                        [&](size_t s) { return forces + 3 * offsets[s]; }, energies, statuses);
  }
  if (nmax >= 0) {
    const size_t n0 = rel[1] - rel[0];
    const int* z0 = (eam && !zbl && n0 > 0) ? atomicNrs + base : nullptr;
    rb::Team* team = (z0 && nSystems >= 1024) ? get_team(pot) : nullptr;
    // homogeneous batches (NEB / dimer images: one species array repeated): the kernel reads ONE copy
    auto check = [&](size_t c0, size_t c1) -> bool {
      if (!team) return false;
      std::atomic<int> same{1};
      team->run(c1 - c0, 128, [&](size_t b, size_t e) {
        for (size_t s = c0 + b; s < c0 + e; ++s)
          if (rel[s + 1] - rel[s] != n0 || memcmp(atomicNrs + offsets[s], z0, sizeof(int) * n0) != 0) {
            same.store(0, std::memory_order_relaxed);
            return;
          }
      });
      return same.load() != 0;
    };
    auto nop = [](size_t, size_t) {};
    std::function<const int*(size_t)> host_Z_of;
    if (eam) host_Z_of = [&](size_t s) -> const int* { return atomicNrs + offsets[s]; };
    return batch_pipeline(pot, nSystems, rel.data(), nmax, positions + 3 * base,
                          atomicNrs ? atomicNrs + base : nullptr, boxes, forces + 3 * base, energies,
                          statuses, team ? z0 : nullptr, host_Z_of, check, nop);
  }

  CU_TRY(pot, pot->b_pos.reserve(sizeof(double) * 3 * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->b_F.reserve(sizeof(double) * 3 * std::max<size_t>(total, 1)));
  CU_TRY(pot, pot->b_E.reserve(sizeof(double) * nSystems));
  CU_TRY(pot, pot->b_boxes.reserve(sizeof(double) * 9 * nSystems));
  CU_TRY(pot, cudaMemcpyAsync(pot->b_pos.p, positions + 3 * base, sizeof(double) * 3 * total,
                              cudaMemcpyHostToDevice, st));
  const int* d_Z = nullptr;
  if (eam) {
    CU_TRY(pot, pot->b_Z.reserve(sizeof(int) * std::max<size_t>(total, 1)));
    CU_TRY(pot, cudaMemcpyAsync(pot->b_Z.p, atomicNrs + base, sizeof(int) * total, cudaMemcpyHostToDevice, st));
    d_Z = pot->b_Z.as<int>();
  }
  CU_TRY(pot, cudaMemcpyAsync(pot->b_boxes.p, boxes, sizeof(double) * 9 * nSystems, cudaMemcpyHostToDevice, st));
  int s = batch_device_impl(pot, nSystems, rel.data(), pot->b_pos.as<double>(), d_Z,
                            pot->b_boxes.as<double>(), boxes, pot->b_F.as<double>(), pot->b_E.as<double>(), st);
  if (s) return s;
  std::function<const int*(size_t)> host_Z_of;
  if (eam) host_Z_of = [&](size_t k) -> const int* { return atomicNrs + offsets[k]; };
  s = batch_finish(pot, nSystems, rel.data(), pot->b_pos.as<double>(), d_Z, boxes,
                   pot->b_boxes.as<double>(), pot->b_F.as<double>(), pot->b_E.as<double>(), st,
                   positions + 3 * base, host_Z_of, statuses);
  if (s >= kStatusCuda) return s;
  CU_TRY(pot, cudaMemcpyAsync(forces + 3 * base, pot->b_F.p, sizeof(double) * 3 * total,
                              cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaMemcpyAsync(energies, pot->b_E.p, sizeof(double) * nSystems, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaStreamSynchronize(st));
  return s;
}

static int force_batch_one(RgpotB200Pot* pot, size_t nSystems, const size_t* nAtoms,
                           const double* const* positions, const int* const* atomicNrs,
                           const double* const* boxes, double* const* forces, double* energies,
                           int* statuses) {
  pot->last_error.clear();
  std::vector<size_t> off(nSystems + 1, 0);
  for (size_t s = 0; s < nSystems; ++s) off[s + 1] = off[s] + nAtoms[s];
  return batch_staged(pot, nSystems, off.data(), [&](size_t s) { return positions[s]; },
                      [&](size_t s) { return atomicNrs[s]; },
                      [&](size_t s) -> const double* { return boxes ? boxes[s] : nullptr; },
                      [&](size_t s) { return forces[s]; }, energies, statuses);
}

int rgpot_b200_force_batch(RgpotB200Pot* pot, size_t nSystems, const size_t* nAtoms,
                           const double* const* positions, const int* const* atomicNrs,
                           const double* const* boxes, double* const* forces, double* energies,
                           int* statuses) {
  if (!pot) return kStatusArgs;
  pot->last_error.clear();
  if (nSystems == 0) return 0;
  if (!nAtoms || !positions || !forces || !energies)
    return fail(pot, kStatusArgs, "rgpot_b200_force_batch: null buffer");
  const bool zbl = pot->cfg.kind == RGPOT_B200_ZBL;
  const bool eam = reads_species(pot->cfg.kind) || zbl;  // carries a per-atom integer array
  if (eam && !atomicNrs) return fail(pot, kStatusArgs, "rgpot_b200: this potential needs atomic numbers");
  std::vector<size_t> before(nSystems + 1, 0);
  for (size_t s = 0; s < nSystems; ++s) {
    before[s + 1] = before[s] + nAtoms[s];
    if (nAtoms[s] > 0 && (!positions[s] || !forces[s] || (eam && !atomicNrs[s])))
      return fail(pot, kStatusArgs, "rgpot_b200_force_batch: null system buffer");
  }
  const size_t parts = shards_for(pot, nSystems, before[nSystems]);
  if (parts > 1 && !zbl) {
    const std::vector<size_t> cut = cut_by_atoms(nSystems, [&](size_t s) { return before[s]; }, parts);
    return run_on_shards(pot, cut, [&](RgpotB200Pot* sp, size_t s0, size_t s1) {
      return force_batch_one(sp, s1 - s0, nAtoms + s0, positions + s0, atomicNrs ? atomicNrs + s0 : nullptr,
                             boxes ? boxes + s0 : nullptr, forces + s0, energies + s0,
                             statuses ? statuses + s0 : nullptr);
    });
  }
  RgpotB200Pot* target = pot->shards.empty() ? pot : pot->shards[0].pot;
  DevGuard g(target->device);
  if (g.err != cudaSuccess) return fail(pot, kStatusCuda, std::string("CUDA failure: ") + cudaGetErrorString(g.err));
  target->msg_sys_base = 0;
  const int st = force_batch_one(target, nSystems, nAtoms, positions, atomicNrs, boxes, forces, energies, statuses);
  if (target != pot) pot->last_error = target->last_error;
  return st;
}

// ------------------------------------------------------------------------------------------------
// exact drop-in for the Fortran entry (FortranPots.cc:33-40): process-wide default handle,
// thread-local error text (rgpot_ferror.f90 keeps one process-global buffer).
// ------------------------------------------------------------------------------------------------
static std::mutex g_cuh2_mu;
static RgpotB200Pot* g_cuh2 = nullptr;
// the handle-less entries take their GPU from the environment (RGPOT_B200_DEVICE, default 0)
static void dropin_device(RgpotB200Config* cfg) {
  if (const char* e = getenv("RGPOT_B200_DEVICE")) {
    char* end = nullptr;
    const long d = strtol(e, &end, 10);
    if (end != e && d >= 0 && d < 1024) cfg->device = (int)d;
  }
}
static thread_local std::string g_last_fortran_error;

int rgpot_cuh2_force(int32_t natoms, const double* positions, const int32_t* atomic_numbers,
                     const double* cell, double* forces, double* energy) {
  std::lock_guard<std::mutex> lk(g_cuh2_mu);
  g_last_fortran_error.clear();
  if (!g_cuh2) {
    RgpotB200Config cfg;
    rgpot_b200_default_config(RGPOT_B200_CUH2, &cfg);
    dropin_device(&cfg);
    char err[256] = {0};
    g_cuh2 = rgpot_b200_create(&cfg, err, sizeof err);
    if (!g_cuh2) {
      g_last_fortran_error = err;
      if (energy) *energy = 0.0;
      if (forces && natoms > 0) memset(forces, 0, sizeof(double) * 3 * (size_t)natoms);
      return kStatusCuda;
    }
  }
  int s = rgpot_b200_force(g_cuh2, natoms, positions, atomic_numbers, forces, energy, nullptr, cell);
  if (s) g_last_fortran_error = g_cuh2->last_error;
  return s;
}

// FortranPots.cc:26-28, rgpot_eam_al_capi.f90:31-58: no atomic numbers in this signature
static RgpotB200Pot* g_eam_al = nullptr;

int rgpot_eam_al_force(int32_t natoms, const double* positions, const double* cell, double* forces,
                       double* energy) {
  std::lock_guard<std::mutex> lk(g_cuh2_mu);
  g_last_fortran_error.clear();
  if (!g_eam_al) {
    RgpotB200Config cfg;
    rgpot_b200_default_config(RGPOT_B200_EAM_AL, &cfg);
    dropin_device(&cfg);
    char err[256] = {0};
    g_eam_al = rgpot_b200_create(&cfg, err, sizeof err);
    if (!g_eam_al) {
      g_last_fortran_error = err;
      if (energy) *energy = 0.0;
      if (forces && natoms > 0) memset(forces, 0, sizeof(double) * 3 * (size_t)natoms);
      return kStatusCuda;
    }
  }
  int s = rgpot_b200_force(g_eam_al, natoms, positions, nullptr, forces, energy, nullptr, cell);
  if (s) g_last_fortran_error = g_eam_al->last_error;
  return s;
}

// FortranPots.cc: rgpot_fehe_force(natoms, positions, atomic_numbers, cell, forces, energy), rgpot_fehe_capi.f90:36-63
static RgpotB200Pot* g_fehe = nullptr;

int rgpot_fehe_force(int32_t natoms, const double* positions, const int32_t* atomic_numbers, const double* cell,
                     double* forces, double* energy) {
  std::lock_guard<std::mutex> lk(g_cuh2_mu);
  g_last_fortran_error.clear();
  if (!g_fehe) {
    RgpotB200Config cfg;
    rgpot_b200_default_config(RGPOT_B200_FEHE, &cfg);
    dropin_device(&cfg);
    char err[256] = {0};
    g_fehe = rgpot_b200_create(&cfg, err, sizeof err);
    if (!g_fehe) {
      g_last_fortran_error = err;
      if (energy) *energy = 0.0;
      if (forces && natoms > 0) memset(forces, 0, sizeof(double) * 3 * (size_t)natoms);
      return kStatusCuda;
    }
  }
  int s = rgpot_b200_force(g_fehe, natoms, positions, atomic_numbers, forces, energy, nullptr, cell);
  if (s) g_last_fortran_error = g_fehe->last_error;
  return s;
}

int rgpot_fortran_last_error(char* buffer, int buffer_len) {
  if (!buffer || buffer_len <= 0) return 0;
  const int n = (int)std::min<size_t>(g_last_fortran_error.size(), (size_t)buffer_len - 1);
  memcpy(buffer, g_last_fortran_error.data(), n);
  buffer[n] = 0;
  return n;
}

}  // extern "C"

#include "core_abi.cuh"

#ifdef RB_STAMPS
// tuning builds only: the stage stamps of the last single-system kernel (nanoseconds, %globaltimer)
extern "C" __attribute__((visibility("default"))) int rgpot_b200_debug_stamps(unsigned long long* out, int n) {
  if (n > 32) n = 32;
  return (int)cudaMemcpyFromSymbol(out, rb::g_stamps, sizeof(unsigned long long) * (size_t)n);
}
#endif
