/**
 * @file rgpot_b200.h
 * @brief C ABI of the B200 (sm_100a) CUDA engine for rgpot's LJ / LJCluster / CuH2 hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types.  Host units are
 * those of rgpot: positions and box in Angstrom, energies in eV, forces in eV/Angstrom;
 * positions and forces are x/y/z interleaved, the box is nine doubles, row-major, one cell
 * vector per row (reference: CppCore/rgpot/ForceStructs.hpp:20-36).
 *
 * Which reference interface each entry point replaces:
 *
 *  - rgpot_b200_abi_version / _create / _destroy / _force / _available follow the reference's
 *    dlopen-engine ABI (CppCore/rgpot/XTBPot/xtb_c_abi.h:49-65,
 *    CppCore/rgpot/MetatomicPot/metatomic_c_abi.h:45-61): opaque handle, 0 = success,
 *    errbuf at create, `variance` may be NULL.  _force is what
 *    `LJPot::forceImpl` (CppCore/rgpot/LennardJones/LJPot.cc:38-84),
 *    `LJClusterPot::forceImpl` (LJClusterPot.cc:33-83) and `CuH2Pot::forceImpl`
 *    (CppCore/rgpot/fortran/FortranPots.cc:78-90) compute.
 *  - rgpot_b200_force_batch is the native implementation of
 *    `Potential<D>::forceBatchImpl(const ForceBatch&)` (CppCore/rgpot/Potential.hpp:235-240,
 *    ForceStructs.hpp:50-54); the reference only has the per-system default loop.
 *  - rgpot_cuh2_force / rgpot_fortran_last_error are the exact symbols the reference's C++ face
 *    binds for the Fortran kernel (FortranPots.cc:33-40, rgpot_cuh2_capi.f90:31-58,
 *    rgpot_ferror.f90:41-56): link librgpot_b200.so instead of the Fortran objects and
 *    `fortranpots::CuH2Pot` runs on the GPU unmodified.
 *  - the *_device entry points take device pointers ("any device" in
 *    rgpot-core/include/rgpot.h:78-88) and never touch host memory on the data path.
 *
 * Status codes (rgpot_cuh2.f90:404-429, rgpot_neighbors.f90:111-115):
 *   0 success; 1 neighbor-search failure (singular box, non-finite coordinate) or size error;
 *   2 an atomic number outside {29, 1} (CuH2 only); 3 two atoms occupy the same position
 *   (CuH2 only); >= 100 CUDA / engine failure.  On any non-zero status the outputs are zeroed,
 *   like the reference, and the message is available from rgpot_b200_last_error.
 *
 * There is no CPU fallback: every entry point fails with a CUDA status if no B200-class
 * device is usable.
 */
#ifndef RGPOT_B200_H
#define RGPOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGPOT_B200_API __attribute__((visibility("default")))
#define RGPOT_B200_ABI_VERSION 3

typedef struct RgpotB200Pot RgpotB200Pot;

/** Potential selector; the names follow rgpot's PotType (CppCore/rgpot/pot_types.hpp:26-45). */
enum {
  RGPOT_B200_LJ = 0,         /**< PotType::LJ        -- periodic, orthorhombic nearest image */
  RGPOT_B200_LJ_CLUSTER = 1, /**< PotType::LJCluster -- free boundaries                       */
  RGPOT_B200_CUH2 = 2,       /**< PotType::CuH2      -- Cu/H EAM, all images inside 6.1 A     */
  RGPOT_B200_MORSE = 3,      /**< PotType::Morse     -- shifted Morse pair potential, same pair search as
                                  LJ (CppCore/rgpot/Morse/MorsePot.{hpp,cc}); SURVEY 8(f) rank 1 */
  RGPOT_B200_ZBL = 4,        /**< PotType::ZBL       -- screened nuclear repulsion with switching, free
                                  boundaries, species-pair tables (CppCore/rgpot/ZBL/ZBLPot.{hpp,cc}) */
  RGPOT_B200_EAM_AL = 5,     /**< PotType::EAMAl     -- double-exponential aluminium EAM, single species, all images
                                  inside 7.1 A (CppCore/rgpot/fortran/rgpot_eam_al.f90); SURVEY 8(f) rank 2.
                                  Atomic numbers are not read (rgpot_eam_al_capi.f90:31-35).            */
  RGPOT_B200_FEHE = 6        /**< PotType::FeHe      -- Fe-He EAM (Ackland iron + s-band helium coupling), two density
                                  channels, all images inside 5.4 A (CppCore/rgpot/fortran/rgpot_fehe.f90);
                                  atomic numbers 26 / 2, anything else is status 2.  Gather kernels only.  */
};

typedef struct RgpotB200Config {
  int kind;      /**< RGPOT_B200_LJ / _LJ_CLUSTER / _CUH2 / _MORSE / _ZBL / _EAM_AL / _FEHE */
  int device;    /**< CUDA device ordinal; -1 keeps the calling thread's current device;
                      -2: every usable device of the box (same as n_devices = all)             */
  double u0;     /**< LJConfig::u0     (LJPot.hpp:30-34), default 1.0  */
  double cutoff; /**< LJConfig::cutoff, default 15.0; ignored for CuH2 (fixed 6.1) and EAM-Al (7.1) */
  double psi;    /**< LJConfig::psi,    default 1.0  */
  /** 0: reference semantics (box diagonal only, one nearest image per pair, LJPot.cc:35).
   *  1: opt-in general triclinic LJ: every periodic image with d2 < rc2 (vesin semantics);
   *     no reference counterpart exists for this mode.                                      */
  int lj_all_images;
  /** Batches on several GPUs of one box (the systems of a ForceBatch are independent,
   *  CppCore/rgpot/ForceStructs.hpp:42-48, Potential.hpp:231-233): with n_devices > 1 (or device = -2)
   *  ONE rgpot_b200_force_batch[_packed] call cuts the batch into contiguous ranges balanced by the
   *  atom count, one per device of `devices`, and drives each range from its own host thread
   *  (streams, pinned staging and helper threads on the CPUs next to that GPU), results written
   *  straight into the caller's buffers; no collective.  Single systems and the *_device entry points
   *  use devices[0].  0 or 1: single device (`device`).                                        */
  int n_devices;
  /** Host threads that move ForceBatch buffers to and from pinned staging, per device (0 = automatic:
   *  the usable CPUs divided by the number of devices, at most 12). */
  int host_threads;
  int reserved[5];
  /** MorseConfig (Morse/MorsePot.hpp:31-36): De, a, re; the cutoff goes in `cutoff` (default 9.5).
   *  Defaults 0.7102 eV, 1.6047 1/A, 2.8970 A (platinum).  Unused by the other kinds.        */
  double morse_De, morse_a, morse_re;
  /** ZBLConfig (ZBL/ZBLPot.hpp:29-32): cut_inner here, cut_global in `cutoff` (defaults 2.0 / 2.5);
   *  create fails unless 0 < cut_inner < cut_global (ZBLPot.cc:76-79).  At most 16 distinct
   *  atomic numbers per call.                                                                */
  double zbl_cut_inner;
  int devices[16]; /**< CUDA ordinals of the batch devices when n_devices > 1 */
} RgpotB200Config;

RGPOT_B200_API int rgpot_b200_abi_version(void);
/** Number of usable CUDA devices of compute capability >= 10.0 (0 if none). */
RGPOT_B200_API int rgpot_b200_available(void);
/** Fill `cfg` with the reference defaults for `kind`. */
RGPOT_B200_API void rgpot_b200_default_config(int kind, RgpotB200Config *cfg);

RGPOT_B200_API RgpotB200Pot *rgpot_b200_create(const RgpotB200Config *cfg, char *errbuf,
                                               size_t errlen);
RGPOT_B200_API void rgpot_b200_destroy(RgpotB200Pot *pot);

/** Energy + forces of one system from HOST buffers.  `forces` (3 * nAtoms) is overwritten. */
RGPOT_B200_API int rgpot_b200_force(RgpotB200Pot *pot, long nAtoms, const double *positions,
                                    const int *atomicNrs, double *forces, double *energy,
                                    double *variance, const double *box);

/** Many independent systems in one call, HOST buffers, parallel arrays like ForceBatch.
 *  forces[i] must hold 3 * nAtoms[i] doubles; energies holds nSystems doubles.
 *  `statuses` (may be NULL) receives one status per system; the return value is the first
 *  non-zero status (0 if all succeeded).                                                    */
RGPOT_B200_API int rgpot_b200_force_batch(RgpotB200Pot *pot, size_t nSystems,
                                          const size_t *nAtoms, const double *const *positions,
                                          const int *const *atomicNrs, const double *const *boxes,
                                          double *const *forces, double *energies, int *statuses);

/** Same batch with the systems stored back to back: system s owns atoms
 *  [offsets[s], offsets[s+1]) of `positions` / `atomicNrs` / `forces` and boxes[9*s .. 9*s+8].
 *  Buffers allocated with rgpot_b200_host_alloc are copied asynchronously.                  */
RGPOT_B200_API int rgpot_b200_force_batch_packed(RgpotB200Pot *pot, size_t nSystems,
                                                 const size_t *offsets, const double *positions,
                                                 const int *atomicNrs, const double *boxes,
                                                 double *forces, double *energies, int *statuses);

/** Device-resident variants: every pointer except `box` (9 host doubles) and `offsets` (host
 *  array of nSystems + 1 atom offsets) is a device pointer on the handle's device; work is
 *  enqueued on `cuda_stream` (a cudaStream_t, NULL = the handle's own stream) and the call
 *  returns without synchronising.  Status flags are checked by rgpot_b200_sync_status.
 *  Lifetime: d_positions / d_atomicNrs / d_forces / d_energy must stay allocated and unchanged until
 *  rgpot_b200_sync_status has returned (a system the all-pairs kernel cannot hold is re-run from
 *  them there), and a handle carries ONE call in flight: the next *_device call on the same handle
 *  reuses its scratch buffers.                                                               */
RGPOT_B200_API int rgpot_b200_force_device(RgpotB200Pot *pot, long nAtoms, const double *d_positions,
                                           const int *d_atomicNrs, double *d_forces,
                                           double *d_energy, const double *box, void *cuda_stream);
RGPOT_B200_API int rgpot_b200_force_batch_device(RgpotB200Pot *pot, size_t nSystems,
                                                 const size_t *offsets, const double *d_positions,
                                                 const int *d_atomicNrs, const double *d_boxes,
                                                 double *d_forces, double *d_energies,
                                                 void *cuda_stream);
/** Wait for the last *_device call and return its status (0 / 1 / 2 / 3 / >= 100). */
RGPOT_B200_API int rgpot_b200_sync_status(RgpotB200Pot *pot);

/** Neighbor table the engine acts on, for pair-set parity checks and list consumers:
 *  writes up to `cap` entries (i, j) into pairs[2k], pairs[2k+1] and the periodic shift S into
 *  shifts[3k..3k+2] (vec = r_j - r_i + S.H), in unspecified order; *n_out receives the total
 *  count.  CuH2 / all-image mode: directed entries, self images dropped
 *  (rgpot_neighbors.f90:126-134).  LJ reference mode: unordered pairs i < j, shifts zero.
 *  `path`: 0 = automatic, 1 = force the cell-list kernels, 2 = force the all-pairs kernels.  */
RGPOT_B200_API int rgpot_b200_neighbors(RgpotB200Pot *pot, long nAtoms, const double *positions,
                                        const int *atomicNrs, const double *box, int path,
                                        long cap, int32_t *pairs, int32_t *shifts, long *n_out);

/** A neighbor list in the shape of vesin's VesinNeighborList (third_party/vesin/vesin.h:111-157):
 *  `length` pairs (i, j), the periodic shift S of each, the vector (p_j - p_i) + S.H and its
 *  length.  Host arrays owned by the struct: release with rgpot_b200_neighbor_list_free.          */
typedef struct RgpotB200NeighborList {
  size_t length;
  size_t *pairs;     /**< 2 * length */
  int32_t *shifts;   /**< 3 * length */
  double *distances; /**< length */
  double *vectors;   /**< 3 * length */
} RgpotB200NeighborList;

/** vesin_neighbors on the GPU cell list (lists for model-driven potentials such
 *  as MetatomicPot, CppCore/rgpot/MetatomicPot/MetatomicPot.cc:487-542): every pair of points and
 *  periodic images with d2 < cutoff^2, evaluated with vesin's own expression
 *  (vesin-single-build.cpp:1199-1202), so the pair set, the vectors and the distances are
 *  bit-identical to vesin's.  `box`: nine doubles, one cell vector per row; `periodic`: three flags,
 *  all set or all clear (NULL = periodic); `full`: both i->j and j->i, else vesin's half list
 *  (:1165-1197); `sorted`: ordered by (i, j, shift).  Any handle serves (it supplies the device,
 *  the stream and the scratch buffers); `out` must be zero-initialised or a previous result.      */
RGPOT_B200_API int rgpot_b200_neighbor_list(RgpotB200Pot *pot, size_t n_points, const double *points,
                                            const double *box, const int *periodic, double cutoff,
                                            int full, int sorted, RgpotB200NeighborList *out);
RGPOT_B200_API void rgpot_b200_neighbor_list_free(RgpotB200NeighborList *nl);

/** Result-cache keys of a packed batch: keys[i] is bit-identical to
 *  Potential<D>::cacheKey(in[i]) of the reference (CppCore/rgpot/Potential.hpp:324-336) --
 *  XXH3_64bits (xxHash 0.8.3, subprojects/xxhash.wrap) of the positions, the atomic numbers and the
 *  box, XOR the hashes of `pot_type` (the PotType enumerator as size_t) and `params_key`
 *  (paramsKey()) -- so the hit / miss compaction of Potential::forceBatch (:251-306) can hash a
 *  whole batch in one launch; with the CPU potential's params_key the keys address the entries the
 *  CPU path wrote, with the plugin classes' own (salted) paramsKey() they stay apart.  Host buffers;
 *  the _device variant takes device pointers for positions / atomicNrs / boxes / keys (offsets stay
 *  on the host), enqueues on `cuda_stream` and does not synchronise.                          */
RGPOT_B200_API int rgpot_b200_cache_keys(RgpotB200Pot *pot, size_t nSystems, const size_t *offsets,
                                         const double *positions, const int *atomicNrs,
                                         const double *boxes, uint64_t pot_type, uint64_t params_key,
                                         uint64_t *keys);
RGPOT_B200_API int rgpot_b200_cache_keys_device(RgpotB200Pot *pot, size_t nSystems, const size_t *offsets,
                                                const double *d_positions, const int *d_atomicNrs,
                                                const double *d_boxes, uint64_t pot_type,
                                                uint64_t params_key, uint64_t *d_keys, void *cuda_stream);

/** XXH3_64bits (xxHash 0.8.3, seed 0) of a host buffer, computed on the host by the same source the
 *  kernel compiles for its short inputs (hash.cuh): the engine uses it for the two words every key
 *  of a handle shares; hosts can use it to key single systems without a launch.  No GPU needed. */
RGPOT_B200_API uint64_t rgpot_b200_hash64(const void *data, size_t len);

/** Message of the last failure on this handle (rgpot_ferror.f90:41-56 semantics, but per
 *  handle): copies at most buffer_len - 1 bytes, NUL-terminates, returns the length. */
RGPOT_B200_API int rgpot_b200_last_error(RgpotB200Pot *pot, char *buffer, int buffer_len);

/** Kernels this handle has launched since creation (evidence for bench.py's gpu_launches). */
RGPOT_B200_API long rgpot_b200_launch_count(RgpotB200Pot *pot);
/** Diagnostics of the last finished single-system call: "launches" (as above), "slow_atoms" (atoms a
 *  brick kernel evaluated on its in-kernel slow path; 0 once the capacities follow the recorded halo
 *  populations), "max_halo" / "brick" / "cap_cand" / "cap_list" (largest halo population met, brick
 *  shape as decimal xyz, staged and list capacities of the plan; max_halo > 0 means the brick
 *  kernels did the work), "pool_used" (survivor-list entries the EAM density pass handed to the
 *  force pass).  -1: unknown key. */
RGPOT_B200_API long rgpot_b200_stat(RgpotB200Pot *pot, const char *key);
/** Tuning / test hooks: force a code path for subsequent calls.
 *  key "path": 0 auto, 1 cell-list, 2 all-pairs; "fine": 1 = never use half-cutoff cells;
 *  "fine_min": atoms from which half-cutoff cells are used; "brick" (cells per brick, decimal
 *  xyz), "brick_tpa", "brick_threads", "brick_list", "brick_slack": brick-kernel tuning, 0 = automatic;
 *  "bricks": 0 = gather kernels only; "handoff": 0 = the EAM force pass repeats the candidate search instead of reusing the density
 *  pass's lists (default 1); "half": 0 = the batch kernel of the EAMs evaluates every pair from both ends (full lists) instead of
 *  once with a fixed-point hand-over to the partner (default 1); "one": 0 = single EAM systems of 64..512 atoms use the batch
 *  kernel (one block) instead of the many-block kernel; "tile": 1 = the tensor-core tile kernel for large LJ systems (default 0:
 *  the brick kernel is faster); "batch_chunk": systems per pipeline chunk of the host-buffer batch entries (0 = automatic).
 *  The handle-less Fortran drop-in entries below take their GPU from the environment variable RGPOT_B200_DEVICE (default 0). */
RGPOT_B200_API int rgpot_b200_set_option(RgpotB200Pot *pot, const char *key, long value);

/** Page-locked host memory for the batch entry points. */
RGPOT_B200_API void *rgpot_b200_host_alloc(size_t bytes);
RGPOT_B200_API void rgpot_b200_host_free(void *ptr);

/** Sustained FP64 FMA rate of the device (TFLOP/s): the measured roofline denominator. */
RGPOT_B200_API double rgpot_b200_measure_fp64_tflops(int device, int iters);

/** Aggregate host<->device copy rate (GB/s) of `n_devices` GPUs moving h2d_bytes in and d2h_bytes out
 *  EACH, both directions and all devices at once, from page-locked buffers placed next to each GPU:
 *  the ceiling of every end-to-end batch number on this box (best of `reps` passes). */
RGPOT_B200_API double rgpot_b200_measure_copy_gbs(const int *devices, int n_devices, size_t h2d_bytes,
                                                  size_t d2h_bytes, int reps);

/** Payload rate (GB/s) at which `threads` host threads move `bytes` from pageable into page-locked
 *  memory and back in 5 KB pieces with the engine's streaming copy: what bounds the pageable
 *  ForceBatch entry when the GPUs are not the limit (best of `reps`). */
RGPOT_B200_API double rgpot_b200_measure_host_copy_gbs(int threads, size_t bytes, int reps);

/* ---- exact drop-in for the reference's Fortran entry (FortranPots.cc:33-40) ------------ */
RGPOT_B200_API int rgpot_cuh2_force(int32_t natoms, const double *positions,
                                    const int32_t *atomic_numbers, const double *cell,
                                    double *forces, double *energy);
/** The reference's EAM-Al entry (FortranPots.cc:26-28, rgpot_eam_al_capi.f90:31-58). */
RGPOT_B200_API int rgpot_eam_al_force(int32_t natoms, const double *positions, const double *cell,
                                      double *forces, double *energy);
/** The reference's Fe-He entry (FortranPots.cc, rgpot_fehe_capi.f90:36-63). */
RGPOT_B200_API int rgpot_fehe_force(int32_t natoms, const double *positions, const int32_t *atomic_numbers,
                                    const double *cell, double *forces, double *energy);
RGPOT_B200_API int rgpot_fortran_last_error(char *buffer, int buffer_len);

#ifdef __cplusplus
}
#endif
#endif /* RGPOT_B200_H */
