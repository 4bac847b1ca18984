/**
 * @file rgpot_b200_core.h
 * @brief The B200 engine behind the rgpot-core C ABI (rgpot-core/include/rgpot.h): a ready-made
 *        `PotentialCallback` over DLPack tensors.
 *
 * The reference's language-neutral boundary is a callback
 *   rgpot_status_t (*PotentialCallback)(void *user_data, const rgpot_force_input_t *,
 *                                       rgpot_force_out_t *)              (rgpot.h:129-131)
 * registered with rgpot_potential_new(cb, user_data, free_fn) (rgpot.h:357-378); the Rust RPC
 * server (rgpot-core/src/rpc/server.rs:77), the eindir objective (eindir.rs:96-215) and
 * PotentialHandle::calculate (include/rgpot/potential.hpp:121-127) all call through it.
 * rgpot_b200_core_callback has exactly that signature with an engine handle as user_data:
 *
 *   RgpotB200Pot *h = rgpot_b200_create(&cfg, err, sizeof err);
 *   rgpot_potential_t *p = rgpot_potential_new(rgpot_b200_core_callback, h, rgpot_b200_core_free);
 *
 * Semantics (rgpot.h:76-117, rgpot-core/src/types.rs:8-26, tensor.rs:57-204):
 *   - inputs are BORROWED DLPack tensors: positions [n,3] f64, atomic_numbers [n] i32 (may be NULL
 *     for the LJ kinds), box_matrix [3,3] f64 (may be NULL), compact row-major;
 *   - output->forces is CALLEE-allocated: an owning DLManagedTensorVersioned (version 1.0,
 *     DLPACK_FLAG_BITMASK_IS_COPIED, row-major strides) whose deleter frees it, released by the
 *     caller through rgpot_tensor_free == tensor->deleter(tensor) (tensor.rs:393-401);
 *   - kDLCPU inputs give a kDLCPU forces tensor (what every in-tree consumer dereferences,
 *     potential.hpp:217, rpc/server.rs:80-86); kDLCUDA inputs ("any device", rgpot.h:78-88) on the
 *     handle's device are evaluated in place (no host copies) and give a kDLCUDA forces tensor;
 *   - no exception crosses the boundary: RGPOT_INVALID_PARAMETER for NULL / mis-shaped /
 *     mis-typed / strided tensors, RGPOT_INTERNAL_ERROR for an engine failure (the message is
 *     kept per thread, rgpot_b200_core_last_error, like rgpot_last_error rgpot.h:269).
 *
 * The DLPack and rgpot-core types are defined here only when the real headers have not been
 * included first (the reference does not vendor dlpack.h; rgpot.h:18 takes it from an install).
 */
#ifndef RGPOT_B200_CORE_H
#define RGPOT_B200_CORE_H

#include <stdint.h>

#include "rgpot_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DLPACK_DLPACK_H_  /* minimal DLPack 1.0 subset, layout-identical to dlpack/dlpack.h */
#define DLPACK_DLPACK_H_
#define DLPACK_MAJOR_VERSION 1
#define DLPACK_MINOR_VERSION 0
typedef struct {
  uint32_t major;
  uint32_t minor;
} DLPackVersion;
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLCUDAManaged = 13 } DLDeviceType;
typedef struct {
  DLDeviceType device_type;
  int32_t device_id;
} DLDevice;
typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U } DLDataTypeCode;
typedef struct {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} DLDataType;
typedef struct {
  void *data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t *shape;
  int64_t *strides; /* NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
#define DLPACK_FLAG_BITMASK_READ_ONLY (1UL << 0UL)
#define DLPACK_FLAG_BITMASK_IS_COPIED (1UL << 1UL)
typedef struct DLManagedTensorVersioned {
  DLPackVersion version;
  void *manager_ctx;
  void (*deleter)(struct DLManagedTensorVersioned *self);
  uint64_t flags;
  DLTensor dl_tensor;
} DLManagedTensorVersioned;
#endif /* DLPACK_DLPACK_H_ */

#ifndef RGPOT_H /* the three rgpot-core types this boundary uses (rgpot.h:28-49, 76-117, 129-131) */
typedef enum rgpot_status_t {
  RGPOT_SUCCESS = 0,
  RGPOT_INVALID_PARAMETER = 1,
  RGPOT_INTERNAL_ERROR = 2,
  RGPOT_RPC_ERROR = 3,
  RGPOT_BUFFER_SIZE_ERROR = 4,
} rgpot_status_t;
typedef struct rgpot_force_input_t {
  DLManagedTensorVersioned *positions;
  DLManagedTensorVersioned *atomic_numbers;
  DLManagedTensorVersioned *box_matrix;
} rgpot_force_input_t;
typedef struct rgpot_force_out_t {
  DLManagedTensorVersioned *forces;
  double energy;
  double variance;
} rgpot_force_out_t;
typedef enum rgpot_status_t (*PotentialCallback)(void *user_data, const struct rgpot_force_input_t *input,
                                                 struct rgpot_force_out_t *output);
#endif /* RGPOT_H */

/** A `PotentialCallback`; `user_data` is the RgpotB200Pot* of rgpot_b200_create. */
RGPOT_B200_API enum rgpot_status_t rgpot_b200_core_callback(void *user_data,
                                                            const struct rgpot_force_input_t *input,
                                                            struct rgpot_force_out_t *output);
/** The `free_fn` of rgpot_potential_new: destroys the engine handle. */
RGPOT_B200_API void rgpot_b200_core_free(void *user_data);
/** Message of the last failure of rgpot_b200_core_callback on this thread ("" if none); valid until
 *  the next call on the thread (rgpot_last_error semantics). */
RGPOT_B200_API const char *rgpot_b200_core_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* RGPOT_B200_CORE_H */
