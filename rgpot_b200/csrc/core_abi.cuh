// core_abi.cuh -- rgpot_b200_core_callback: the engine behind the rgpot-core PotentialCallback
// (include/rgpot_b200_core.h; reference: rgpot-core/include/rgpot.h:76-131, the C++ trampoline
// include/rgpot/potential.hpp:202-250 and the owned tensors of rgpot-core/src/tensor.rs:150-204).
// Included at the end of engine.cu (it uses the engine's entry points directly).
#pragma once

#include <new>

#include "../../include/rgpot_b200_core.h"

namespace {

thread_local std::string g_core_error;

struct OwnedForces {  // manager_ctx of the forces tensor handed to the caller
  int64_t shape[2];
  int64_t strides[2];
  void* data;
  bool on_device;
  int device;
};

void owned_forces_deleter(DLManagedTensorVersioned* self) {
  if (!self) return;
  auto* ctx = static_cast<OwnedForces*>(self->manager_ctx);
  if (ctx) {
    if (ctx->on_device) {
      int prev = 0;
      cudaGetDevice(&prev);
      cudaSetDevice(ctx->device);
      cudaFree(ctx->data);
      cudaSetDevice(prev);
    } else {
      free(ctx->data);
    }
    delete ctx;
  }
  delete self;
}

DLManagedTensorVersioned* make_forces_tensor(long n, bool on_device, int device) {
  auto* ctx = new (std::nothrow) OwnedForces{};
  auto* t = new (std::nothrow) DLManagedTensorVersioned{};
  if (!ctx || !t) {
    delete ctx;
    delete t;
    return nullptr;
  }
  const size_t bytes = sizeof(double) * 3 * (size_t)std::max<long>(n, 1);
  if (on_device) {
    if (cudaMalloc(&ctx->data, bytes) != cudaSuccess) ctx->data = nullptr;
  } else {
    ctx->data = calloc(1, bytes);
  }
  if (!ctx->data) {
    delete ctx;
    delete t;
    return nullptr;
  }
  ctx->shape[0] = n;
  ctx->shape[1] = 3;
  ctx->strides[0] = 3;
  ctx->strides[1] = 1;
  ctx->on_device = on_device;
  ctx->device = device;
  t->version.major = 1;
  t->version.minor = 0;
  t->manager_ctx = ctx;
  t->deleter = owned_forces_deleter;
  t->flags = DLPACK_FLAG_BITMASK_IS_COPIED;
  t->dl_tensor.data = ctx->data;
  t->dl_tensor.device.device_type = on_device ? kDLCUDA : kDLCPU;
  t->dl_tensor.device.device_id = on_device ? device : 0;
  t->dl_tensor.ndim = 2;
  t->dl_tensor.dtype.code = kDLFloat;
  t->dl_tensor.dtype.bits = 64;
  t->dl_tensor.dtype.lanes = 1;
  t->dl_tensor.shape = ctx->shape;
  t->dl_tensor.strides = ctx->strides;
  t->dl_tensor.byte_offset = 0;
  return t;
}

// compact row-major tensor of the given rank / dtype?  (strides may be NULL)
bool tensor_ok(const DLTensor& t, int ndim, uint8_t code, uint8_t bits) {
  if (t.ndim != ndim || t.dtype.code != code || t.dtype.bits != bits || t.dtype.lanes != 1) return false;
  if (!t.shape) return false;
  if (t.strides) {
    int64_t expect = 1;
    for (int k = ndim - 1; k >= 0; --k) {
      if (t.shape[k] > 1 && t.strides[k] != expect) return false;
      expect *= t.shape[k];
    }
  }
  return true;
}

const void* tensor_data(const DLTensor& t) { return static_cast<const char*>(t.data) + t.byte_offset; }

}  // namespace

extern "C" {

const char* rgpot_b200_core_last_error(void) { return g_core_error.c_str(); }

void rgpot_b200_core_free(void* user_data) { rgpot_b200_destroy(static_cast<RgpotB200Pot*>(user_data)); }

enum rgpot_status_t rgpot_b200_core_callback(void* user_data, const struct rgpot_force_input_t* input,
                                             struct rgpot_force_out_t* output) {
  g_core_error.clear();
  auto bad = [&](const char* msg) {
    g_core_error = msg;
    return RGPOT_INVALID_PARAMETER;
  };
  try {
    auto* pot = static_cast<RgpotB200Pot*>(user_data);
    if (!pot || !input || !output) return bad("rgpot_b200_core_callback: NULL handle, input or output");
    output->forces = nullptr;
    output->energy = 0.0;
    output->variance = 0.0;
    if (!input->positions) return bad("rgpot_b200_core_callback: positions tensor is NULL");
    const DLTensor& P = input->positions->dl_tensor;
    if (!tensor_ok(P, 2, kDLFloat, 64) || P.shape[1] != 3 || P.shape[0] < 0)
      return bad("rgpot_b200_core_callback: positions must be a compact [n_atoms, 3] float64 tensor");
    const long n = (long)P.shape[0];
    const bool on_device = P.device.device_type == kDLCUDA;
    if (!on_device && P.device.device_type != kDLCPU && P.device.device_type != kDLCUDAHost)
      return bad("rgpot_b200_core_callback: positions must live on kDLCPU or kDLCUDA");
    if (on_device && P.device.device_id != pot->device)
      return bad("rgpot_b200_core_callback: CUDA tensors must live on the engine's device");
    const int* Z = nullptr;
    if (input->atomic_numbers) {
      const DLTensor& A = input->atomic_numbers->dl_tensor;
      if (!tensor_ok(A, 1, kDLInt, 32) || A.shape[0] != n)
        return bad("rgpot_b200_core_callback: atomic_numbers must be a compact [n_atoms] int32 tensor");
      if ((A.device.device_type == kDLCUDA) != on_device)
        return bad("rgpot_b200_core_callback: positions and atomic_numbers live on different devices");
      Z = static_cast<const int*>(tensor_data(A));
    }
    double box[9];
    const double* boxp = nullptr;
    if (input->box_matrix) {
      const DLTensor& B = input->box_matrix->dl_tensor;
      if (!tensor_ok(B, 2, kDLFloat, 64) || B.shape[0] != 3 || B.shape[1] != 3)
        return bad("rgpot_b200_core_callback: box_matrix must be a compact [3, 3] float64 tensor");
      if (B.device.device_type == kDLCUDA) {  // nine doubles: the grid is planned on the host
        if (cudaMemcpy(box, tensor_data(B), sizeof box, cudaMemcpyDeviceToHost) != cudaSuccess)
          return bad("rgpot_b200_core_callback: cannot read the box tensor");
      } else {
        memcpy(box, tensor_data(B), sizeof box);
      }
      boxp = box;
    }
    // periodic kinds need the cell (the host entry refuses a missing one too); a NULL box_matrix is
    // a caller error, not an engine failure
    if (!boxp && pot->cfg.kind != RGPOT_B200_LJ_CLUSTER && pot->cfg.kind != RGPOT_B200_ZBL)
      return bad("rgpot_b200_core_callback: box_matrix is NULL, but this potential is periodic");
    // the forces tensor and the scratch below live on the handle's device, whatever device the
    // calling thread had current; the caller's choice is put back on exit
    DevGuard guard(pot->device);
    if (guard.err != cudaSuccess) {
      g_core_error = "rgpot_b200_core_callback: cannot select the engine's device";
      return RGPOT_INTERNAL_ERROR;
    }
    DLManagedTensorVersioned* F = make_forces_tensor(n, on_device, pot->device);
    if (!F) {
      g_core_error = "rgpot_b200_core_callback: cannot allocate the forces tensor";
      return RGPOT_INTERNAL_ERROR;
    }
    int st = 0;
    double energy = 0.0, variance = 0.0;
    if (on_device) {
      double* d_E = nullptr;
      if (cudaMalloc(&d_E, sizeof(double)) != cudaSuccess) {
        owned_forces_deleter(F);
        g_core_error = "rgpot_b200_core_callback: cannot allocate device scratch";
        return RGPOT_INTERNAL_ERROR;
      }
      st = rgpot_b200_force_device(pot, n, static_cast<const double*>(tensor_data(P)), Z,
                                   static_cast<double*>(F->dl_tensor.data), d_E, boxp, nullptr);
      if (st == 0) st = rgpot_b200_sync_status(pot);
      if (st == 0 && cudaMemcpy(&energy, d_E, sizeof energy, cudaMemcpyDeviceToHost) != cudaSuccess) st = 100;
      cudaFree(d_E);
    } else {
      st = rgpot_b200_force(pot, n, static_cast<const double*>(tensor_data(P)), Z,
                            static_cast<double*>(F->dl_tensor.data), &energy, &variance, boxp);
    }
    if (st != 0) {
      char msg[512];
      rgpot_b200_last_error(pot, msg, (int)sizeof msg);
      g_core_error = std::string("rgpot_b200 engine failed (status ") + std::to_string(st) + "): " + msg;
      owned_forces_deleter(F);
      return RGPOT_INTERNAL_ERROR;
    }
    output->forces = F;
    output->energy = energy;
    output->variance = variance;
    return RGPOT_SUCCESS;
  } catch (const std::exception& e) {
    g_core_error = std::string("rgpot_b200_core_callback: ") + e.what();
    return RGPOT_INTERNAL_ERROR;
  } catch (...) {
    g_core_error = "rgpot_b200_core_callback: unknown exception";
    return RGPOT_INTERNAL_ERROR;
  }
}

}  // extern "C"
