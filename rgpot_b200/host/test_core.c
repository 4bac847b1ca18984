/* test_core.c -- the rgpot-core boundary (include/rgpot_b200_core.h) from plain C: what
 * rgpot_potential_new(rgpot_b200_core_callback, handle, rgpot_b200_core_free) +
 * rgpot_potential_calculate (rgpot-core/include/rgpot.h:357-378) do, without the Rust crate.
 * Reproduces CppCore/tests/CuH2PotTest.cc:11-31 through DLPack tensors. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rgpot_b200_core.h"

static DLManagedTensorVersioned borrow(void *data, int ndim, int64_t *shape, uint8_t code, uint8_t bits) {
  DLManagedTensorVersioned t;
  memset(&t, 0, sizeof t);
  t.version.major = 1;
  t.dl_tensor.data = data;
  t.dl_tensor.device.device_type = kDLCPU;
  t.dl_tensor.ndim = ndim;
  t.dl_tensor.dtype.code = code;
  t.dl_tensor.dtype.bits = bits;
  t.dl_tensor.dtype.lanes = 1;
  t.dl_tensor.shape = shape;
  return t;
}

int main(void) {
  double pos[12] = {0.63940268750835, 0.90484742551374, 6.97516498544584,  3.19652040936288, 0.90417430354811,
                    6.97547796369474, 8.98363230369760, 9.94703496017833,  7.83556854923689, 7.64080177576300,
                    9.94703114803832, 7.83556986121272};
  int Z[4] = {29, 29, 1, 1};
  double box[9] = {15.345599999999999, 0, 0, 0, 21.702000000000002, 0, 0, 0, 100.00000000000000};
  int64_t sp[2] = {4, 3}, sz[1] = {4}, sb[2] = {3, 3};
  DLManagedTensorVersioned tp = borrow(pos, 2, sp, kDLFloat, 64), tz = borrow(Z, 1, sz, kDLInt, 32),
                           tb = borrow(box, 2, sb, kDLFloat, 64);
  RgpotB200Config cfg;
  char err[512];
  rgpot_b200_default_config(RGPOT_B200_CUH2, &cfg);
  RgpotB200Pot *h = rgpot_b200_create(&cfg, err, sizeof err);
  if (!h) {
    printf("create failed: %s\n", err);
    return 2;
  }
  PotentialCallback cb = rgpot_b200_core_callback; /* the reference's callback type, unchanged */
  rgpot_force_input_t in = {&tp, &tz, &tb};
  rgpot_force_out_t out = {0, 0.0, 0.0};
  rgpot_status_t st = cb(h, &in, &out);
  if (st != RGPOT_SUCCESS) {
    printf("callback failed (%d): %s\n", (int)st, rgpot_b200_core_last_error());
    return 1;
  }
  const double *F = (const double *)out.forces->dl_tensor.data;
  int ok = fabs(out.energy - (-2.711409624266224)) < 1e-9 * 2.72 && out.variance == 0.0 &&
           out.forces->dl_tensor.ndim == 2 && out.forces->dl_tensor.shape[0] == 4 &&
           out.forces->dl_tensor.shape[1] == 3 && out.forces->version.major == 1 &&
           (out.forces->flags & DLPACK_FLAG_BITMASK_IS_COPIED);
  ok = ok && fabs(F[0] - 1.4919411183978113) < 1e-8; /* CuH2PotTest.cc:27 */
  out.forces->deleter(out.forces);                    /* rgpot_tensor_free */
  /* a malformed input never reaches the engine */
  sp[1] = 2;
  ok = ok && cb(h, &in, &out) == RGPOT_INVALID_PARAMETER && out.forces == 0;
  rgpot_b200_core_free(h);
  printf(ok ? "test_core: ok\n" : "test_core: FAILED\n");
  return ok ? 0 : 1;
}
