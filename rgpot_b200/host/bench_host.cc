// bench_host.cc -- the C++ drop-in entries timed from C++ (bench.py runs this and reports the line):
//   Potential::forceBatch(const ForceBatch&)  (CppCore/rgpot/Potential.hpp:251-312) on nSystems
//       separately allocated, pageable systems -- what an eOn NEB client holds -- through
//       CuH2CudaPot::forceBatchImpl, on one device or on all listed devices in ONE call;
//   forceImpl on a single 218-atom system (the operator() shape, :165-211).
// usage: bench_host <slab218.bin> <configs> <steps> <n_devices> [host_threads]
//   slab218.bin: 218 x 3 doubles (positions), 218 int32 (atomic numbers), 9 doubles (box)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "cuda_pots.hpp"

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: bench_host <slab218.bin> <configs> <steps> <n_devices> [host_threads]\n");
    return 2;
  }
  const size_t n = 218;
  std::vector<double> base(3 * n), box(9);
  std::vector<int> Z(n);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(base.data(), 8, 3 * n, f) != 3 * n || fread(Z.data(), 4, n, f) != n || fread(box.data(), 8, 9, f) != 9) {
    fprintf(stderr, "bench_host: cannot read %s\n", argv[1]);
    return 2;
  }
  fclose(f);
  const size_t nsys = strtoul(argv[2], nullptr, 10);
  const int steps = atoi(argv[3]), ndev = atoi(argv[4]);
  rgpot::LJCudaConfig cfg;
  if (ndev > 1)
    for (int d = 0; d < ndev; ++d) cfg.devices.push_back(d);
  else
    cfg.device = 0;
  if (argc > 5) cfg.host_threads = atoi(argv[5]);
  rgpot::CuH2CudaPot pot(cfg);

  // every system owns its buffers: separate heap allocations, pageable
  std::vector<std::unique_ptr<double[]>> pos(nsys), frc(nsys), bx(nsys);
  std::vector<std::unique_ptr<int[]>> zz(nsys);
  std::vector<rgpot::ForceInput> in;  // (ForceInput has const members: built in place, never assigned)
  std::vector<rgpot::ForceOut> out;
  in.reserve(nsys);
  out.reserve(nsys);
  unsigned long long lcg = 88172645463325252ull;
  for (size_t s = 0; s < nsys; ++s) {
    pos[s].reset(new double[3 * n]);
    frc[s].reset(new double[3 * n]);
    bx[s].reset(new double[9]);
    zz[s].reset(new int[n]);
    for (size_t k = 0; k < 3 * n; ++k) {
      lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
      pos[s][k] = base[k] + 0.1 * ((double)(lcg >> 11) / 9007199254740992.0 - 0.5);
    }
    memcpy(zz[s].get(), Z.data(), 4 * n);
    memcpy(bx[s].get(), box.data(), 72);
    in.push_back(rgpot::ForceInput{n, pos[s].get(), zz[s].get(), bx[s].get()});
    out.push_back(rgpot::ForceOut{frc[s].get(), 0.0, 0.0});
  }
  rgpot::ForceBatch batch{nsys, in.data(), out.data()};
  pot.forceBatch(batch);  // warm-up
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < steps; ++k) pot.forceBatch(batch);
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps;
  double esum = 0.0;
  for (size_t s = 0; s < nsys; ++s) esum += out[s].energy;

  // one 218-atom system through forceImpl (what operator() calls)
  std::vector<double> F1(3 * n);
  rgpot::ForceInput one{n, base.data(), Z.data(), box.data()};
  rgpot::ForceOut o1{F1.data(), 0.0, 0.0};
  for (int k = 0; k < 100; ++k) pot.forceImpl(one, &o1);
  const auto t1 = std::chrono::steady_clock::now();
  const int reps = 2000;
  for (int k = 0; k < reps; ++k) pot.forceImpl(one, &o1);
  const double us = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count() / reps * 1e6;
  printf("{\"what\": \"cxx_plugin\", \"api\": \"rgpot::CuH2CudaPot::forceBatch(const ForceBatch&), %zu separately allocated pageable systems\", "
         "\"configs\": %zu, \"devices\": %d, \"ms_per_batch\": %.4f, \"atom_evals_per_s\": %.6e, \"energy_sum\": %.10f, "
         "\"single_218_forceImpl_us\": %.2f, \"single_218_energy\": %.12f}\n",
         nsys, nsys, ndev, dt * 1e3, (double)(nsys * n) / dt, esum, us, o1.energy);
  return 0;
}
