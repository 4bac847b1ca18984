// cuda_pots.cc -- host glue between rgpot's plugin interface and the engine's C ABI.
#include "cuda_pots.hpp"

#include <array>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace rgpot {
namespace b200detail {

namespace {
// FNV-1a over the config fields plus a kernel-version salt, the scheme of
// CppCore/rgpot/ParamHash.hpp:19-51 / LJPot.hpp:54-59.  The salt differs from the CPU kernels'
// (1) so a shared result cache never mixes CPU and GPU results.
constexpr uint64_t kKernelVersion = 0xB2000001ull;
struct Fnv {
  uint64_t h = 0xcbf29ce484222325ull;
  void bytes(const void *p, size_t n) {
    const auto *b = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n; ++i) {
      h ^= b[i];
      h *= 0x100000001b3ull;
    }
  }
  void u64(uint64_t v) { bytes(&v, sizeof v); }
  void f64(double v) {
    uint64_t b;
    std::memcpy(&b, &v, sizeof b);
    bytes(&b, sizeof b);
  }
};
} // namespace

Engine::Engine(int kind, const LJCudaConfig &c) {
  RgpotB200Config cfg;
  rgpot_b200_default_config(kind, &cfg);
  cfg.device = c.device;
  cfg.host_threads = c.host_threads;
  if (c.devices.size() > 16) throw std::invalid_argument("rgpot_b200: at most 16 devices");
  if (!c.devices.empty()) {
    cfg.device = c.devices[0];
    cfg.n_devices = static_cast<int>(c.devices.size());
    for (size_t k = 0; k < c.devices.size(); ++k) cfg.devices[k] = c.devices[k];
  }
  Fnv fp;
  fp.u64(kKernelVersion);
  if (kind == RGPOT_B200_EAM_AL || kind == RGPOT_B200_FEHE) {
    fp.u64((uint64_t)kind);  // parameter-free like CuH2; the kind keeps the two keys apart
  } else if (kind != RGPOT_B200_CUH2) {
    cfg.u0 = c.u0;
    cfg.cutoff = c.cutoff;
    cfg.psi = c.psi;
    cfg.lj_all_images = c.all_images ? 1 : 0;
    fp.f64(c.u0);
    fp.f64(c.cutoff);
    fp.f64(c.psi);
  }
  key_ = fp.h;
  char err[512] = {0};
  h_ = rgpot_b200_create(&cfg, err, sizeof err);
  if (!h_) throw std::runtime_error(std::string("rgpot_b200 engine: ") + err);
}

Engine::Engine(const MorseCudaConfig &c) {
  RgpotB200Config cfg;
  rgpot_b200_default_config(RGPOT_B200_MORSE, &cfg);
  cfg.device = c.device;
  cfg.morse_De = c.De;
  cfg.morse_a = c.a;
  cfg.morse_re = c.re;
  cfg.cutoff = c.cutoff;
  Fnv fp;  // MorsePot.hpp:62-68
  fp.u64(kKernelVersion);
  fp.f64(c.De);
  fp.f64(c.a);
  fp.f64(c.re);
  fp.f64(c.cutoff);
  key_ = fp.h;
  char err[512] = {0};
  h_ = rgpot_b200_create(&cfg, err, sizeof err);
  if (!h_) throw std::runtime_error(std::string("rgpot_b200 engine: ") + err);
}

Engine::Engine(const ZBLCudaConfig &c) {
  if (!(c.cut_inner > 0.0) || !(c.cut_inner < c.cut_global))
    throw std::invalid_argument("ZBLPot: invalid cutoffs, require 0.0 < cut_inner < cut_global.");
  RgpotB200Config cfg;
  rgpot_b200_default_config(RGPOT_B200_ZBL, &cfg);
  cfg.device = c.device;
  cfg.zbl_cut_inner = c.cut_inner;
  cfg.cutoff = c.cut_global;
  Fnv fp;  // ZBLPot.cc:80-84
  fp.u64(kKernelVersion);
  fp.f64(c.cut_inner);
  fp.f64(c.cut_global);
  key_ = fp.h;
  char err[512] = {0};
  h_ = rgpot_b200_create(&cfg, err, sizeof err);
  if (!h_) throw std::runtime_error(std::string("rgpot_b200 engine: ") + err);
}

Engine::~Engine() { rgpot_b200_destroy(h_); }

// FortranPots.cc:44-56: "<pot> potential failed (status k): <message>"
void Engine::raise(const char *label, int status) const {
  std::array<char, 512> buffer{};
  const int written = rgpot_b200_last_error(h_, buffer.data(), static_cast<int>(buffer.size()));
  std::string message = std::string(label) + " potential failed (status " + std::to_string(status) + ")";
  if (written > 0) {
    message += ": ";
    message.append(buffer.data(), static_cast<size_t>(written));
  }
  throw std::runtime_error(message);
}

void Engine::force(const char *label, const ForceInput &in, ForceOut *out) const {
  const int status = rgpot_b200_force(h_, static_cast<long>(in.nAtoms), in.pos, in.atmnrs, out->F,
                                      &out->energy, &out->variance, in.box);
  if (status != 0) raise(label, status);
  out->variance = 0.0;
}

void Engine::forceBatch(const char *label, const ForceBatch &batch) const {
  const size_t n = batch.nSystems;
  if (n == 0) return;
  // per-thread scratch, kept between calls: a 65 536-image batch would otherwise allocate (and fault in) 3 MB of
  // pointer tables on every call
  thread_local std::vector<size_t> natoms;
  thread_local std::vector<const double *> pos, box;
  thread_local std::vector<const int *> z;
  thread_local std::vector<double *> frc;
  thread_local std::vector<double> energy;
  natoms.resize(n);
  pos.resize(n);
  box.resize(n);
  z.resize(n);
  frc.resize(n);
  energy.resize(n);
  for (size_t i = 0; i < n; ++i) {
    natoms[i] = batch.in[i].nAtoms;
    pos[i] = batch.in[i].pos;
    z[i] = batch.in[i].atmnrs;
    box[i] = batch.in[i].box;
    frc[i] = batch.out[i].F;
  }
  const int status = rgpot_b200_force_batch(h_, n, natoms.data(), pos.data(), z.data(), box.data(),
                                            frc.data(), energy.data(), nullptr);
  for (size_t i = 0; i < n; ++i) {
    batch.out[i].energy = energy[i];
    batch.out[i].variance = 0.0;
  }
  if (status != 0) raise(label, status);
}

std::vector<uint64_t> Engine::cacheKeys(const char *label, const ForceBatch &batch, PotType type) const {
  const size_t n = batch.nSystems;
  std::vector<uint64_t> keys(n);
  if (n == 0) return keys;
  // pack the systems back to back (the engine hashes one contiguous batch in one launch)
  std::vector<size_t> offsets(n + 1, 0);
  for (size_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + batch.in[i].nAtoms;
  std::vector<double> pos(3 * offsets[n] + 1), boxes(9 * n);
  std::vector<int> z(offsets[n] + 1);
  for (size_t i = 0; i < n; ++i) {
    std::copy(batch.in[i].pos, batch.in[i].pos + 3 * batch.in[i].nAtoms, pos.begin() + 3 * offsets[i]);
    std::copy(batch.in[i].atmnrs, batch.in[i].atmnrs + batch.in[i].nAtoms, z.begin() + offsets[i]);
    std::copy(batch.in[i].box, batch.in[i].box + 9, boxes.begin() + 9 * i);
  }
  const int status = rgpot_b200_cache_keys(h_, n, offsets.data(), pos.data(), z.data(), boxes.data(),
                                           static_cast<uint64_t>(static_cast<size_t>(type)), key_, keys.data());
  if (status != 0) raise(label, status);
  return keys;
}

} // namespace b200detail

double MorseCudaPot::energyShift() const noexcept {
  const double d = 1.0 - std::exp(-m_config.a * (m_config.cutoff - m_config.re));
  return m_config.De * d * d - m_config.De;
}

} // namespace rgpot
