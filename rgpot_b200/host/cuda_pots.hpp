// cuda_pots.hpp -- rgpot plugin classes backed by the B200 engine (include/rgpot_b200.h).
//
// Drop-in `Potential<Derived>` subclasses (CppCore/rgpot/Potential.hpp:138-343):
//   rgpot::LJCudaPot         <->  rgpot::LJPot          (LennardJones/LJPot.hpp:40-86)
//   rgpot::LJClusterCudaPot  <->  rgpot::LJClusterPot   (LennardJones/LJClusterPot.hpp:47-100)
//   rgpot::CuH2CudaPot       <->  fortranpots::CuH2Pot  (fortran/FortranPots.hpp:29-45)
// Same PotType (so cache keys and potserv's selector strings stay valid), same config aggregates,
// same exception text on failure (FortranPots.cc:44-56); forceBatchImpl is native and
// caps().batched is true.  Build against the real rgpot headers with
// -DRGPOT_B200_USE_REFERENCE_HEADERS -I<rgpot>/CppCore, or standalone against rgpot_iface.hpp.
#pragma once

#ifdef RGPOT_B200_USE_REFERENCE_HEADERS
#include "rgpot/ForceStructs.hpp"
#include "rgpot/Potential.hpp"
#include "rgpot/pot_caps.hpp"
#include "rgpot/pot_types.hpp"
#else
#include "rgpot_iface.hpp"
#endif

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/rgpot_b200.h"

namespace rgpot {

/// LJConfig-compatible aggregate (LJPot.hpp:30-34) plus the device ordinal.
struct LJCudaConfig {
  double u0 = 1.0;
  double cutoff = 15.0;
  double psi = 1.0;
  int device = -1;
  bool all_images = false; ///< opt-in triclinic LJ (no reference counterpart)
  /// Batches on several GPUs of the box: with more than one ordinal here (or device = -2 for "all"),
  /// ONE forceBatch call cuts the batch across them (RgpotB200Config.devices); single systems use the first.
  std::vector<int> devices;
  int host_threads = 0; ///< helper threads per device for packing ForceBatch buffers (0 = automatic)
};

/// MorseConfig-compatible aggregate (Morse/MorsePot.hpp:31-36) plus the device ordinal.
struct MorseCudaConfig {
  double De = 0.7102;
  double a = 1.6047;
  double re = 2.8970;
  double cutoff = 9.5;
  int device = -1;
};

/// ZBLConfig-compatible aggregate (ZBL/ZBLPot.hpp:29-32) plus the device ordinal.
struct ZBLCudaConfig {
  double cut_inner = 2.0;
  double cut_global = 2.5;
  int device = -1;
};

namespace b200detail {

/// Owns one engine handle; shared by the three classes below.
class Engine {
public:
  Engine(int kind, const LJCudaConfig &c);
  explicit Engine(const MorseCudaConfig &c);
  explicit Engine(const ZBLCudaConfig &c);
  ~Engine();
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;
  void force(const char *label, const ForceInput &in, ForceOut *out) const;
  void forceBatch(const char *label, const ForceBatch &batch) const;
  /// Result-cache keys of a batch, one launch: keys[i] equals
  /// Potential<D>::cacheKey(batch.in[i]) (Potential.hpp:324-336, XXH3 over geometry, species, cell,
  /// type and paramsKey), so forceBatch's hit / miss compaction (:251-306) can be fed from the GPU.
  [[nodiscard]] std::vector<uint64_t> cacheKeys(const char *label, const ForceBatch &batch, PotType type) const;
  [[nodiscard]] uint64_t paramsKey() const noexcept { return key_; }

private:
  [[noreturn]] void raise(const char *label, int status) const;
  RgpotB200Pot *h_ = nullptr;
  uint64_t key_ = 0;
};

} // namespace b200detail

#define RGPOT_B200_POT_CLASS(ClassName, PotTypeValue, Kind, Label, Periodic)                       \
  class ClassName : public Potential<ClassName> {                                                 \
  public:                                                                                          \
    ClassName() : ClassName(LJCudaConfig{}) {}                                                     \
    explicit ClassName(const LJCudaConfig &c)                                                      \
        : Potential(PotType::PotTypeValue), m_config(c),                                           \
          m_engine(std::make_shared<b200detail::Engine>(Kind, c)) {}                               \
    void forceImpl(const ForceInput &in, ForceOut *out) const override {                           \
      m_engine->force(Label, in, out);                                                             \
    }                                                                                              \
    void forceBatchImpl(const ForceBatch &batch) const override {                                  \
      m_engine->forceBatch(Label, batch);                                                          \
    }                                                                                              \
    [[nodiscard]] PotCaps caps() const noexcept override {                                         \
      PotCaps c;                                                                                   \
      c.reentrancy = Reentrancy::PerInstance; /* one stream + scratch per handle */                \
      c.batched = true;                                                                            \
      c.periodic = Periodic;                                                                       \
      return c;                                                                                    \
    }                                                                                              \
    [[nodiscard]] uint64_t paramsKey() const noexcept override { return m_engine->paramsKey(); }   \
    [[nodiscard]] std::vector<uint64_t> cacheKeys(const ForceBatch &batch) const {                 \
      return m_engine->cacheKeys(Label, batch, get_type());                                        \
    }                                                                                              \
    [[nodiscard]] const LJCudaConfig &config() const noexcept { return m_config; }                 \
                                                                                                   \
  private:                                                                                         \
    LJCudaConfig m_config;                                                                         \
    std::shared_ptr<b200detail::Engine> m_engine;                                                  \
  }

RGPOT_B200_POT_CLASS(LJCudaPot, LJ, RGPOT_B200_LJ, "LJ", true);
RGPOT_B200_POT_CLASS(LJClusterCudaPot, LJCluster, RGPOT_B200_LJ_CLUSTER, "LJCluster", false);
RGPOT_B200_POT_CLASS(CuH2CudaPot, CuH2, RGPOT_B200_CUH2, "CuH2", true);
/// fortranpots::EAMAlPot on the engine (fortran/FortranPots.hpp:43): atomic numbers are not read.
RGPOT_B200_POT_CLASS(EAMAlCudaPot, EAMAl, RGPOT_B200_EAM_AL, "EAM-Al", true);
/// fortranpots::FeHePot on the engine (fortran/FortranPots.hpp:44): atomic numbers 26 / 2.
RGPOT_B200_POT_CLASS(FeHeCudaPot, FeHe, RGPOT_B200_FEHE, "FeHe", true);

#undef RGPOT_B200_POT_CLASS

/// rgpot::MorsePot on the engine (Morse/MorsePot.hpp:44-101): same PotType, same energyShift().
class MorseCudaPot : public Potential<MorseCudaPot> {
public:
  MorseCudaPot() : MorseCudaPot(MorseCudaConfig{}) {}
  explicit MorseCudaPot(const MorseCudaConfig &c)
      : Potential(PotType::Morse), m_config(c), m_engine(std::make_shared<b200detail::Engine>(c)) {}
  void forceImpl(const ForceInput &in, ForceOut *out) const override { m_engine->force("Morse", in, out); }
  void forceBatchImpl(const ForceBatch &batch) const override { m_engine->forceBatch("Morse", batch); }
  [[nodiscard]] PotCaps caps() const noexcept override {
    PotCaps c;
    c.reentrancy = Reentrancy::PerInstance;
    c.batched = true;
    return c;
  }
  [[nodiscard]] uint64_t paramsKey() const noexcept override { return m_engine->paramsKey(); }
  [[nodiscard]] std::vector<uint64_t> cacheKeys(const ForceBatch &batch) const {
    return m_engine->cacheKeys("Morse", batch, get_type());
  }
  [[nodiscard]] const MorseCudaConfig &config() const noexcept { return m_config; }
  [[nodiscard]] double energyShift() const noexcept;

private:
  MorseCudaConfig m_config;
  std::shared_ptr<b200detail::Engine> m_engine;
};

/// rgpot::ZBLPot on the engine (ZBL/ZBLPot.hpp:60-130): free boundaries, species-pair tables.
class ZBLCudaPot : public Potential<ZBLCudaPot> {
public:
  ZBLCudaPot() : ZBLCudaPot(ZBLCudaConfig{}) {}
  /// @throws std::invalid_argument unless 0 < cut_inner < cut_global (ZBLPot.cc:76-79)
  explicit ZBLCudaPot(const ZBLCudaConfig &c)
      : Potential(PotType::ZBL), m_config(c), m_engine(std::make_shared<b200detail::Engine>(c)) {}
  void forceImpl(const ForceInput &in, ForceOut *out) const override { m_engine->force("ZBL", in, out); }
  void forceBatchImpl(const ForceBatch &batch) const override { m_engine->forceBatch("ZBL", batch); }
  [[nodiscard]] PotCaps caps() const noexcept override {
    PotCaps c;
    c.reentrancy = Reentrancy::PerInstance;
    c.batched = true;
    c.periodic = false;
    return c;
  }
  [[nodiscard]] uint64_t paramsKey() const noexcept override { return m_engine->paramsKey(); }
  [[nodiscard]] std::vector<uint64_t> cacheKeys(const ForceBatch &batch) const {
    return m_engine->cacheKeys("ZBL", batch, get_type());
  }
  [[nodiscard]] const ZBLCudaConfig &config() const noexcept { return m_config; }

private:
  ZBLCudaConfig m_config;
  std::shared_ptr<b200detail::Engine> m_engine;
};

} // namespace rgpot
