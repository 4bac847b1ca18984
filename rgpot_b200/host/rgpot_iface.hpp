// rgpot_iface.hpp -- stand-in for rgpot's plugin interface, used ONLY when the reference headers
// are not on the include path (e.g. on the GPU box, where /root/reference does not exist).
//
// Same names, argument meaning and behaviour as the subset of
//   CppCore/rgpot/ForceStructs.hpp:20-54, pot_caps.hpp:20-45, pot_types.hpp:26-45,
//   types/AtomMatrix.hpp:30-160, PotHelpers.hpp:27-89, Potential.hpp:40-312
// that the CUDA potentials touch, written from the interface description, not copied.
// With `-I<rgpot>/CppCore` and -DRGPOT_B200_USE_REFERENCE_HEADERS the real headers are used
// instead and this file is not read at all (see cuda_pots.hpp).
#pragma once

#include <array>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <tuple>
#include <vector>

namespace rgpot {

struct ForceInput {
  const size_t nAtoms;
  const double *pos;
  const int *atmnrs;
  const double *box;
};
struct ForceOut {
  double *F;
  double energy;
  double variance;
};
struct ForceBatch {
  size_t nSystems;
  const ForceInput *in;
  ForceOut *out;
};

enum class Reentrancy : uint8_t { SharedInstance, PerInstance, ProcessSerial };
struct PotCaps {
  Reentrancy reentrancy = Reentrancy::SharedInstance;
  bool perImageInstances = false;
  bool batched = false;
  bool periodic = true;
};

enum class PotType {
  UNKNOWN = 0, CuH2, LJ, XTB, TBLite, Metatomic, NWChem, CPMD, Morse, LJCluster, ZBL, SWSi, EDIP,
  LenoskySi, TersoffSi, EAMAl, FeHe, WaterH
};

namespace types {
class AtomMatrix {
public:
  AtomMatrix() = default;
  AtomMatrix(size_t r, size_t c) : r_(r), c_(c), d_(r * c) {}
  AtomMatrix(std::initializer_list<std::initializer_list<double>> rows)
      : r_(rows.size()), c_(rows.size() ? rows.begin()->size() : 0) {
    d_.reserve(r_ * c_);
    for (const auto &row : rows) d_.insert(d_.end(), row.begin(), row.end());
  }
  static AtomMatrix Zero(size_t r, size_t c) { return AtomMatrix(r, c); }
  double &operator()(size_t i, size_t j) { return d_[i * c_ + j]; }
  const double &operator()(size_t i, size_t j) const { return d_[i * c_ + j]; }
  size_t rows() const { return r_; }
  size_t cols() const { return c_; }
  size_t size() const { return d_.size(); }
  double *data() { return d_.data(); }
  const double *data() const { return d_.data(); }

private:
  size_t r_ = 0, c_ = 0;
  std::vector<double> d_;
};
} // namespace types

template <typename T> struct registry {
  static std::atomic<size_t> count;
  static std::atomic<size_t> forceCalls;
  registry() { ++count; }
  registry(const registry &) { ++count; }
  ~registry() { --count; }
  static void incrementForceCalls() { forceCalls.fetch_add(1, std::memory_order_relaxed); }
};
template <typename T> std::atomic<size_t> registry<T>::count{0};
template <typename T> std::atomic<size_t> registry<T>::forceCalls{0};

class PotentialBase {
public:
  explicit PotentialBase(PotType t) : m_type(t) {}
  virtual ~PotentialBase() = default;
  virtual std::tuple<double, types::AtomMatrix, double>
  operator()(const types::AtomMatrix &positions, const std::vector<int> &atmtypes,
             const std::array<std::array<double, 3>, 3> &box) = 0;
  [[nodiscard]] virtual PotCaps caps() const noexcept { return {}; }
  virtual void forceBatch(const ForceBatch &) {
    throw std::runtime_error("PotentialBase::forceBatch called directly");
  }
  [[nodiscard]] virtual uint64_t paramsKey() const noexcept { return 0; }
  [[nodiscard]] PotType get_type() const { return m_type; }

protected:
  PotType m_type;
};

template <typename Derived> class Potential : public PotentialBase, public registry<Derived> {
public:
  using PotentialBase::PotentialBase;
  std::tuple<double, types::AtomMatrix, double>
  operator()(const types::AtomMatrix &positions, const std::vector<int> &atmtypes,
             const std::array<std::array<double, 3>, 3> &box) override {
    const size_t n = positions.rows();
    types::AtomMatrix forces = types::AtomMatrix::Zero(n, 3);
    double flat[9];
    std::memcpy(flat, static_cast<const void *>(&box), sizeof flat);
    ForceInput fi{n, positions.data(), atmtypes.data(), flat};
    ForceOut fo{forces.data(), 0.0, 0.0};
    static_cast<Derived *>(this)->forceImpl(fi, &fo);
    registry<Derived>::incrementForceCalls();
    return {fo.energy, std::move(forces), fo.variance};
  }
  virtual void forceImpl(const ForceInput &in, ForceOut *out) const = 0;
  virtual void forceBatchImpl(const ForceBatch &batch) const {
    for (size_t i = 0; i < batch.nSystems; ++i)
      static_cast<const Derived *>(this)->forceImpl(batch.in[i], &batch.out[i]);
  }
  void forceBatch(const ForceBatch &batch) override {
    if (batch.nSystems == 0) return;
    static_cast<Derived *>(this)->forceBatchImpl(batch);
    for (size_t i = 0; i < batch.nSystems; ++i) registry<Derived>::incrementForceCalls();
  }
};

} // namespace rgpot
