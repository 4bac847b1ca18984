// test_host.cc -- the reference's own C++ tests for this path, re-expressed without Catch2 and
// run through the plugin classes (operator(), forceBatch).  Exit code 0 = all checks passed.
//   CuH2PotTest.cc:8-43, LJClusterPotTest.cc:55-126, FortranPots.cc:44-56 (error text)
#include <cmath>
#include <cstdio>
#include <string>

#include "cuda_pots.hpp"

using rgpot::types::AtomMatrix;
static int failures = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
      ++failures;                                                       \
    }                                                                   \
  } while (0)

static bool within_rel(double a, double b, double tol) { return std::fabs(a - b) <= tol * std::fabs(b); }

int main() {
  // ---- CuH2PotTest.cc:8-43
  rgpot::CuH2CudaPot cuh2;
  AtomMatrix pos{{0.63940268750835, 0.90484742551374, 6.97516498544584},
                 {3.19652040936288, 0.90417430354811, 6.97547796369474},
                 {8.98363230369760, 9.94703496017833, 7.83556854923689},
                 {7.64080177576300, 9.94703114803832, 7.83556986121272}};
  std::vector<int> types{29, 29, 1, 1};
  std::array<std::array<double, 3>, 3> box{{{15.345599999999999, 0, 0}, {0, 21.702000000000002, 0}, {0, 0, 100.0}}};
  auto [energy, forces, variance] = cuh2(pos, types, box);
  const double expF[4][3] = {{1.4919411183978113, -3.9273058476626193E-004, 1.8260603127768336E-004},
                             {-1.4919411183978113, 3.9273058476626193E-004, -1.8260603127768336E-004},
                             {-4.9118653085630006, -1.3944215503304855E-005, 4.7990036210569753E-006},
                             {4.9118653085630006, 1.3944215503304855E-005, -4.7990036210569753E-006}};
  CHECK(within_rel(energy, -2.7114096242662238, 1e-10));
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) CHECK(std::fabs(forces(i, j) - expF[i][j]) < 1e-9);
  CHECK(variance == 0.0);
  CHECK(cuh2.get_type() == rgpot::PotType::CuH2);
  CHECK(cuh2.caps().batched && cuh2.caps().reentrancy == rgpot::Reentrancy::PerInstance);
  CHECK(rgpot::registry<rgpot::CuH2CudaPot>::forceCalls.load() == 1);

  // foreign species: exception text of FortranPots.cc:44-56 + rgpot_cuh2.f90:481-486
  std::vector<int> foreign{29, 29, 26, 1};
  try {
    cuh2(pos, foreign, box);
    CHECK(!"no exception");
  } catch (const std::runtime_error &e) {
    CHECK(std::string(e.what()) ==
          "CuH2 potential failed (status 2): rgpot_cuh2: atom 3 has atomic number 26; this potential "
          "covers Cu (29) and H (1) only");
  }

  // ---- forceBatch (Potential.hpp:251-312): three systems, different sizes
  AtomMatrix dimer{{0.0, 0.0, 0.0}, {1.5, 0.0, 0.0}};
  std::vector<int> dz{29, 1};
  const double dbox[9] = {10, 0, 0, 0, 10, 0, 0, 0, 10};
  double flat[9];
  std::memcpy(flat, &box, sizeof flat);
  AtomMatrix f0 = AtomMatrix::Zero(4, 3), f1 = AtomMatrix::Zero(2, 3), f2 = AtomMatrix::Zero(4, 3);
  rgpot::ForceInput in[3] = {{4, pos.data(), types.data(), flat}, {2, dimer.data(), dz.data(), dbox},
                             {4, pos.data(), types.data(), flat}};
  rgpot::ForceOut out[3] = {{f0.data(), 0, 0}, {f1.data(), 0, 0}, {f2.data(), 0, 0}};
  cuh2.forceBatch(rgpot::ForceBatch{3, in, out});
  CHECK(within_rel(out[0].energy, -2.7114096242662238, 1e-10));
  CHECK(within_rel(out[1].energy, -0.678808, 1e-6)); // tests/rpc_integ.py:76
  CHECK(within_rel(f1(0, 0), -7.5565221, 1e-6) && within_rel(f1(1, 0), 7.5565221, 1e-6));
  CHECK(out[2].energy == out[0].energy);
  CHECK(rgpot::registry<rgpot::CuH2CudaPot>::forceCalls.load() == 4);
  {  // result-cache keys of the same batch (Potential.hpp:324-336): equal systems, equal keys; the
     // potential type enters the key
    const std::vector<uint64_t> keys = cuh2.cacheKeys(rgpot::ForceBatch{3, in, out});
    CHECK(keys.size() == 3 && keys[0] == keys[2] && keys[0] != keys[1] && keys[0] != 0);
  }

  // ---- LJClusterPotTest.cc:55-126
  AtomMatrix cluster{{50.594227, 52.017165, 52.277482}, {51.578540, 52.642148, 51.195222},
                     {51.243758, 52.919086, 52.218383}, {50.490127, 52.736378, 51.417663},
                     {51.522643, 50.884535, 51.656125}, {50.927263, 51.743087, 51.272643},
                     {51.240676, 50.960436, 50.573042}, {51.275727, 52.061877, 50.284160},
                     {50.434110, 50.976423, 51.879062}, {51.695007, 51.919472, 52.038680},
                     {51.363596, 52.199101, 53.064923}, {52.017483, 51.639074, 51.008389},
                     {51.187843, 51.165217, 52.678225}};
  std::array<std::array<double, 3>, 3> cell{{{101.942400, 0, 0}, {0, 103.142600, 0}, {0, 0, 102.605500}}};
  std::vector<int> ones(13, 1);
  rgpot::LJClusterCudaPot ljc;
  rgpot::LJCudaPot lj;
  auto [e_c, f_c, v_c] = ljc(cluster, ones, cell);
  auto [e_p, f_p, v_p] = lj(cluster, ones, cell);
  CHECK(within_rel(e_c, -39.965379, 1e-4));
  CHECK(within_rel(e_c, -39.9653513626179, 1e-9));
  double worst = 0;
  for (size_t i = 0; i < 13; ++i)
    worst = std::fmax(worst, std::sqrt(f_c(i, 0) * f_c(i, 0) + f_c(i, 1) * f_c(i, 1) + f_c(i, 2) * f_c(i, 2)));
  CHECK(within_rel(worst, 0.00470419517301762, 1e-9));
  CHECK(within_rel(e_c, e_p, 1e-12));
  for (size_t i = 0; i < 13; ++i)
    for (size_t j = 0; j < 3; ++j) CHECK(std::fabs(f_c(i, j) - f_p(i, j)) < 1e-12);
  std::array<std::array<double, 3>, 3> tight{{{5, 0, 0}, {0, 5, 0}, {0, 0, 5}}};
  auto [e_t, f_t, v_t] = ljc(cluster, ones, tight);
  CHECK(within_rel(e_t, e_c, 1e-12));
  CHECK(!ljc.caps().periodic && lj.caps().periodic);
  rgpot::LJCudaConfig wider;
  wider.cutoff = 20.0;
  CHECK(rgpot::LJClusterCudaPot{}.paramsKey() == ljc.paramsKey());
  CHECK(rgpot::LJClusterCudaPot{wider}.paramsKey() != ljc.paramsKey());
  CHECK(ljc.get_type() == rgpot::PotType::LJCluster && lj.get_type() == rgpot::PotType::LJ);
  CHECK(ljc.paramsKey() == lj.paramsKey());
  (void)v_c; (void)v_p; (void)v_t; (void)f_t;

  // ---- MorsePotTest.cc:62-117
  {
    rgpot::MorseCudaPot morse;
    CHECK(std::fabs(morse.energyShift() - (-3.5537997279178057e-05)) < 1e-14);
    AtomMatrix two{{10.0, 10.0, 10.0}, {10.0 + 2.8970, 10.0, 10.0}};
    std::vector<int> pt{78, 78};
    std::array<std::array<double, 3>, 3> b40{{{40, 0, 0}, {0, 40, 0}, {0, 0, 40}}};
    auto [e2, f2, v2] = morse(two, pt, b40);
    (void)v2;
    CHECK(std::fabs(e2 - (-0.71016446200272088)) < 1e-12);
    for (size_t i = 0; i < 2; ++i)
      for (size_t j = 0; j < 3; ++j) CHECK(std::fabs(f2(i, j)) < 1e-12);
    auto [em, fm, vm] = morse(cluster, ones, cell);
    (void)vm;
    CHECK(within_rel(em, 7483.1166203812, 1e-9));
    const double row0[3] = {-2056.95286055812, 333.678513150438, 1022.83140169994};
    for (size_t j = 0; j < 3; ++j) CHECK(within_rel(fm(0, j), row0[j], 1e-9));
    double wm = 0, net[3] = {0, 0, 0};
    for (size_t i = 0; i < 13; ++i) {
      wm = std::fmax(wm, std::sqrt(fm(i, 0) * fm(i, 0) + fm(i, 1) * fm(i, 1) + fm(i, 2) * fm(i, 2)));
      for (size_t j = 0; j < 3; ++j) net[j] += fm(i, j);
    }
    CHECK(within_rel(wm, 2480.28362300107, 1e-9));
    for (double s : net) CHECK(std::fabs(s) < 1e-9);
    rgpot::MorseCudaConfig deeper;
    deeper.De = 0.8;
    CHECK(rgpot::MorseCudaPot{}.paramsKey() == morse.paramsKey());
    CHECK(rgpot::MorseCudaPot{deeper}.paramsKey() != morse.paramsKey());
    CHECK(morse.get_type() == rgpot::PotType::Morse);
  }

  // ---- ZBLPotTest.cc:36-142
  {
    rgpot::ZBLCudaPot zbl;
    AtomMatrix dimer2{{9.40, 9.35, 9.30}, {10.60, 10.65, 10.70}};  // siAuDimer(), ZBLPotTest.cc:26-28
    std::array<std::array<double, 3>, 3> c20{{{20, 0, 0}, {0, 20, 0}, {0, 0, 20}}};
    std::array<std::array<double, 3>, 3> c3{{{3, 0, 0}, {0, 3, 0}, {0, 0, 3}}};
    auto [e, f, v] = zbl(dimer2, std::vector<int>{14, 79}, c20);
    (void)v;
    CHECK(std::fabs(e - 0.38537731) < 1e-8);
    const double ef[2][3] = {{-2.37926, -2.57753, -2.7758}, {2.37926, 2.57753, 2.7758}};
    for (size_t i = 0; i < 2; ++i)
      for (size_t j = 0; j < 3; ++j) CHECK(std::fabs(f(i, j) - ef[i][j]) < 1e-5);
    auto [e_au, f_au, v_au] = zbl(dimer2, std::vector<int>{79, 79}, c20);
    auto [e_si, f_si, v_si] = zbl(dimer2, std::vector<int>{14, 14}, c20);
    auto [e_again, f_again, v_again] = zbl(dimer2, std::vector<int>{14, 79}, c20);
    (void)f_au; (void)v_au; (void)f_si; (void)v_si; (void)f_again; (void)v_again;
    CHECK(within_rel(e, 0.38537730883, 1e-9));
    CHECK(within_rel(e_au, 0.961609764914, 1e-9));
    CHECK(within_rel(e_si, 0.167604056792, 1e-9));
    CHECK(e_again == e);
    auto [e_tight, f_tight, v_tight] = zbl(dimer2, std::vector<int>{14, 79}, c3);
    (void)f_tight; (void)v_tight;
    CHECK(e_tight == e);
    CHECK(!zbl.caps().periodic);
    AtomMatrix at_cut{{5.0, 5.0, 5.0}, {7.5, 5.0, 5.0}};
    auto [e_cut, f_cut, v_cut] = zbl(at_cut, std::vector<int>{14, 79}, c20);
    (void)v_cut;
    CHECK(std::fabs(e_cut) < 1e-12);
    for (size_t i = 0; i < 2; ++i)
      for (size_t j = 0; j < 3; ++j) CHECK(std::fabs(f_cut(i, j)) < 1e-12);
    bool threw = false;
    try {
      rgpot::ZBLCudaConfig bad;
      bad.cut_inner = 2.5;
      bad.cut_global = 2.0;
      rgpot::ZBLCudaPot p{bad};
    } catch (const std::invalid_argument &) {
      threw = true;
    }
    CHECK(threw);
    rgpot::ZBLCudaConfig wide;
    wide.cut_global = 4.5;
    CHECK(rgpot::ZBLCudaPot{}.paramsKey() == zbl.paramsKey());
    CHECK(rgpot::ZBLCudaPot{wide}.paramsKey() != zbl.paramsKey());
  }

  {  // EAM aluminium matches the eOn reference, CppCore/tests/FortranPotsTest.cc:135-144
    rgpot::EAMAlCudaPot al;
    AtomMatrix fcc{{7.97500, 7.97500, 7.97500}, {10.00000, 10.00000, 7.97500}, {10.00000, 7.97500, 10.00000},
                   {7.97500, 10.00000, 10.00000}};
    const std::vector<int> types(4, 13);
    std::array<std::array<double, 3>, 3> cube{{{20.0, 0.0, 0.0}, {0.0, 20.0, 0.0}, {0.0, 0.0, 20.0}}};
    auto [e, f, v] = al(fcc, types, cube);
    (void)v;
    CHECK(within_rel(e, -5.217864, 1e-4));
    double worst = 0.0, net[3] = {0, 0, 0};
    for (size_t i = 0; i < 4; ++i) {
      worst = std::max(worst, std::sqrt(f(i, 0) * f(i, 0) + f(i, 1) * f(i, 1) + f(i, 2) * f(i, 2)));
      for (size_t j = 0; j < 3; ++j) net[j] += f(i, j);
    }
    CHECK(within_rel(worst, 0.968647, 1e-3));
    CHECK(std::fabs(net[0]) < 1e-9 && std::fabs(net[1]) < 1e-9 && std::fabs(net[2]) < 1e-9);
    CHECK(al.get_type() == rgpot::PotType::EAMAl);
  }

  std::printf(failures ? "test_host: %d FAILURES\n" : "test_host: ok\n", failures);
  return failures ? 1 : 0;
}
