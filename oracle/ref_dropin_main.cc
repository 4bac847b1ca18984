// ref_dropin_main.cc -- the reference's OWN C++ face of the CuH2 kernel
// (CppCore/rgpot/fortran/FortranPots.cc, compiled unmodified from /root/reference) linked against
// librgpot_b200.so instead of the Fortran objects: `rgpot::fortranpots::CuH2Pot` and `EAMAlPot` then
// evaluate on the B200 without a single changed line upstream.  TEST INFRASTRUCTURE (built into oracle/_ref).
// Runs CppCore/tests/CuH2PotTest.cc:8-43 and the error path of FortranPots.cc:44-56.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "rgpot/fortran/FortranPots.hpp"
#include "rgpot/types/AtomMatrix.hpp"

// FortranPots.cc also references the five potentials this engine does not implement
extern "C" {
#define STUB5(name) int name(int32_t, const double *, const double *, double *, double *) { return 99; }
#define STUB6(name) int name(int32_t, const double *, const int32_t *, const double *, double *, double *) { return 99; }
STUB5(rgpot_sw_force) STUB5(rgpot_edip_force) STUB5(rgpot_lenosky_force) STUB5(rgpot_tersoff_force)
STUB6(rgpot_water_h_force)
}

int main() {
  using rgpot::types::AtomMatrix;
  auto pot = rgpot::fortranpots::CuH2Pot();
  AtomMatrix positions{{0.63940268750835, 0.90484742551374, 6.97516498544584},
                       {3.19652040936288, 0.90417430354811, 6.97547796369474},
                       {8.98363230369760, 9.94703496017833, 7.83556854923689},
                       {7.64080177576300, 9.94703114803832, 7.83556986121272}};
  std::vector<int> atmtypes{29, 29, 1, 1};
  std::array<std::array<double, 3>, 3> box{{{15.345599999999999, 0, 0}, {0, 21.702000000000002, 0}, {0, 0, 100.0}}};
  auto [energy, forces, variance] = pot(positions, atmtypes, box);
  (void)variance;
  int bad = 0;
  const double expE = -2.7114096242662238;
  const double expF[4][3] = {{1.4919411183978113, -3.9273058476626193E-004, 1.8260603127768336E-004},
                             {-1.4919411183978113, 3.9273058476626193E-004, -1.8260603127768336E-004},
                             {-4.9118653085630006, -1.3944215503304855E-005, 4.7990036210569753E-006},
                             {4.9118653085630006, 1.3944215503304855E-005, -4.7990036210569753E-006}};
  if (std::fabs(energy - expE) > 1e-10 * std::fabs(expE)) ++bad;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j)
      if (std::fabs(forces(i, j) - expF[i][j]) > 1e-9) ++bad;
  atmtypes[2] = 26;
  try {
    pot(positions, atmtypes, box);
    ++bad;
  } catch (const std::runtime_error &e) {
    if (std::string(e.what()) != "CuH2 potential failed (status 2): rgpot_cuh2: atom 3 has atomic number 26; "
                                 "this potential covers Cu (29) and H (1) only")
      ++bad;
  }
  {  // CppCore/tests/FortranPotsTest.cc:135-144 through the reference's own EAMAlPot class
    auto al = rgpot::fortranpots::EAMAlPot();
    AtomMatrix fcc{{7.97500, 7.97500, 7.97500}, {10.00000, 10.00000, 7.97500}, {10.00000, 7.97500, 10.00000},
                   {7.97500, 10.00000, 10.00000}};
    std::array<std::array<double, 3>, 3> cube{{{20.0, 0.0, 0.0}, {0.0, 20.0, 0.0}, {0.0, 0.0, 20.0}}};
    auto [e_al, f_al, v_al] = al(fcc, std::vector<int>(4, 13), cube);
    (void)v_al;
    double worst = 0.0;
    for (int i = 0; i < 4; ++i)
      worst = std::fmax(worst, std::sqrt(f_al(i, 0) * f_al(i, 0) + f_al(i, 1) * f_al(i, 1) + f_al(i, 2) * f_al(i, 2)));
    if (std::fabs(e_al - (-5.217864)) > 1e-4 * 5.217864 || std::fabs(worst - 0.968647) > 1e-3 * 0.968647) ++bad;
  }
  {  // CppCore/tests/FortranPotsTest.cc:146-155 through the reference's own FeHePot class
    auto fehe = rgpot::fortranpots::FeHePot();
    AtomMatrix bcc{{7.13000, 7.13000, 7.13000},    {8.56500, 8.56500, 8.56500},   {7.13000, 7.13000, 10.00000},
                   {8.56500, 8.56500, 11.43500},   {7.13000, 10.00000, 7.13000},  {8.56500, 11.43500, 8.56500},
                   {7.13000, 10.00000, 10.00000},  {8.56500, 11.43500, 11.43500}, {10.00000, 7.13000, 7.13000},
                   {11.43500, 8.56500, 8.56500},   {10.00000, 7.13000, 10.00000}, {11.43500, 8.56500, 11.43500},
                   {10.00000, 10.00000, 7.13000},  {11.43500, 11.43500, 8.56500}, {10.00000, 10.00000, 10.00000},
                   {11.43500, 11.43500, 11.43500}};
    std::array<std::array<double, 3>, 3> cube{{{20.0, 0.0, 0.0}, {0.0, 20.0, 0.0}, {0.0, 0.0, 20.0}}};
    auto [e_fe, f_fe, v_fe] = fehe(bcc, std::vector<int>(16, 26), cube);
    (void)v_fe;
    double worst = 0.0;
    for (int i = 0; i < 16; ++i)
      worst = std::fmax(worst, std::sqrt(f_fe(i, 0) * f_fe(i, 0) + f_fe(i, 1) * f_fe(i, 1) + f_fe(i, 2) * f_fe(i, 2)));
    if (std::fabs(e_fe - (-43.959774)) > 1e-4 * 43.959774 || std::fabs(worst - 1.245094) > 1e-3 * 1.245094) ++bad;
  }
  std::printf("reference fortranpots::CuH2Pot / EAMAlPot / FeHePot over librgpot_b200.so: E = %.16f, %s\n", energy,
              bad ? "MISMATCH" : "ok");
  return bad ? 1 : 0;
}
