// physics.cuh -- closed-form pair / density / embedding terms evaluated in FP64.
//
//   LJ   : CppCore/rgpot/LennardJones/LJPot.cc:61-81
//   CuH2 : CppCore/rgpot/fortran/rgpot_cuh2.f90:143-218 (pair), :228-291 (density),
//          :298-370 (embedding), :118-135 (cutoff shifts, computed on the host with libm so they
//          are the same doubles the reference subtracts)
// The CuH2 EAM is fully analytic (no spline tables exist upstream), so the coefficients live in
// __constant__ memory and the kernels evaluate the same closed forms.
#pragma once

#include "common.cuh"

namespace rb {

struct EamParams {
  // pair kind: 0 CuCu, 1 HCu, 2 HH
  double da[3], aa[3], db[3], ab[3], phicut[3];
  // species: 0 Cu, 1 H
  double scale[2], ba[2], bb[2], gamma[2], rhocut[2];
  double f_cu[8], f_h[5];
  double rcut, rc2;
};

__constant__ EamParams c_eam;

// ---- Lennard-Jones -----------------------------------------------------------------------
// one directed entry of the gather form: half the pair energy, the whole force on i.
// d = r_j - r_i (folded / shifted), so F_i -= fs * d   (LJPot.cc:70-77 with d = r_i - r_j).
RB_D void lj_pair(const LJParams& p, double dx, double dy, double dz, double r2, double& e,
                  double& fx, double& fy, double& fz) {
  const double invR2 = 1.0 / r2;
  const double sr2 = p.psi2 * invR2;
  const double a = sr2 * sr2 * sr2;
  const double b = 4.0 * p.u0 * a;
  e += 0.5 * (b * (a - 1.0) - p.shiftU);
  const double fs = 6.0 * b * invR2 * (2.0 * a - 1.0);
  fx -= fs * dx;
  fy -= fs * dy;
  fz -= fs * dz;
}

// ---- CuH2 EAM ------------------------------------------------------------------------------
RB_D int pair_kind(int zi, int zj) { return zi + zj; }  // Cu=0, H=1: 0 CuCu, 1 HCu, 2 HH

// r ** neta with neta_cu = 6 (square-and-multiply like libgfortran: r^2 * r^4), neta_h = 0
RB_D double rpow(int species, double r) {
  if (species == 1) return 1.0;
  const double r2 = r * r;
  return r2 * (r2 * r2);
}

// density a neighbour of `species` deposits at distance r (shifted)   :228-262
RB_D double eam_rho(int species, double r) {
  const double t1 = exp(-c_eam.ba[species] * r);
  // gamma_h = 0: the reference multiplies exp(0) by zero, which is exactly 0
  const double t2 = species == 0 ? c_eam.gamma[0] * exp(-c_eam.bb[0] * r) : 0.0;
  const double t3 = c_eam.scale[species] * rpow(species, r);
  return t3 * (t1 + t2) - c_eam.rhocut[species];
}

// (drho/dr) / r   :241-291
RB_D double eam_drho(int species, double r, double rinv) {
  const double t1 = exp(-c_eam.ba[species] * r);
  const double t2 = species == 0 ? c_eam.gamma[0] * exp(-c_eam.bb[0] * r) : 0.0;
  const double t3 = c_eam.scale[species] * rpow(species, r);
  const double raw = t3 * (t1 + t2);
  const double neta = species == 0 ? 6.0 : 0.0;
  return (neta * raw * rinv - t3 * (c_eam.ba[species] * t1 + c_eam.bb[species] * t2)) * rinv;
}

// pair energy and (dphi/dr)/r   :143-218
RB_D void eam_phi(int kind, double r, double rinv, double& phi, double& dphi_r) {
  const double t1 = c_eam.da[kind] * exp(-c_eam.aa[kind] * r);
  const double t2 = c_eam.db[kind] * exp(-c_eam.ab[kind] * r);
  phi = t1 + t2 - c_eam.phicut[kind];
  dphi_r = -(c_eam.aa[kind] * t1 + c_eam.ab[kind] * t2) * rinv;
}

// embedding energy and derivative, explicit powers and left-to-right sums like :298-370
// (the Cu polynomial cancels by ~2 orders of magnitude; keep the reference's association)
RB_D void eam_embed(int species, double x, double& emb, double& demb) {
  const double x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x;
  if (species == 0) {
    const double x6 = x5 * x, x7 = x6 * x, x8 = x7 * x;
    const double* f = c_eam.f_cu;
    emb = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(
              __dmul_rn(f[0], x), __dmul_rn(f[1], x2)), __dmul_rn(f[2], x3)), __dmul_rn(f[3], x4)),
              __dmul_rn(f[4], x5)), __dmul_rn(f[5], x6)), __dmul_rn(f[6], x7)), __dmul_rn(f[7], x8));
    demb = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(
               f[0], __dmul_rn(__dmul_rn(2.0, f[1]), x)), __dmul_rn(__dmul_rn(3.0, f[2]), x2)),
               __dmul_rn(__dmul_rn(4.0, f[3]), x3)), __dmul_rn(__dmul_rn(5.0, f[4]), x4)),
               __dmul_rn(__dmul_rn(6.0, f[5]), x5)), __dmul_rn(__dmul_rn(7.0, f[6]), x6)),
               __dmul_rn(__dmul_rn(8.0, f[7]), x7));
  } else {
    const double* f = c_eam.f_h;
    emb = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(
              __dmul_rn(f[0], x), __dmul_rn(f[1], x2)), __dmul_rn(f[2], x3)), __dmul_rn(f[3], x4)),
              __dmul_rn(f[4], x5));
    demb = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(
               f[0], __dmul_rn(__dmul_rn(2.0, f[1]), x)), __dmul_rn(__dmul_rn(3.0, f[2]), x2)),
               __dmul_rn(__dmul_rn(4.0, f[3]), x3)), __dmul_rn(__dmul_rn(5.0, f[4]), x4));
  }
}

// one accepted directed entry of the force pass (atom_force :552-586).
// vec = r_j - r_i + S.H ; returns half pair energy and accumulates f_i += w * vec.
RB_D void eam_force_entry(int zi, int zj, double dfd_i, double dfd_j, double vx, double vy,
                          double vz, double r, double& e, double& fx, double& fy, double& fz) {
  const double rinv = 1.0 / r;
  double phi, w;
  eam_phi(pair_kind(zi, zj), r, rinv, phi, w);
  e += 0.5 * phi;
  const double drho_ij = eam_drho(zj, r, rinv);                    // what j deposits on i
  const double drho_ji = (zi == zj) ? drho_ij : eam_drho(zi, r, rinv);  // what i deposits on j
  w = w + drho_ij * dfd_i + drho_ji * dfd_j;
  fx += w * vx;
  fy += w * vy;
  fz += w * vz;
}

}  // namespace rb
