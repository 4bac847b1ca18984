// physics.cuh -- closed-form pair / density / embedding terms evaluated in FP64.
//
//   LJ   : CppCore/rgpot/LennardJones/LJPot.cc:61-81
//   CuH2 : CppCore/rgpot/fortran/rgpot_cuh2.f90:143-218 (pair), :228-291 (density),
//          :298-370 (embedding), :118-135 (cutoff shifts, computed on the host with libm so they
//          are the same doubles the reference subtracts)
// The CuH2 EAM is fully analytic (no spline tables exist upstream), so the coefficients live in
// __constant__ memory and the kernels evaluate the same closed forms.
//
// The radial terms are exp-bound (the reference evaluates 10 exponentials per directed pair).
// Here a directed Cu-Cu pair costs 3:
//   * alpha_A(CuCu) == 2 alpha_B(CuCu) exactly, so exp(-alpha_A r) = exp(-alpha_B r)^2;
//   * beta_b(Cu) = 2 beta_a(Cu) - 1e-5, so exp(-beta_b r) = exp(-beta_a r)^2 * exp(1e-5 r) with
//     the last factor a 3-term Taylor series (|1e-5 r| < 6.1e-5, truncation < 6e-19);
//   * exp itself is a branch-free range-limited kernel (arguments lie in [-38, 0]).
// Each identity is checked on the host against the coefficient values before it is enabled
// (EamParams::cucu_square / cu_taylor); otherwise the generic two-exponential form runs.
// All of this changes results by a few ulp, far inside the 1e-10 parity budget; accept/reject
// decisions never depend on it.
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
  // physics threshold on d2: pairs with d2 >= rc2_phys are skipped.  It equals
  // min(rc2, smallest d2 whose correctly rounded sqrt reaches rcut), i.e. exactly the
  // reference's `d2 < rc2` (vesin filter) AND `sqrt(d2) < rcut` (rgpot_cuh2.f90:522,554).
  double rc2_phys;
  double cu_delta;  // 2 beta_a(Cu) - beta_b(Cu)
  int cucu_square;  // alpha_A(CuCu) == 2 alpha_B(CuCu)
  int cu_taylor;    // |cu_delta| * rcut small enough for the 3-term series
  double gauge;     // EAM-Al: g_transform (rgpot_eam_al.f90:47), 0 for CuH2
};

// Two analytic EAM models share the kernels.  EAM_CUH2: the Cu-H potential of rgpot_cuh2.f90.
// EAM_AL: the single-species double-exponential aluminium of rgpot_eam_al.f90 -- the same pair /
// density shapes as the Cu-Cu terms (slot 0 of every table) plus a gauge term g: phi carries
// -2 g rho(r), the embedding +g rho (rgpot_eam_al.f90:159-171, 188-213).  Each model has its own
// __constant__ block (handles of both kinds may run concurrently) selected at compile time, so the
// coefficients stay immediate operands of the FP64 instructions.
enum { EAM_CUH2 = 0, EAM_AL = 1, EAM_FEHE = 2 };  // (EAM_FEHE: the batch kernel only; its terms are the fehe_* functions below)
__constant__ EamParams c_eam;
__constant__ EamParams c_eam_al;

template <int M>
RB_D const EamParams& eam_par() {
  if (M == EAM_AL) return c_eam_al;
  return c_eam;
}
// table cutoff of a model (rgpot_fehe.f90:167-176: the largest of the five Fe-He cutoffs)
template <int M>
RB_D double model_rcut() {
  if (M == EAM_FEHE) return 5.4;
  return eam_par<M>().rcut;
}

// ---- pair potentials on the LJ pair search ------------------------------------------------------
// Pair energy U and fs with F_i += fs * (r_i - r_j) for one accepted pair at squared distance r2.
//   LJ    : CppCore/rgpot/LennardJones/LJPot.cc:61-81
//   Morse : CppCore/rgpot/Morse/MorsePot.cc:53-62  (sqrt and exp as the reference calls them)
//   ZBL   : CppCore/rgpot/ZBL/ZBLPot.cc:38-60 (e_zbl, dzbldr), :233-240 (switching)
RB_D void pair_eval(const LJParams& p, double r2, int ti, int tj, double& u, double& fs) {
  if (p.morse == 2) {
    const double* c = p.zbl_table + 10 * (ti * p.zbl_nt + tj);
    const double r = sqrt(r2);
    const double r_inv = 1.0 / r;
    const double e1 = exp(-c[0] * r), e2 = exp(-c[1] * r), e3 = exp(-c[2] * r), e4 = exp(-c[3] * r);
    const double sum = 0.18175 * e1 + 0.50986 * e2 + 0.28022 * e3 + 0.02817 * e4;
    const double sum_p = -0.18175 * c[0] * e1 - 0.50986 * c[1] * e2 - 0.28022 * c[2] * e3 - 0.02817 * c[3] * e4;
    u = c[4] * sum * r_inv + c[9];
    double dEdr = c[4] * (sum_p - sum * r_inv) * r_inv;
    if (r2 > p.zbl_rin2) {
      const double t = r - p.zbl_rin;
      u += t * t * t * (c[7] + c[8] * t);
      dEdr += t * t * (c[5] + c[6] * t);
    }
    fs = -dEdr / r;
  } else if (p.morse == 1) {
    const double r = sqrt(r2);
    const double d = 1.0 - exp(-p.m_a * (r - p.m_re));
    u = p.m_de * d * d - p.m_de - p.m_ecut;
    fs = p.m_two_de_a * d * (d - 1.0) / r;
  } else {
    const double invR2 = 1.0 / r2;
    const double sr2 = p.psi2 * invR2;
    const double a = sr2 * sr2 * sr2;
    const double b = 4.0 * p.u0 * a;
    u = b * (a - 1.0) - p.shiftU;
    fs = 6.0 * b * invR2 * (2.0 * a - 1.0);
  }
}

// one directed entry of the gather form: half the pair energy, the whole force on i.
// d = r_j - r_i (folded / shifted), so F_i -= fs * d   (LJPot.cc:70-77 with d = r_i - r_j).
RB_D void lj_pair(const LJParams& p, int ti, int tj, double dx, double dy, double dz, double r2,
                  double& e, double& fx, double& fy, double& fz) {
  double u, fs;
  pair_eval(p, r2, ti, tj, u, fs);
  e += 0.5 * u;
  fx -= fs * dx;
  fy -= fs * dy;
  fz -= fs * dz;
}

// ---- FP64 building blocks --------------------------------------------------------------------
// exp(x) for x in about [-700, 700] with no special cases:  x = (64 k + j) ln2 / 64 + r,
// |r| <= ln2 / 128, so exp(x) = 2^k * 2^(j/64) * e^r.  n = 64 k + j = rint(64 x log2 e) by the
// magic-number add, a two-term Cody-Waite reduction, 2^(j/64) from a 64-entry table in shared
// memory (correctly rounded doubles, uploaded by the host), e^r as a degree-5 Taylor polynomial
// (truncation r^6 / 720 < 3.4e-17), 2^k patched into the exponent with integer arithmetic.
// About 1 ulp; 10 FP64 instructions and one shared-memory load, no branch.  (The radial terms of
// the EAM are exp-bound: this is 6 FP64 instructions fewer than a table-free degree-11 kernel.)
__constant__ double c_exp2tab[64];
// the same table in global memory: a block fills its shared copy with ONE coalesced 512-byte load
// (64 lanes reading 64 different __constant__ addresses are serialised by the constant cache, eight
// line misses per warp: microseconds at the head of every block)
__device__ double g_exp2tab[64];

RB_D double* exp_table() {
  __shared__ double tab[64];
  return tab;
}
// every thread of the block calls this once, before any thread returns; includes the barrier
RB_D void exp_table_init() {
  double* tab = exp_table();
  for (int t = threadIdx.x; t < 64; t += blockDim.x) tab[t] = g_exp2tab[t];
  __syncthreads();
}

RB_D double exp_fast(double x) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 92.332482616893656758, kMagic);  // 64 log2 e
  const int n = __double2loint(t);
  const double kf = t - kMagic;
  double r = fma(kf, -6.93147180369123816490e-01 / 64.0, x);
  r = fma(kf, -1.90821492927058770002e-10 / 64.0, r);
#ifdef RB_EXP_ESTRIN
  // e^r = (1 + r) + r^2 ((1/2 + r/6) + r^2 (1/24 + r/120)): three levels instead of a five-deep chain
  const double r2 = r * r;
  const double q0 = 1.0 + r, q1 = fma(r, 1.0 / 6.0, 0.5), q2 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  double p = fma(r2, fma(r2, q2, q1), q0);
#else
  double p = 1.0 / 120.0;
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
#endif
  p *= exp_table()[n & 63];
  return __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
}

// 1/sqrt(a) for normal positive a: hardware seed (rsqrt.approx, ~2^-22) + one cubic Newton
// step (error ~ e^3), ~1 ulp, 5 FP64 instructions.
RB_D double rsqrt_fast(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = a * y;
  const double e = fma(-h, y, 1.0);
  const double c = fma(e, 0.375, 0.5);
  return fma(y, e * c, y);
}

// ---- CuH2 EAM ------------------------------------------------------------------------------
RB_D int pair_kind(int zi, int zj) { return zi + zj; }  // Cu=0, H=1: 0 CuCu, 1 HCu, 2 HH

// the two exponentials of the Cu density and r^6 (square-and-multiply like libgfortran: r^2*r^4)
struct CuDensityTerms {
  double ea, eb, r6;
};
template <int M = EAM_CUH2>
RB_D CuDensityTerms cu_density_terms(double r) {
  const EamParams& P = eam_par<M>();
  CuDensityTerms t;
  t.ea = exp_fast(-P.ba[0] * r);
  if (P.cu_taylor) {
    const double z = P.cu_delta * r;
    const double corr = fma(z, fma(z, fma(z, 1.0 / 6.0, 0.5), 1.0), 1.0);
    t.eb = t.ea * t.ea * corr;
  } else {
    t.eb = exp_fast(-P.bb[0] * r);
  }
  const double r2 = r * r;
  t.r6 = r2 * (r2 * r2);
  return t;
}

// density a neighbour of `species` deposits at distance r (shifted)   :228-262
template <int M = EAM_CUH2>
RB_D double eam_rho(int species, double r) {
  if (M == EAM_AL || species == 0) {
    const EamParams& P = eam_par<M>();
    const CuDensityTerms t = cu_density_terms<M>(r);
    return P.scale[0] * t.r6 * fma(P.gamma[0], t.eb, t.ea) - P.rhocut[0];
  }
  // gamma_h = 0 and neta_h = 0: a single exponential (the reference multiplies exp(0) by zero)
  return c_eam.scale[1] * exp_fast(-c_eam.ba[1] * r) - c_eam.rhocut[1];
}

// (drho/dr) / r   :241-291
RB_D double eam_drho(int species, double r, double rinv) {
  if (species == 0) {
    const CuDensityTerms t = cu_density_terms(r);
    const double t2 = c_eam.gamma[0] * t.eb;
    const double t3 = c_eam.scale[0] * t.r6;
    const double raw = t3 * (t.ea + t2);
    return (6.0 * raw * rinv - t3 * fma(c_eam.bb[0], t2, c_eam.ba[0] * t.ea)) * rinv;
  }
  const double t1 = exp_fast(-c_eam.ba[1] * r);
  return -(c_eam.scale[1] * c_eam.ba[1] * t1) * rinv;
}

// pair energy and (dphi/dr)/r   :143-218
RB_D void eam_phi(int kind, double r, double rinv, double& phi, double& dphi_r) {
  double ea, eb;
  eb = exp_fast(-c_eam.ab[kind] * r);
  if (kind == 0 && c_eam.cucu_square) {
    ea = eb * eb;
  } else {
    ea = exp_fast(-c_eam.aa[kind] * r);
  }
  const double t1 = c_eam.da[kind] * ea;
  const double t2 = c_eam.db[kind] * eb;
  phi = t1 + t2 - c_eam.phicut[kind];
  dphi_r = -fma(c_eam.aa[kind], t1, c_eam.ab[kind] * t2) * rinv;
}

// embedding energy and derivative, explicit powers and left-to-right sums like :298-370
// (the Cu polynomial cancels by ~2 orders of magnitude; keep the reference's association)
template <int M = EAM_CUH2>
RB_D void eam_embed(int species, double x, double& emb, double& demb) {
  const double x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x;
  if (M == EAM_AL) {
    // F = g rho + sum_k c_k rho^k, F' = g + c_1 + sum_{k>=2} k c_k rho^(k-1), summed in the
    // reference's order (rgpot_eam_al.f90:188-213)
    const double x6 = x5 * x, x7 = x6 * x, x8 = x7 * x;
    const EamParams& P = eam_par<EAM_AL>();
    const double* f = P.f_cu;
    const double pw[8] = {x, x2, x3, x4, x5, x6, x7, x8};
    emb = __dmul_rn(P.gauge, x);
    demb = __dadd_rn(P.gauge, f[0]);
#pragma unroll
    for (int k = 0; k < 8; ++k) emb = __dadd_rn(emb, __dmul_rn(f[k], pw[k]));
#pragma unroll
    for (int k = 1; k < 8; ++k) demb = __dadd_rn(demb, __dmul_rn(__dmul_rn(f[k], (double)(k + 1)), pw[k - 1]));
  } else if (species == 0) {
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

// r and 1/r of an accepted entry.  d2 < rc2_phys has been checked by the caller, so neither the
// correctly rounded sqrt nor a division is needed any more.
RB_D void radial(double d2, double& r, double& rinv) {
  rinv = rsqrt_fast(d2);
  r = d2 * rinv;
}

// one accepted directed entry of the density pass (atom_density :518-529)
template <int M = EAM_CUH2>
RB_D double eam_density_entry(int zj, double d2) {
  double r, rinv;
  radial(d2, r, rinv);
  return eam_rho<M>(zj, r);
}

// one accepted directed entry of the force pass (atom_force :552-586).
// vec = r_j - r_i + S.H ; returns half pair energy and accumulates f_i += w * vec.
// the pair energy phi and the radial factor w of one accepted pair (symmetric in i <-> j: the same w acts on
// both atoms with opposite sign, which is what the half-list mode of the batch kernel uses)
template <int M = EAM_CUH2>
RB_D void eam_force_pair(int zi, int zj, double dfd_i, double dfd_j, double d2, double& phi, double& w) {
  double r, rinv;
  radial(d2, r, rinv);
  if (M == EAM_AL || (zi | zj) == 0) {
    // Cu-Cu (the bulk of every slab) and Al-Al: every coefficient index is static, so the
    // constants are operands of the FP64 instructions instead of separately loaded registers
    const EamParams& P = eam_par<M>();
    const double eb = exp_fast(-P.ab[0] * r);
    const double ea = P.cucu_square ? eb * eb : exp_fast(-P.aa[0] * r);
    const double t1 = P.da[0] * ea;
    const double t2 = P.db[0] * eb;
    phi = t1 + t2 - P.phicut[0];
    w = -fma(P.aa[0], t1, P.ab[0] * t2) * rinv;
    const CuDensityTerms t = cu_density_terms<M>(r);
    const double g2 = P.gamma[0] * t.eb;
    const double t3 = P.scale[0] * t.r6;
    const double raw = t3 * (t.ea + g2);
    const double drho = (6.0 * raw * rinv - t3 * fma(P.bb[0], g2, P.ba[0] * t.ea)) * rinv;
    if (M == EAM_AL) {
      // gauge partner of the embedding's +g rho: phi -= 2 g rho(r), phi' -= 2 g rho'(r)
      phi = fma(-2.0 * P.gauge, raw - P.rhocut[0], phi);
      w = fma(drho, (dfd_i + dfd_j) - 2.0 * P.gauge, w);
    } else {
      w = fma(drho, dfd_i + dfd_j, w);  // the same density derivative acts at both ends
    }
  } else {
    eam_phi(pair_kind(zi, zj), r, rinv, phi, w);
    const double drho_ij = eam_drho(zj, r, rinv);                         // what j deposits on i
    const double drho_ji = (zi == zj) ? drho_ij : eam_drho(zi, r, rinv);  // what i deposits on j
    w = w + drho_ij * dfd_i + drho_ji * dfd_j;
  }
}

template <int M = EAM_CUH2>
RB_D void eam_force_entry(int zi, int zj, double dfd_i, double dfd_j, double vx, double vy,
                          double vz, double d2, double& e, double& fx, double& fy, double& fz) {
  double phi, w;
  eam_force_pair<M>(zi, zj, dfd_i, dfd_j, d2, phi, w);
  e += 0.5 * phi;
  fx = fma(w, vx, fx);
  fy = fma(w, vy, fy);
  fz = fma(w, vz, fz);
}

// ---- Fe-He EAM (CppCore/rgpot/fortran/rgpot_fehe.f90) ----------------------------------------------
// Three pair terms, two density channels (d-band: Fe-Fe pairs, embedded on Fe; s-band: Fe-He pairs,
// embedded on both), five cutoffs.  Every branch of the reference is a comparison on r = sqrt(d2)
// (the table's IEEE square root), so the kernels take the same correctly rounded sqrt and compare
// the same way: which term, which spline piece and which cutoff applies is decided exactly as
// upstream.  Species: 0 Fe, 1 He.
struct FeHeParams {
  double rcut_fe_pair, rcut_fe_rho, rcut_fehe_pair, rcut_fehe_rho, rcut_he_pair;
  double fe_zbl_scale, fe_zbl_weight[4], fe_zbl_decay[4];
  double fe_bridge[4], fe_bridge_lo, fe_bridge_hi;
  double fe_pair_knot[13], fe_pair_coeff[13];
  double fe_rho_knot[3], fe_rho_coeff[3];
  double fe_emb_quadratic, fe_emb_quartic;
  double fehe_knot[8], fehe_coeff[8];
  double sband_amp, sband_decay, sband_emb_sqrt, sband_emb_quadratic, sband_emb_quartic;
  double he_rm, he_c6, he_c8, he_c10, he_aa, he_alpha, he_beta, he_damp, he_eps;
  double rho_floor;
};
__constant__ FeHeParams c_fehe;

template <int N>
RB_D void knot_spline(const double* knot, const double* coeff, double r, double& v, double& dv) {
  // sum_k a_k max(r_k - r, 0)^3 and its derivative -3 sum_k a_k max(r_k - r, 0)^2   :194, :223
  double s3 = 0.0, s2 = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double t = fmax(knot[k] - r, 0.0);
    const double t2 = t * t;
    s2 += coeff[k] * t2;
    s3 += coeff[k] * (t2 * t);
  }
  v = s3;
  dv = -3.0 * s2;
}

// pair energy and dV/dr of the species combination   :182-304, :421-438
RB_D void fehe_pair(int zi, int zj, double r, double& v, double& dv) {
  const FeHeParams& P = c_fehe;
  v = dv = 0.0;
  if ((zi | zj) == 0) {  // Fe-Fe
    if (r > P.rcut_fe_pair) return;
    if (r < P.fe_bridge_lo) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double sc = P.fe_zbl_weight[k] * exp(-P.fe_zbl_decay[k] * r);
        s1 += sc;
        s2 += P.fe_zbl_decay[k] * sc;
      }
      v = P.fe_zbl_scale / r * s1;
      dv = -P.fe_zbl_scale / (r * r) * s1 - P.fe_zbl_scale / r * s2;
    } else if (r < P.fe_bridge_hi) {
      const double expo = exp(P.fe_bridge[0] + P.fe_bridge[1] * r + P.fe_bridge[2] * r * r + P.fe_bridge[3] * r * r * r);
      v = expo;
      dv = expo * (P.fe_bridge[1] + 2.0 * P.fe_bridge[2] * r + 3.0 * P.fe_bridge[3] * r * r);
    } else {
      knot_spline<13>(P.fe_pair_knot, P.fe_pair_coeff, r, v, dv);
    }
  } else if ((zi & zj) == 1) {  // He-He, HFD-B
    if (r > P.rcut_he_pair) return;
    const double x = r / P.he_rm;
    const double x2 = x * x, x4 = x2 * x2, x6 = x2 * x4, x8 = x4 * x4, x10 = x2 * x8;
    const double dispersion = P.he_c6 / x6 + P.he_c8 / x8 + P.he_c10 / x10;
    const double ddispersion = -(6.0 * P.he_c6 / (x6 * x) + 8.0 * P.he_c8 / (x8 * x) + 10.0 * P.he_c10 / (x10 * x));
    const double core = P.he_aa * exp(-P.he_alpha * x + P.he_beta * x * x);
    double damp = 1.0, ddamp = 0.0;
    if (x < P.he_damp) {
      const double q = P.he_damp / x - 1.0;
      damp = exp(-(q * q));
      ddamp = damp * 2.0 * P.he_damp * q / (x * x);
    }
    v = P.he_eps * (core - damp * dispersion);
    dv = P.he_eps * (core * (-P.he_alpha + 2.0 * P.he_beta * x) - ddamp * dispersion - damp * ddispersion) / P.he_rm;
  } else {  // Fe-He
    if (r > P.rcut_fehe_pair) return;
    knot_spline<8>(P.fehe_knot, P.fehe_coeff, r, v, dv);
  }
}

// density the pair feeds into ITS channel (d-band for Fe-Fe, s-band for Fe-He, none for He-He) and
// the radial derivative   :311-355, :442-473.  Returns the channel: 0 d-band, 1 s-band, -1 none.
RB_D int fehe_density(int zi, int zj, double r, double& rho, double& drho) {
  const FeHeParams& P = c_fehe;
  rho = drho = 0.0;
  if ((zi | zj) == 0) {
    if (!(r > P.rcut_fe_rho)) knot_spline<3>(P.fe_rho_knot, P.fe_rho_coeff, r, rho, drho);
    return 0;
  }
  if (zi != zj) {
    if (!(r > P.rcut_fehe_rho)) {
      const double ex = exp(-2.0 * P.sband_decay * r);
      rho = P.sband_amp * (r * (r * r)) * ex;
      drho = P.sband_amp * (r * r) * ex * (3.0 - 2.0 * P.sband_decay * r);
    }
    return 1;
  }
  return -1;
}

// embedding of both channels   :363-413
RB_D void fehe_embed(int species, double rho_d, double rho_s, double& e, double& dfd, double& dfs) {
  const FeHeParams& P = c_fehe;
  double ed = 0.0, es = 0.0;
  dfd = dfs = 0.0;
  if (species == 0 && !(rho_d < P.rho_floor)) {
    const double sq = sqrt(rho_d), r2 = rho_d * rho_d;
    ed = -sq + P.fe_emb_quadratic * r2 + P.fe_emb_quartic * (r2 * r2);
    dfd = -0.5 / sq + 2.0 * P.fe_emb_quadratic * rho_d + 4.0 * P.fe_emb_quartic * (rho_d * r2);
  }
  if (!(rho_s < P.rho_floor)) {
    const double sq = sqrt(rho_s), r2 = rho_s * rho_s;
    es = P.sband_emb_sqrt * sq + P.sband_emb_quadratic * r2 + P.sband_emb_quartic * (r2 * r2);
    dfs = 0.5 * P.sband_emb_sqrt / sq + 2.0 * P.sband_emb_quadratic * rho_s + 4.0 * P.sband_emb_quartic * (rho_s * r2);
  }
  e = ed + es;
}

}  // namespace rb
