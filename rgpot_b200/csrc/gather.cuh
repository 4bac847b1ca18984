// gather.cuh -- full-list gather kernels over the GPU cell list (one large system).
//
// Every atom walks the stencil of its cell and owns all of its own sums (full lists, no
// Newton-3 scatter), so there are no floating-point atomics and the summation order is fixed.
//
// Three geometries decide which (j, image) candidates exist and how the separation is formed;
// each reproduces the reference's floating-point expression for the accept test exactly:
//   GEO_LJ_SHIFT  LJ reference semantics, fast form.  Valid when every periodic direction has
//                 >= 3 cells and no atom sits outside the cell: the nearest image of a stencil
//                 neighbour is the one the cell wrap predicts, and
//                 d - w*round(d/w) == d + w*q bit for bit (w*q is exact for q in {-1,0,1}).
//                 (vesin_visit.hpp:97-104)
//   GEO_LJ_FOLD   LJ reference semantics, literal fold arithmetic; any box, any cutoff; each
//                 distinct cell visited once.                      (vesin_visit.hpp:23-25,97-104)
//   GEO_IMAGES    vesin semantics: every image with d2 < rc2, vec = (r_j - r_i) + S.H,
//                 self images dropped.  (vesin-single-build.cpp:1199-1202,1489-1513;
//                 rgpot_neighbors.f90:126-134)
#pragma once

#include "cells.cuh"
#include "physics.cuh"

namespace rb {

enum { GEO_LJ_SHIFT = 0, GEO_LJ_FOLD = 1, GEO_IMAGES = 2 };

struct GatherArgs {
  const double4* spos;  // cell-sorted atoms
  const int* scell;     // cell of each sorted atom
  const int* starts;    // CSR cell starts
  int n;
  Box box;
  Grid grid;
  LJParams lj;
  double rc2;           // GEO_IMAGES accept threshold (strict <)
  const Status* st_in;  // flags from binning (kFlagAtomShift selects the per-pair shift form)
  int only_if_big_shift;   // 1: run only when some atom sits more than one cell vector outside the
                           //    cell (the brick kernels of brick.cuh cover everything else)
  int keep_self_images;    // GEO_IMAGES: keep (i, i, S != 0) entries like vesin does (the EAM tables
                           //    of rgpot drop them, rgpot_neighbors.f90:132)
};

// optional pair emission for rgpot_b200_neighbors
struct EmitArgs {
  int32_t* pairs;   // 2 per entry
  int32_t* shifts;  // 3 per entry
  unsigned long long* count;
  long cap;
  double* vecs;     // optional: 3 per entry, (p_j - p_i) + S.H exactly as evaluated for the test
  double* dists;    // optional: sqrt(d2) (IEEE, like vesin's)
  int half;         // GEO_IMAGES: vesin's half list (vesin-single-build.cpp:1165-1197)
};

// Walk all candidates of sorted atom `i`.  body(j, mj, dx, dy, dz, d2, S0, S1, S2) is called for
// every ACCEPTED candidate with d = r_j - r_i (+ image), in a fixed order.
template <int GEO, bool ATOM_SHIFTS, class Body>
RB_D void walk_neighbors(const GatherArgs& a, int i, const double4 pi, Body&& body) {
  const Grid& g = a.grid;
  const int cell = a.scell[i];
  const int cx = cell % g.nc[0];
  const int cy = (cell / g.nc[0]) % g.nc[1];
  const int cz = cell / (g.nc[0] * g.nc[1]);
  const unsigned long long mi = (unsigned long long)__double_as_longlong(pi.w);
  const int oi = meta_orig(mi);
  const int si0 = meta_shift(mi, 0), si1 = meta_shift(mi, 1), si2 = meta_shift(mi, 2);

  for (int dz = -g.ns_lo[2]; dz <= g.ns_hi[2]; ++dz) {
    int qz = 0, rz = cz + dz;
    if (a.box.periodic[2]) {
      divmod_i(cz + dz, g.nc[2], qz, rz);
    } else if (rz < 0 || rz >= g.nc[2]) {
      continue;
    }
    for (int dy = -g.ns_lo[1]; dy <= g.ns_hi[1]; ++dy) {
      int qy = 0, ry = cy + dy;
      if (a.box.periodic[1]) {
        divmod_i(cy + dy, g.nc[1], qy, ry);
      } else if (ry < 0 || ry >= g.nc[1]) {
        continue;
      }
      for (int dx = -g.ns_lo[0]; dx <= g.ns_hi[0]; ++dx) {
        int qx = 0, rx = cx + dx;
        if (a.box.periodic[0]) {
          divmod_i(cx + dx, g.nc[0], qx, rx);
        } else if (rx < 0 || rx >= g.nc[0]) {
          continue;
        }
        const int nb = (rz * g.nc[1] + ry) * g.nc[0] + rx;
        const int jb = a.starts[nb], je = a.starts[nb + 1];
        if (jb == je) continue;

        // image offset common to the whole stencil cell (atoms all inside the box)
        double svx = 0.0, svy = 0.0, svz = 0.0;
        if (GEO == GEO_LJ_SHIFT) {
          svx = xmul(a.lj.w[0], (double)qx);
          svy = xmul(a.lj.w[1], (double)qy);
          svz = xmul(a.lj.w[2], (double)qz);
        } else if (GEO == GEO_IMAGES && !ATOM_SHIFTS) {
          svx = shift_component(a.box.H, 0, (double)qx, (double)qy, (double)qz);
          svy = shift_component(a.box.H, 1, (double)qx, (double)qy, (double)qz);
          svz = shift_component(a.box.H, 2, (double)qx, (double)qy, (double)qz);
        }

        for (int j = jb; j < je; ++j) {
          const double4 pj = a.spos[j];
          const unsigned long long mj = (unsigned long long)__double_as_longlong(pj.w);
          double ddx, ddy, ddz, d2;
          int S0 = qx, S1 = qy, S2 = qz;
          if (GEO == GEO_LJ_SHIFT) {
            if (j == i) continue;
            ddx = xadd(xsub(pj.x, pi.x), svx);
            ddy = xadd(xsub(pj.y, pi.y), svy);
            ddz = xadd(xsub(pj.z, pi.z), svz);
            d2 = xnorm2(ddx, ddy, ddz);
            if (!(d2 <= a.lj.rc2)) continue;
          } else if (GEO == GEO_LJ_FOLD) {
            if (j == i) continue;
            ddx = fold_component(xsub(pj.x, pi.x), a.lj.w[0], a.lj.inv[0]);
            ddy = fold_component(xsub(pj.y, pi.y), a.lj.w[1], a.lj.inv[1]);
            ddz = fold_component(xsub(pj.z, pi.z), a.lj.w[2], a.lj.inv[2]);
            d2 = xnorm2(ddx, ddy, ddz);
            if (!(d2 <= a.lj.rc2)) continue;
          } else {
            if (meta_orig(mj) == oi && !a.keep_self_images) continue;  // self images carry no usable direction
            if (ATOM_SHIFTS) {
              S0 = qx + si0 - meta_shift(mj, 0);
              S1 = qy + si1 - meta_shift(mj, 1);
              S2 = qz + si2 - meta_shift(mj, 2);
              svx = shift_component(a.box.H, 0, (double)S0, (double)S1, (double)S2);
              svy = shift_component(a.box.H, 1, (double)S0, (double)S1, (double)S2);
              svz = shift_component(a.box.H, 2, (double)S0, (double)S1, (double)S2);
            }
            if (meta_orig(mj) == oi && (S0 | S1 | S2) == 0) continue;  // the atom itself
            ddx = xadd(xsub(pj.x, pi.x), svx);
            ddy = xadd(xsub(pj.y, pi.y), svy);
            ddz = xadd(xsub(pj.z, pi.z), svz);
            d2 = xnorm2(ddx, ddy, ddz);
            if (!(d2 < a.rc2)) continue;
          }
          body(j, mj, ddx, ddy, ddz, d2, S0, S1, S2);
        }
      }
    }
  }
}

// Pick the geometry form the binning flags allow: GEO_LJ_SHIFT is only valid while no atom sits
// outside the cell (otherwise the literal fold decides), GEO_IMAGES switches to per-pair shifts.
template <int GEO, class Body>
RB_D void walk_dispatch(const GatherArgs& a, int i, const double4 pi, Body&& body) {
  const bool atom_shifts = (a.st_in->flags & kFlagAtomShift) != 0;
  if (GEO == GEO_LJ_SHIFT) {
    if (atom_shifts) {
      walk_neighbors<GEO_LJ_FOLD, false>(a, i, pi, body);
    } else {
      walk_neighbors<GEO_LJ_SHIFT, false>(a, i, pi, body);
    }
  } else if (GEO == GEO_LJ_FOLD) {
    walk_neighbors<GEO_LJ_FOLD, false>(a, i, pi, body);
  } else {
    if (atom_shifts) {
      walk_neighbors<GEO_IMAGES, true>(a, i, pi, body);
    } else {
      walk_neighbors<GEO_IMAGES, false>(a, i, pi, body);
    }
  }
}

// ---- LJ ------------------------------------------------------------------------------------
// GEO_LJ_SHIFT / GEO_LJ_FOLD: reference semantics.  GEO_IMAGES: opt-in all-image triclinic LJ.
template <int GEO>
__global__ void __launch_bounds__(128)
k_lj_gather(GatherArgs a, double* __restrict__ F, double* __restrict__ eatom) {
  if (a.only_if_big_shift && !(a.st_in->flags & kFlagBigShift)) return;
  // grid-stride: as a stand-by behind the brick kernels the launch is a few blocks per SM
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    const double4 pi = a.spos[i];
    const int ti = meta_species((unsigned long long)__double_as_longlong(pi.w));
    auto body = [&](int, unsigned long long mj, double dx, double dy, double dz, double d2, int, int,
                    int) { lj_pair(a.lj, ti, meta_species(mj), dx, dy, dz, d2, e, fx, fy, fz); };
    walk_dispatch<GEO>(a, i, pi, body);
    const int o = meta_orig((unsigned long long)__double_as_longlong(pi.w));
    F[3 * (size_t)o] = fx;
    F[3 * (size_t)o + 1] = fy;
    F[3 * (size_t)o + 2] = fz;
    eatom[i] = e;  // per-atom energies in sorted order: the fold has one order whatever path wrote them
  }
}

// ---- CuH2 pass 1 + 2: density, then embedding energy and derivative ---------------------------
// (rgpot_cuh2.f90:436-451; the embedding is per-atom work on a finished rho_i, so it rides at
// the tail of the density kernel instead of paying a launch and an 8 B/atom round trip)
template <int M>
__global__ void __launch_bounds__(128)
k_eam_density(GatherArgs a, double* __restrict__ dfd_sorted, double* __restrict__ eemb_sorted,
              Status* __restrict__ st) {
  if (a.only_if_big_shift && !(a.st_in->flags & kFlagBigShift)) return;
  exp_table_init();
  const double rc2_phys = eam_par<M>().rc2_phys;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const double4 pi = a.spos[i];
    double rho = 0.0;
    bool coincident = false;
    auto body = [&](int, unsigned long long mj, double, double, double, double d2, int, int, int) {
      if (d2 == 0.0) coincident = true;  // dist <= 0 guard, rgpot_cuh2.f90:422-428
      if (!(d2 < rc2_phys)) return;      // :522 -- sqrt(d2) can round up onto the cutoff
      rho += eam_density_entry<M>(meta_species(mj), d2);
    };
    walk_dispatch<GEO_IMAGES>(a, i, pi, body);
    if (coincident) atomicOr(&st->flags, kFlagCoincident);
    double emb, demb;
    eam_embed<M>(meta_species((unsigned long long)__double_as_longlong(pi.w)), rho, emb, demb);
    dfd_sorted[i] = demb;
    eemb_sorted[i] = emb;
  }
}

// ---- CuH2 pass 3: pair energy + whole force on each atom (rgpot_cuh2.f90:454-457) -----------
template <int M>
__global__ void __launch_bounds__(128)
k_eam_force(GatherArgs a, const double* __restrict__ dfd_sorted,
            const double* __restrict__ eemb_sorted, double* __restrict__ F,
            double* __restrict__ eatom) {
  if (a.only_if_big_shift && !(a.st_in->flags & kFlagBigShift)) return;
  exp_table_init();
  const double rc2_phys = eam_par<M>().rc2_phys;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    const double4 pi = a.spos[i];
    const unsigned long long mi = (unsigned long long)__double_as_longlong(pi.w);
    const int zi = meta_species(mi);
    const double dfd_i = dfd_sorted[i];
    auto body = [&](int j, unsigned long long mj, double dx, double dy, double dz, double d2, int,
                    int, int) {
      if (!(d2 < rc2_phys)) return;
      eam_force_entry<M>(zi, meta_species(mj), dfd_i, dfd_sorted[j], dx, dy, dz, d2, e, fx, fy, fz);
    };
    walk_dispatch<GEO_IMAGES>(a, i, pi, body);
    e += eemb_sorted[i];
    const int o = meta_orig(mi);
    F[3 * (size_t)o] = fx;
    F[3 * (size_t)o + 1] = fy;
    F[3 * (size_t)o + 2] = fz;
    eatom[i] = e;
  }
}

// ---- Fe-He (rgpot_fehe.f90:573-620): the three passes on the gather walk ---------------------------
// First correct CUDA path of this potential: thread per atom, all system sizes (the brick and
// batch kernels do not carry its two density channels yet).
static __global__ void __launch_bounds__(128)
k_fehe_density(GatherArgs a, double* __restrict__ dfd_sorted, double* __restrict__ dfs_sorted,
               double* __restrict__ eemb_sorted) {
  if (a.only_if_big_shift && !(a.st_in->flags & kFlagBigShift)) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const double4 pi = a.spos[i];
    const int zi = meta_species((unsigned long long)__double_as_longlong(pi.w));
    double rho_d = 0.0, rho_s = 0.0;
    auto body = [&](int, unsigned long long mj, double, double, double, double d2, int, int, int) {
      double rho, drho;
      const int ch = fehe_density(zi, meta_species(mj), sqrt(d2), rho, drho);  // atom_densities :480-500
      if (ch == 0) rho_d += rho;
      if (ch == 1) rho_s += rho;
    };
    walk_dispatch<GEO_IMAGES>(a, i, pi, body);
    double e, dfd, dfs;
    fehe_embed(zi, rho_d, rho_s, e, dfd, dfs);
    dfd_sorted[i] = dfd;
    dfs_sorted[i] = dfs;
    eemb_sorted[i] = e;
  }
}

static __global__ void __launch_bounds__(128)
k_fehe_force(GatherArgs a, const double* __restrict__ dfd_sorted, const double* __restrict__ dfs_sorted,
             const double* __restrict__ eemb_sorted, double* __restrict__ F, double* __restrict__ eatom) {
  if (a.only_if_big_shift && !(a.st_in->flags & kFlagBigShift)) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    const double4 pi = a.spos[i];
    const unsigned long long mi = (unsigned long long)__double_as_longlong(pi.w);
    const int zi = meta_species(mi);
    const double dfd_i = dfd_sorted[i], dfs_i = dfs_sorted[i];
    auto body = [&](int j, unsigned long long mj, double dx, double dy, double dz, double d2, int, int, int) {
      const int zj = meta_species(mj);
      const double r = sqrt(d2);  // table%dist
      double v, dv, rho, drho;
      fehe_pair(zi, zj, r, v, dv);
      const int ch = fehe_density(zi, zj, r, rho, drho);
      e += 0.5 * v;  // atom_contribution :509-538
      double radial = dv;
      if (ch == 0) radial += (dfd_i + dfd_sorted[j]) * drho;
      if (ch == 1) radial += (dfs_i + dfs_sorted[j]) * drho;
      fx += radial * (dx / r);
      fy += radial * (dy / r);
      fz += radial * (dz / r);
    };
    walk_dispatch<GEO_IMAGES>(a, i, pi, body);
    e += eemb_sorted[i];
    const int o = meta_orig(mi);
    F[3 * (size_t)o] = fx;
    F[3 * (size_t)o + 1] = fy;
    F[3 * (size_t)o + 2] = fz;
    eatom[i] = e;
  }
}

// ---- neighbor-table emission (pair-set parity, list consumers) ----------------------------------
template <int GEO>
__global__ void k_emit_pairs(GatherArgs a, EmitArgs em) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double4 pi = a.spos[i];
  const int oi = meta_orig((unsigned long long)__double_as_longlong(pi.w));
  auto body = [&](int, unsigned long long mj, double vx, double vy, double vz, double d2, int S0, int S1,
                  int S2) {
    const int oj = meta_orig(mj);
    if (GEO != GEO_IMAGES && oj < oi) return;  // LJ: unordered pairs, i < j
    if (GEO == GEO_IMAGES && em.half) {
      // vesin's half list: first <= second, and one of the two images of a self pair
      if (oi > oj) return;
      if (oi == oj) {
        const int sum = S0 + S1 + S2;
        if (sum < 0 || (sum == 0 && (S2 < 0 || (S2 == 0 && S1 < 0)))) return;
      }
    }
    const unsigned long long k = atomicAdd(em.count, 1ull);
    if ((long)k < em.cap) {
      em.pairs[2 * k] = oi;
      em.pairs[2 * k + 1] = oj;
      em.shifts[3 * k] = GEO == GEO_IMAGES ? S0 : 0;
      em.shifts[3 * k + 1] = GEO == GEO_IMAGES ? S1 : 0;
      em.shifts[3 * k + 2] = GEO == GEO_IMAGES ? S2 : 0;
      if (em.vecs) {
        em.vecs[3 * k] = vx;
        em.vecs[3 * k + 1] = vy;
        em.vecs[3 * k + 2] = vz;
      }
      if (em.dists) em.dists[k] = sqrt(d2);
    }
  };
  walk_dispatch<GEO>(a, i, pi, body);
}

}  // namespace rb
