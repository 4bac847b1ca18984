"""second_opinion.py -- TEST INFRASTRUCTURE ONLY: an independent evaluation of the two EAM potentials whose
C oracles (oracle/eam_al_oracle.c, oracle/fehe_oracle.c) are pinned by the reference only loosely
(CppCore/tests/FortranPotsTest.cc:135-155: E to 1e-4, max |F| to 1e-3; the Fortran cannot be compiled in
this image).  Written a second time, by formula, from

    CppCore/rgpot/fortran/rgpot_eam_al.f90   (parameters :44-76, shapes :103-122, density :124-162,
                                              pair :164-195, embedding :197-226, passes :228-309)
    CppCore/rgpot/fortran/rgpot_fehe.f90     (parameters :63-156, pair terms :180-304, densities
                                              :311-355, embeddings :363-413, passes :417-538, driver :573-620)

in 40-digit arithmetic (mpmath), with the neighbor table replaced by a brute-force enumeration of
periodic images (every j /= i and every cell shift with |r_j - r_i + S.H|^2 < rcut^2; the table drops
i == j entries, rgpot_neighbors.f90:126-134).  It shares no code with the C restatements: a formula
misread in one of them shows up as a disagreement far above rounding.
"""
from __future__ import annotations

import itertools

import mpmath as mp
import numpy as np

mp.mp.dps = 40


def _mpf(x):
    return mp.mpf(float(x))  # exact: a double is a dyadic rational


def _images(pos, box, rcut):
    """[(i, j, vec, r)] for every j != i and every image inside rcut (brute force)"""
    n = len(pos)
    H = [[_mpf(v) for v in row] for row in np.asarray(box, dtype=float)]
    P = [[_mpf(v) for v in p] for p in np.asarray(pos, dtype=float)]
    Hn = np.asarray(box, dtype=float)
    # how many images per direction can matter: rcut over the face distance, plus atoms outside the cell
    vol = abs(np.linalg.det(Hn))
    face = [vol / np.linalg.norm(np.cross(Hn[(k + 1) % 3], Hn[(k + 2) % 3])) for k in range(3)]
    frac = np.asarray(pos, dtype=float) @ np.linalg.inv(Hn)
    spread = np.ceil(frac.max(axis=0) - frac.min(axis=0)).astype(int)
    M = [int(np.ceil(rcut / face[k])) + int(spread[k]) + 1 for k in range(3)]
    rc2 = _mpf(rcut) * _mpf(rcut)
    out = []
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            base = [P[j][k] - P[i][k] for k in range(3)]
            for S in itertools.product(*[range(-M[k], M[k] + 1) for k in range(3)]):
                vec = [base[k] + S[0] * H[0][k] + S[1] * H[1][k] + S[2] * H[2][k] for k in range(3)]
                d2 = vec[0] ** 2 + vec[1] ** 2 + vec[2] ** 2
                if d2 < rc2:
                    out.append((i, j, vec, mp.sqrt(d2)))
    return out


# ---------------------------------------------------------------------------------------- EAM-Al
class _Al:
    g = _mpf(-0.60647886749203)
    A, a = _mpf(2294.3609145535), _mpf(3.0205380362464)
    B, b = _mpf(-192.06894637533), _mpf(1.5102690181232)
    scale = _mpf(0.87928364657088)
    ba, bb, w = _mpf(3.3068828537934), _mpf(6.6137657075868), _mpf(512.0)
    eta = 6
    c = [_mpf(v) for v in (4.4572157051836, 193.21775368064, -1173.6502684704, 4203.3500116196,
                           -8785.1680280827, 10632.994102532, -6921.0410455328, 1875.2698365752)]
    rcut = _mpf(7.1)
    trunc = mp.sqrt(_mpf(0.9999)) * _mpf(7.1)


def _al_pair_shape(r):
    return _Al.A * mp.exp(-_Al.a * r) + _Al.B * mp.exp(-_Al.b * r)


def _al_density_shape(r):
    return r ** _Al.eta * (mp.exp(-_Al.ba * r) + _Al.w * mp.exp(-_Al.bb * r))


def _al_density(r):
    if r >= _Al.trunc:
        return mp.mpf(0)
    return _Al.scale * (_al_density_shape(r) - _al_density_shape(_Al.rcut))


def _al_density_deriv(r):
    if r >= _Al.trunc:
        return mp.mpf(0)
    da, db = mp.exp(-_Al.ba * r), _Al.w * mp.exp(-_Al.bb * r)
    return _Al.scale * (_Al.eta * r ** (_Al.eta - 1) * (da + db) - r ** _Al.eta * (_Al.ba * da + _Al.bb * db))


def _al_pair_energy(r):
    if r >= _Al.trunc:
        return mp.mpf(0)
    return _al_pair_shape(r) - _al_pair_shape(_Al.rcut) - 2 * _Al.g * _al_density(r)


def _al_pair_deriv(r):
    if r >= _Al.trunc:
        return mp.mpf(0)
    return -(_Al.A * _Al.a * mp.exp(-_Al.a * r) + _Al.B * _Al.b * mp.exp(-_Al.b * r)) - 2 * _Al.g * _al_density_deriv(r)


def _al_embedding(rho):
    return _Al.g * rho + sum(ck * rho ** (k + 1) for k, ck in enumerate(_Al.c))


def _al_embedding_deriv(rho):
    return _Al.g + _Al.c[0] + sum(ck * (k + 1) * rho ** k for k, ck in enumerate(_Al.c) if k >= 1)


def eam_al(pos, box):
    """(energy, forces) of rgpot_eam_al.f90 in 40-digit arithmetic"""
    n = len(pos)
    table = _images(pos, box, 7.1)
    rho = [mp.mpf(0)] * n
    for i, j, vec, r in table:
        rho[i] += _al_density(r)
    dembed = [_al_embedding_deriv(x) for x in rho]
    e = sum(_al_embedding(x) for x in rho)
    F = [[mp.mpf(0)] * 3 for _ in range(n)]
    for i, j, vec, r in table:
        e += mp.mpf("0.5") * _al_pair_energy(r)
        dedr = _al_pair_deriv(r) + (dembed[i] + dembed[j]) * _al_density_deriv(r)
        for k in range(3):
            F[i][k] += dedr * vec[k] / r
    return float(e), np.array([[float(v) for v in row] for row in F])


# ----------------------------------------------------------------------------------------- Fe-He
class _FH:
    rcut_fe_pair, rcut_fe_rho = _mpf(5.3), _mpf(4.2)
    rcut_fehe_pair, rcut_fehe_rho, rcut_he_pair = _mpf(3.9028), _mpf(4.10), _mpf(5.4)
    zbl_scale = _mpf(9.7342365892908e+03)
    zbl_w = [_mpf(v) for v in (0.18180, 0.50990, 0.28020, 0.02817)]
    zbl_d = [_mpf(v) for v in (2.8616724320005e+01, 8.4267310396064e+00, 3.6030244464156e+00, 1.8028536321603e+00)]
    bridge = [_mpf(v) for v in (7.4122709384068e+00, -6.4180690713367e-01, -2.6043547961722e+00, 6.2625393931230e-01)]
    bridge_lo, bridge_hi = _mpf(1.0), _mpf(2.05)
    pk = [_mpf(v) for v in (2.2, 2.3, 2.4, 2.5, 2.6, 2.7, 2.8, 3.0, 3.3, 3.7, 4.2, 4.7, 5.3)]
    pc = [_mpf(v) for v in (-2.7444805994228e+01, 1.5738054058489e+01, 2.2077118733936e+00, -2.4989799053251e+00,
                            4.2099676494795e+00, -7.7361294129713e-01, 8.0656414937789e-01, -2.3194358924605e+00,
                            2.6577406128280e+00, -1.0260416933564e+00, 3.5018615891957e-01, -5.8531821042271e-02,
                            -3.0458824556234e-03)]
    rk = [_mpf(v) for v in (2.4, 3.2, 4.2)]
    rcf = [_mpf(v) for v in (1.1686859407970e+01, -1.4710740098830e-02, 4.7193527075943e-01)]
    emb_q, emb_t = _mpf(-6.7314115586063e-04), _mpf(7.6514905604792e-08)
    hk = [_mpf(v) for v in (1.5440, 1.6155, 1.6896, 1.8017, 2.0482, 2.3816, 3.5067, 3.9028)]
    hc = [_mpf(v) for v in (559.804426025391, -45.916354995728, 35.550312671661, 164.319865173340,
                            -1.727464405060, 0.106771826237, 0.073715269849, 0.038235287677)]
    s_amp = _mpf(20.0)
    # the Fortran evaluates this quotient in double precision when the derived type is initialised
    s_decay = _mpf(1.5312426703208 / 0.529177210818181818)
    s_sqrt, s_q, s_t = _mpf(0.220813420485), _mpf(1.367508764130), _mpf(3.382256025271)
    he_rm, he_c6, he_c8, he_c10 = _mpf(2.9683), _mpf(1.35186623), _mpf(0.41495143), _mpf(0.17151143)
    he_aa, he_alpha, he_beta, he_damp = _mpf(186924.404), _mpf(10.5717543), _mpf(-2.07758779), _mpf(1.438)
    he_eps = _mpf(10.956 * 1.380658e-23 / 1.602177e-19)  # (a double-precision constant expression upstream)
    floor = _mpf(1.0e-35)


def _knots(knots, coeff, r, power):
    return sum(c * max(k - r, mp.mpf(0)) ** power for k, c in zip(knots, coeff))


def _fh_pair(zi, zj, r):
    """(V, dV/dr); species 0 = Fe, 1 = He"""
    P = _FH
    if zi == 0 and zj == 0:
        if r > P.rcut_fe_pair:
            return mp.mpf(0), mp.mpf(0)
        if r < P.bridge_lo:
            screen = [w * mp.exp(-d * r) for w, d in zip(P.zbl_w, P.zbl_d)]
            return (P.zbl_scale / r * sum(screen),
                    -P.zbl_scale / (r * r) * sum(screen) - P.zbl_scale / r * sum(d * s for d, s in zip(P.zbl_d, screen)))
        if r < P.bridge_hi:
            expo = mp.exp(P.bridge[0] + P.bridge[1] * r + P.bridge[2] * r * r + P.bridge[3] * r * r * r)
            return expo, expo * (P.bridge[1] + 2 * P.bridge[2] * r + 3 * P.bridge[3] * r * r)
        return _knots(P.pk, P.pc, r, 3), -3 * _knots(P.pk, P.pc, r, 2)
    if zi == 1 and zj == 1:
        if r > P.rcut_he_pair:
            return mp.mpf(0), mp.mpf(0)
        x = r / P.he_rm
        disp = P.he_c6 / x ** 6 + P.he_c8 / x ** 8 + P.he_c10 / x ** 10
        ddisp = -(6 * P.he_c6 / x ** 7 + 8 * P.he_c8 / x ** 9 + 10 * P.he_c10 / x ** 11)
        core = P.he_aa * mp.exp(-P.he_alpha * x + P.he_beta * x * x)
        damp, ddamp = mp.mpf(1), mp.mpf(0)
        if x < P.he_damp:
            damp = mp.exp(-(P.he_damp / x - 1) ** 2)
            ddamp = damp * 2 * P.he_damp * (P.he_damp / x - 1) / (x * x)
        return (P.he_eps * (core - damp * disp),
                P.he_eps * (core * (-P.he_alpha + 2 * P.he_beta * x) - ddamp * disp - damp * ddisp) / P.he_rm)
    if r > P.rcut_fehe_pair:
        return mp.mpf(0), mp.mpf(0)
    return _knots(P.hk, P.hc, r, 3), -3 * _knots(P.hk, P.hc, r, 2)


def _fh_density(zi, zj, r):
    """(rho_d, rho_s, drho_d, drho_s) a pair feeds into its channel"""
    P = _FH
    z = mp.mpf(0)
    if zi == 0 and zj == 0:
        if r > P.rcut_fe_rho:
            return z, z, z, z
        return _knots(P.rk, P.rcf, r, 3), z, -3 * _knots(P.rk, P.rcf, r, 2), z
    if zi != zj:
        if r > P.rcut_fehe_rho:
            return z, z, z, z
        ex = mp.exp(-2 * P.s_decay * r)
        return z, P.s_amp * r ** 3 * ex, z, P.s_amp * r ** 2 * ex * (3 - 2 * P.s_decay * r)
    return z, z, z, z


def fehe(pos, Z, box):
    """(energy, forces) of rgpot_fehe.f90 in 40-digit arithmetic; Z: atomic numbers 26 / 2"""
    P = _FH
    n = len(pos)
    sp = [0 if int(z) == 26 else 1 for z in Z]
    table = _images(pos, box, 5.4)
    rho_d, rho_s = [mp.mpf(0)] * n, [mp.mpf(0)] * n
    for i, j, vec, r in table:
        d, s, _, _ = _fh_density(sp[i], sp[j], r)
        rho_d[i] += d
        rho_s[i] += s
    e = mp.mpf(0)
    dfd, dfs = [mp.mpf(0)] * n, [mp.mpf(0)] * n
    for i in range(n):
        if sp[i] == 0 and not rho_d[i] < P.floor:
            e += -mp.sqrt(rho_d[i]) + P.emb_q * rho_d[i] ** 2 + P.emb_t * rho_d[i] ** 4
            dfd[i] = -mp.mpf("0.5") / mp.sqrt(rho_d[i]) + 2 * P.emb_q * rho_d[i] + 4 * P.emb_t * rho_d[i] ** 3
        if not rho_s[i] < P.floor:
            e += P.s_sqrt * mp.sqrt(rho_s[i]) + P.s_q * rho_s[i] ** 2 + P.s_t * rho_s[i] ** 4
            dfs[i] = mp.mpf("0.5") * P.s_sqrt / mp.sqrt(rho_s[i]) + 2 * P.s_q * rho_s[i] + 4 * P.s_t * rho_s[i] ** 3
    F = [[mp.mpf(0)] * 3 for _ in range(n)]
    for i, j, vec, r in table:
        v, dv = _fh_pair(sp[i], sp[j], r)
        _, _, dd, ds = _fh_density(sp[i], sp[j], r)
        e += mp.mpf("0.5") * v
        radial = dv + (dfd[i] + dfd[j]) * dd + (dfs[i] + dfs[j]) * ds
        for k in range(3):
            F[i][k] += radial * vec[k] / r
    return float(e), np.array([[float(v) for v in row] for row in F])
