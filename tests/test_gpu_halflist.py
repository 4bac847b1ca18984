"""The half-list mode of the batch kernel (k_eam_small, `half` option, default on): every pair is evaluated
once by its owner and the partner's share arrives through 64-bit fixed-point accumulators in shared memory.
Checked here: agreement with the oracle and with the full-list mode (option half=0) on systems that exercise
the ownership rule (odd and even atom counts, one to a few hundred atoms), the hydrogen queue (more mixed
pairs than it holds), atoms outside the cell, EAM-Al; bit-reproducibility from call to call (integer sums do
not depend on the arrival order); and the exits: terms outside the fixed-point range and boxes with several
images per pair leave the mode for the general code and still match the oracle.

Reference: rgpot_cuh2.f90:506-587 (the two pair passes; `w` of :566-580 is symmetric in i <-> j, which is
what lets one evaluation serve both atoms), rgpot_neighbors.f90:129-134 (self images dropped).
"""
import numpy as np
import pytest

import rgpot_b200
from rgpot_b200 import workloads

pytestmark = pytest.mark.gpu

F_RTOL, F_ATOL = 1e-10, 1e-9


def _close(e, f, e_o, f_o, what=None):
    assert abs(e - e_o) <= 1e-10 * abs(e_o) + 1e-12, (e, e_o, what)
    err = np.abs(f - f_o)
    assert (err <= F_ATOL + F_RTOL * np.abs(f_o)).all(), (float(err.max()), what)


def _cu_block(n_cu, n_h, seed, box_pad=8.0):
    """n_cu copper atoms cut from a jittered fcc block plus n_h hydrogens dropped between them"""
    rng = np.random.default_rng(seed)
    cells = int(np.ceil((n_cu / 4) ** (1 / 3)))
    a = 3.615
    g = np.arange(cells, dtype=float)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    cu = a * (corner[:, None, :] + workloads._FCC_BASIS[None, :, :]).reshape(-1, 3)[:n_cu]
    cu = cu + rng.normal(0, 0.04, cu.shape)
    L = cells * a
    h = []
    while len(h) < n_h:  # octahedral-ish sites with jitter, at least 0.9 A from everything placed so far
        p = rng.random(3) * L
        pts = np.vstack([cu] + ([np.array(h)] if h else []))
        if np.min(np.linalg.norm(pts - p, axis=1)) > (1.45 if len(h) % 2 else 0.9):
            h.append(p)
    pos = np.vstack([cu, np.array(h).reshape(-1, 3)]) if n_h else cu
    Z = np.concatenate([np.full(len(cu), 29, np.int32), np.ones(n_h, np.int32)])
    box = np.diag([L + box_pad, L + box_pad + 1.5, L + box_pad + 3.0])
    return np.ascontiguousarray(pos), Z, box


@pytest.mark.parametrize("n_cu,n_h", [(1, 0), (2, 0), (3, 1), (7, 2), (32, 0), (63, 2), (64, 2), (107, 6), (216, 2), (255, 0),
                                      (256, 4), (300, 30), (490, 10), (512, 0)])
def test_half_lists_match_the_oracle_and_the_full_lists(oracle, n_cu, n_h):
    pos, Z, box = _cu_block(n_cu, n_h, 5 + n_cu)
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    half = rgpot_b200.CuH2Pot()
    full = rgpot_b200.CuH2Pot()
    full.set_option("half", 0)
    shifted = pos + np.random.default_rng(3).integers(-1, 2, size=pos.shape).astype(float) @ box
    e_s, f_s = oracle.cuh2_force(shifted, Z, box)
    out_h = half.forceBatch([(pos, Z, box), (shifted, Z, box), (pos, Z, box)])
    out_f = full.forceBatch([(pos, Z, box), (shifted, Z, box)])
    _close(out_h[0][0], out_h[0][1], e_o, f_o, "half")
    _close(out_h[1][0], out_h[1][1], e_s, f_s, "half, atoms outside the cell")
    _close(out_f[0][0], out_f[0][1], e_o, f_o, "full")
    # integer accumulation: the same bits whatever order the lanes arrived in
    assert out_h[0][0] == out_h[2][0] and np.array_equal(out_h[0][1], out_h[2][1])
    again = half.forceBatch([(pos, Z, box)])
    assert again[0][0] == out_h[0][0] and np.array_equal(again[0][1], out_h[0][1])
    # the two modes differ by the fixed-point rounding and the summation order only
    assert np.abs(out_h[0][1] - out_f[0][1]).max() <= 1e-10


def test_more_hydrogen_pairs_than_the_queue_holds(oracle):
    """120 Cu + 120 H: thousands of mixed pairs, the queue holds 512; the rest are evaluated in place"""
    pos, Z, box = _cu_block(120, 120, 77, box_pad=7.0)
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    pot = rgpot_b200.CuH2Pot()
    (e, f, _), = pot.forceBatch([(pos, Z, box)])
    _close(e, f, e_o, f_o)
    (e2, f2, _), = pot.forceBatch([(pos, Z, box)])
    assert e2 == e and np.array_equal(f2, f)


def test_terms_outside_the_fixed_point_range_take_the_general_path(oracle):
    """two coppers 0.35 A apart: |w| ~ 1e5 eV/A^2, far outside the 2^-40 accumulators' per-term range; the
    system is handed to the cell-list kernels (kFlagListOverflow) and the neighbours in the batch are not disturbed"""
    pos, Z, box = _cu_block(64, 2, 9)
    crushed = pos.copy()
    crushed[1] = crushed[0] + np.array([0.35, 0.0, 0.0])
    e_o, f_o = oracle.cuh2_force(crushed, Z, box)
    e_n, f_n = oracle.cuh2_force(pos, Z, box)
    assert np.abs(f_o).max() > 2.0e3  # (the case really is outside the range)
    pot = rgpot_b200.CuH2Pot()
    out = pot.forceBatch([(pos, Z, box), (crushed, Z, box), (pos, Z, box)])
    _close(out[1][0], out[1][1], e_o, f_o, "crushed")
    _close(out[0][0], out[0][1], e_n, f_n)
    assert out[0][0] == out[2][0] and np.array_equal(out[0][1], out[2][1])


def test_small_boxes_keep_the_full_lists(oracle):
    """a box narrower than two cutoffs has several images per pair: the mode is not used, results as before"""
    pos, Z, box = _cu_block(32, 2, 21, box_pad=0.4)
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    (e, f, _), = rgpot_b200.CuH2Pot().forceBatch([(pos, Z, box)])
    _close(e, f, e_o, f_o)


def test_eam_al_half_lists(oracle):
    rng = np.random.default_rng(4)
    a = 4.05
    g = np.arange(4, dtype=float)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    pos = a * (corner[:, None, :] + workloads._FCC_BASIS[None, :, :]).reshape(-1, 3) + rng.normal(0, 0.05, (256, 3))
    box = np.diag([4 * a, 4 * a, 4 * a + 12.0])
    Z = np.full(256, 13, np.int32)
    e_o, f_o = oracle.eam_al_force(pos, box)
    pot = rgpot_b200.EAMAlPot()
    (e, f, _), (e2, f2, _) = pot.forceBatch([(pos, Z, box), (pos, Z, box)])
    _close(e, f, e_o, f_o)
    assert e2 == e and np.array_equal(f2, f)
