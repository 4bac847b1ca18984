"""Adversarial parity tests for the prefilters: the FP16 search of the brick
kernels and the FP32 search of the batch kernel are supersets of the exact accept test by a derived
bound; here that bound is attacked instead of sampled.  Pairs are placed at r^2 = rc^2 (1 +- k ulp),
with one partner at a cell corner, across periodic faces, in strongly sheared cells, and at coordinates
hundreds of box lengths from the origin.  The reference decides such pairs with its own floating-point
expression (vesin_visit.hpp:97-104 `r2 <= rc*rc`; vesin-single-build.cpp:1199-1202 `d2 < rc*rc`); the
oracle restates it bit for bit, and a pair lost by a prefilter or kept wrongly moves a force by
>= 4e-3 (the pair force at the cutoff), a million times the tolerance.
"""
import numpy as np
import pytest

import rgpot_b200
from rgpot_b200 import workloads

pytestmark = pytest.mark.gpu

F_RTOL, F_ATOL = 1e-10, 1e-9


def assert_forces(f, ref, what=None):
    err = np.abs(f - ref)
    assert (err <= F_ATOL + F_RTOL * np.abs(ref)).all(), (float(err.max()), what)


def nudge(x, ulps):
    for _ in range(abs(ulps)):
        x = np.nextafter(x, np.inf if ulps > 0 else 0.0)
    return x


def lj_fluid_with_boundary_pairs(n, rc, ulps, seed):
    """a jittered simple-cubic fluid plus 24 pairs exactly (+- ulps) at the cutoff, first partner at a
    corner of the engine's half-cutoff cell grid, random direction (many cross a periodic face)"""
    rng = np.random.default_rng(seed)
    m = int(np.ceil(n ** (1 / 3)))
    L = (n / 0.8442) ** (1 / 3)
    g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] + 0.5) * (L / m)
    pos = np.mod(g + rng.normal(0, 0.08, g.shape), L)
    cells = int(np.floor(2 * L / (rc * (1 + 1e-6))))
    edge = L / cells
    extra = []
    trial = 0
    while len(extra) < 48:
        trial += 1
        corner = rng.integers(0, cells, size=3) * edge
        p0 = corner + (1e-9 if trial % 2 else -1e-9)
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        p1 = p0 + u * np.sqrt(nudge(rc * rc, ulps))
        ok = True
        for q in extra:  # no two inserted atoms on top of each other (periodic distance)
            for p in (p0, p1):
                d = np.mod(p - q + 0.5 * L, L) - 0.5 * L
                ok = ok and (d ** 2).sum() > 0.9
        if ok:
            extra += [p0, p1]
    ins = np.mod(np.array(extra), L)
    keep = np.ones(n, bool)
    for p in ins:  # thin the fluid around the inserted atoms so forces stay moderate
        d = pos - p
        d -= L * np.round(d / L)
        keep &= (d ** 2).sum(1) > 0.8
    return np.vstack([pos[keep], ins]), np.diag([L, L, L])


@pytest.mark.parametrize("ulps", [0, 1, 2, 8, -1, -2, -8])
def test_brick_lj_pairs_at_the_cutoff(oracle, ulps):
    rc = 2.5
    pos, box = lj_fluid_with_boundary_pairs(4000, rc, ulps, 131 + ulps)
    z = np.ones(len(pos), np.int32)
    e_o, f_o, _ = oracle.lj_force(pos, box, 1.0, rc, 1.0)
    for opts in ({"brick_tpa": 1}, {"brick_tpa": 2}, {"brick_tpa": 4}, {"fine": 1, "brick_tpa": 2}, {"brick": 222},
                 {"brick_list": 4}):
        pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, rc, 1.0))
        pot.set_path(1)
        pot.set_option("fine_min", 0)
        pot.set_option("tile", 0)
        for k, v in opts.items():
            pot.set_option(k, v)
        e, f, _ = pot(pos, z, box)
        assert pot.stat("max_halo") > 0 and pot.stat("tile_used") == 0  # the brick kernel did the work
        assert abs(e - e_o) <= 1e-10 * abs(e_o)
        assert_forces(f, f_o, opts)


def sheared_cu_with_boundary_pairs(cells, ulps, seed, shear_deg=60.0):
    """fcc Cu block in a cell sheared by `shear_deg` in the xy plane (lattice-commensurate would need a
    special angle: here the crystal is simply cut by the cell, atoms near the faces removed), plus
    Cu-Cu pairs at d2 = rc^2 (1 +- ulps) whose first partner sits on a cell corner / face"""
    a, rc = 3.615, 6.1
    rng = np.random.default_rng(seed)
    span = cells * a
    t = np.tan(np.radians(90.0 - shear_deg))  # 60 degrees between the a and b cell vectors
    box = np.array([[span, 0.0, 0.0], [span * t, span, 0.0], [0.3 * span * t, 0.2 * span, span + 12.0]])
    inv = np.linalg.inv(box)
    g = np.arange(-cells, 2 * cells, dtype=float)
    ix, iy, iz = np.meshgrid(g, g, np.arange(cells, dtype=float), indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    sites = a * (corner[:, None, :] + workloads._FCC_BASIS[None, :, :]).reshape(-1, 3) + 0.11
    frac = sites @ inv
    inside = ((frac > 0.03) & (frac < 0.97)).all(axis=1)  # no pairs closer than 1.7 A across a face
    pos = sites[inside] + rng.normal(0, 0.03, (int(inside.sum()), 3))
    extra = []
    for trial in range(16):
        f0 = rng.integers(0, 2, size=3) * 1.0 * (trial % 2) + rng.random(3) * (1 - trial % 2)
        f0[2] = rng.random() * 0.7 + 0.05
        p0 = f0 @ box + (1e-9 if trial % 3 else -1e-9)
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        extra += [p0, p0 + u * np.sqrt(nudge(rc * rc, ulps))]
    ins = np.array(extra)
    keep = np.ones(len(pos), bool)
    for p in ins:
        d = (pos - p) @ inv
        d -= np.round(d)
        keep &= ((d @ box) ** 2).sum(1) > 2.0 ** 2
    pos = np.vstack([pos[keep], ins])
    # also keep the inserted atoms apart from each other
    ok = np.ones(len(pos), bool)
    for i in range(len(pos) - len(ins), len(pos)):
        d = (pos[:i] - pos[i]) @ inv
        d -= np.round(d)
        r = np.sqrt(((d @ box) ** 2).sum(1))
        partner = i - 1 if (i - (len(pos) - len(ins))) % 2 else -1
        for j in np.nonzero(r < 1.9)[0]:
            if j != partner:
                ok[i] = False
    pos = pos[ok]
    return pos, np.full(len(pos), 29, np.int32), box


@pytest.mark.parametrize("ulps", [0, 1, 8, -1, -8])
def test_brick_eam_pairs_at_the_cutoff_in_a_sheared_cell(oracle, ulps):
    pos, Z, box = sheared_cu_with_boundary_pairs(7, ulps, 105 + ulps)
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    for opts in ({"brick_tpa": 2}, {"brick_tpa": 4}, {"fine": 1, "brick_tpa": 4}, {"handoff": 0, "brick_tpa": 4}):
        pot = rgpot_b200.CuH2Pot()
        pot.set_path(1)
        pot.set_option("fine_min", 0)
        pot.set_option("tile", 0)
        for k, v in opts.items():
            pot.set_option(k, v)
        e, f, _ = pot(pos, Z, box)
        assert pot.stat("max_halo") > 0
        assert abs(e - e_o) <= 1e-10 * abs(e_o), opts
        assert_forces(f, f_o, opts)
        # unwrapped input: whole cell vectors added to some atoms (wrap counts in the image codes)
        shift = np.random.default_rng(1).integers(-1, 2, size=(len(pos), 3)).astype(float) @ box
        e2, f2, _ = pot(pos + shift, Z, box)
        e_o2, f_o2 = oracle.cuh2_force(pos + shift, Z, box)
        assert abs(e2 - e_o2) <= 1e-10 * abs(e_o2), opts
        assert_forces(f2, f_o2, opts)


@pytest.mark.parametrize("cells,ulps", [(3, 0), (3, 2), (3, -2), (4, 1), (4, -1), (5, 0), (5, 8), (5, -8)])
def test_batch_kernel_pairs_at_the_cutoff(oracle, cells, ulps):
    """k_eam_small (FP32 search in fractional space, half lists in a batch, full lists for one system): 60-degree cell, pairs on the cutoff, systems sized
    for 4 / 2 / 1 lanes per atom; a batch of the system and of copies moved by whole cell vectors"""
    pos, Z, box = sheared_cu_with_boundary_pairs(cells, ulps, 40 + cells + ulps)
    if len(pos) > 500:
        pos, Z = pos[-500:], Z[-500:]  # (the inserted pairs are the last atoms)
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    pot = rgpot_b200.CuH2Pot()
    pot.set_path(2)  # all-pairs kernels only
    rng = np.random.default_rng(2)
    moved = pos + rng.integers(-1, 2, size=(len(pos), 3)).astype(float) @ box
    e_m, f_m = oracle.cuh2_force(moved, Z, box)
    out = pot.forceBatch([(pos, Z, box), (moved, Z, box), (pos, Z, box)])
    for (e, f, _), (er, fr) in zip(out, [(e_o, f_o), (e_m, f_m), (e_o, f_o)]):
        assert abs(e - er) <= 1e-10 * abs(er), (cells, ulps)
        assert_forces(f, fr, (cells, ulps))
    assert out[0][0] == out[2][0] and np.array_equal(out[0][1], out[2][1])
    # forced all-pairs path: the one-launch single-system call runs the same kernel with full lists (a batch
    # evaluates every pair once and hands the partner its share in fixed point: another summation order)
    e1, f1, _ = pot(pos, Z, box)
    assert abs(e1 - e_o) <= 1e-10 * abs(e_o), (cells, ulps)
    assert_forces(f1, f_o, (cells, ulps))
    assert np.abs(f1 - out[0][1]).max() <= 1e-10
    pot.set_option("half", 0)
    full = pot.forceBatch([(pos, Z, box)])
    assert e1 == full[0][0] and np.array_equal(f1, full[0][1])  # full lists in both: the same bits
    pot.set_option("half", 1)
    # the automatic single-system path (several blocks per system, single.cuh): other summation order
    auto = rgpot_b200.CuH2Pot()
    for p_, (er, fr) in ((pos, (e_o, f_o)), (moved, (e_m, f_m))):
        e2, f2, _ = auto(p_, Z, box)
        assert abs(e2 - er) <= 1e-10 * abs(er), (cells, ulps)
        assert_forces(f2, fr, (cells, ulps))


@pytest.mark.parametrize("ulps", [0, 2, -2])
def test_far_displaced_input_takes_the_reference_fold(oracle, ulps):
    """coordinates ~400 box lengths from the origin (nine bits of the mantissa gone): the literal fold
    d - w * round(d / w) of the reference decides the boundary pairs, not a wrapped copy"""
    rc = 2.5
    pos, box = lj_fluid_with_boundary_pairs(4000, rc, ulps, 77 + ulps)
    L = box[0, 0]
    rng = np.random.default_rng(9)
    far = pos + rng.integers(-400, 401, size=pos.shape) * L
    z = np.ones(len(far), np.int32)
    e_o, f_o, _ = oracle.lj_force(far, box, 1.0, rc, 1.0)
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, rc, 1.0))
    pot.set_path(1)
    pot.set_option("fine_min", 0)
    e, f, _ = pot(far, z, box)
    assert abs(e - e_o) <= 1e-10 * abs(e_o)
    assert_forces(f, f_o)
