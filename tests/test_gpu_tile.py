"""The large-system tile kernels (rgpot_b200/csrc/tile.cu) against the CPU oracle, and against themselves:
an atom's result must not depend on how bricks, staged capacities or window chunks were sized (groups
are aligned to the cell grid, and every summation order is a function of the cell list alone), nor on
whether the in-kernel slow path evaluated it.

Tolerances (BASELINE.json north_star): energies 1e-10 relative, forces 1e-10 relative / 1e-9 absolute;
pair sets are checked through the pair counts / energies of adversarial geometries below.
"""
import numpy as np
import pytest

import rgpot_b200
from rgpot_b200 import workloads

pytestmark = pytest.mark.gpu

E_RTOL = 1e-10
F_RTOL, F_ATOL = 1e-10, 1e-9


def assert_energy(e, ref, what=None):
    assert abs(e - ref) <= E_RTOL * abs(ref) + 1e-12, (e, ref, what)


def assert_forces(f, ref, what=None):
    err = np.abs(f - ref)
    assert (err <= F_ATOL + F_RTOL * np.abs(ref)).all(), (float(err.max()), what)


def fluid(n, rho, seed, shift=0):
    """jittered simple-cubic fluid; shift = 1 moves atoms by whole box lengths in {-1, 0, 1}: input that
    is not wrapped into the cell but stays within one cell vector (the tile kernels' own case)"""
    rng = np.random.default_rng(seed)
    m = int(np.ceil(n ** (1 / 3)))
    L = (n / rho) ** (1 / 3)
    g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] + 0.5) * (L / m)
    pos = np.mod(g + rng.normal(0, 0.08, g.shape), L)
    if shift:
        pos = pos + rng.integers(-shift, shift + 1, size=g.shape) * L
    return pos, np.diag([L, L, L])


TILE_KEYS = ("fine", "tile_group", "tile_brick", "tile_cap", "tile_wcap")
TILE_VARIANTS = [
    {},                                         # automatic
    {"tile_cap": 200},                          # staged capacity far too small: boxes are cut down to one group
    {"tile_wcap": 64},                          # windows searched in many chunks
    {"tile_brick": 884},                        # large bricks, two rounds of groups per warp
    {"tile_brick": 222, "tile_cap": 640},
    {"tile_group": 111},                        # one cell per group
    {"tile_group": 111, "tile_cap": 120, "tile_wcap": 64},
    {"tile_group": 321, "tile_brick": 642},
    {"fine": 1},                                # cutoff-sized cells, 27-cell stencil
    {"fine": 1, "tile_group": 111, "tile_wcap": 128},
]


def _with_options(pot, opts):
    pot.set_option("fine_min", 0)
    pot.set_option("tile", 1)
    for key in TILE_KEYS:
        pot.set_option(key, opts.get(key, 0))


def _key(opts):  # what the summation order may depend on: the cell grid and the group shape
    return (opts.get("fine", 0), opts.get("tile_group", 0))


@pytest.mark.parametrize("shift", [0, 1])
def test_tile_lj_variants_vs_oracle(oracle, shift):
    n = 4000
    pos, box = fluid(n, 0.8442, 77 + shift, shift)
    z = np.ones(n, np.int32)
    e_o, f_o, _ = oracle.lj_force(pos, box, 1.0, 2.5, 1.0)
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
    pot.set_path(1)
    first = {}
    for opts in TILE_VARIANTS:
        _with_options(pot, opts)
        for rep in range(2):  # the second call is planned from the halo populations the first one met
            e, f, _ = pot(pos, z, box)
            assert pot.stat("tile_used") == 1 and pot.stat("max_halo") > 0, opts  # the tile kernels did the work
            assert_energy(e, e_o, opts)
            assert_forces(f, f_o, opts)
            if _key(opts) in first:
                assert np.array_equal(f, first[_key(opts)][0]) and e == first[_key(opts)][1], (opts, rep)
            else:
                first[_key(opts)] = (f, e)


def test_tile_cuh2_variants_vs_oracle(oracle):
    pos, Z, box = workloads.cu_slab_triclinic(cells=8, n_h2_side=3)  # 2066 atoms, sheared cell
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    pot = rgpot_b200.CuH2Pot()
    pot.set_path(1)
    first = {}
    for opts in TILE_VARIANTS:
        for handoff in (1, 0):
            _with_options(pot, opts)
            pot.set_option("handoff", handoff)
            e, f, _ = pot(pos, Z, box)
            e, f, _ = pot(pos, Z, box)  # (the second call sizes its windows from what the first one met)
            assert pot.stat("tile_used") == 1 and pot.stat("max_halo") > 0, opts
            assert_energy(e, e_o, (opts, handoff))
            assert_forces(f, f_o, (opts, handoff))
            if handoff and opts.get("tile_wcap", 0) == 0 and opts.get("tile_cap", 0) == 0:
                assert pot.stat("pool_used") > 0  # the force pass reused the density pass's masks
            if _key(opts) in first:
                assert np.array_equal(f, first[_key(opts)][0]) and e == first[_key(opts)][1], (opts, handoff)
            else:
                first[_key(opts)] = (f, e)


def test_tile_unwrapped_triclinic_slab(oracle):
    """atoms up to one cell vector outside a sheared cell: wrap counts folded into the image codes"""
    pos, Z, box = workloads.cu_slab_triclinic(cells=8, n_h2_side=3)
    rng = np.random.default_rng(3)
    pos = pos + rng.integers(-1, 2, size=(len(pos), 3)).astype(float) @ box
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    pot = rgpot_b200.CuH2Pot()
    pot.set_path(1)
    _with_options(pot, {})
    e, f, _ = pot(pos, Z, box)
    assert pot.stat("tile_used") == 1
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    _with_options(pot, {"tile_cap": 300, "tile_group": 211})
    e2, f2, _ = pot(pos, Z, box)
    assert_energy(e2, e_o)
    assert_forces(f2, f_o)


def test_tile_other_potentials(oracle):
    n = 3000
    pos, box = fluid(n, 0.041, 5)  # 2.9 A lattice: the 9.5 A cutoff reaches several shells
    z = np.ones(n, np.int32)
    pot = rgpot_b200.MorsePot()
    pot.set_path(1)
    _with_options(pot, {})
    e, f, _ = pot(pos, z, box)
    assert pot.stat("tile_used") == 1
    e_o, f_o, _ = oracle.morse_force(pos, box)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    rng = np.random.default_rng(11)
    zz = rng.choice([1, 6, 14, 29], size=n).astype(np.int32)
    cl = rng.random((n, 3)) * (n / 0.09) ** (1 / 3)
    zb = rgpot_b200.ZBLPot()
    zb.set_path(1)
    _with_options(zb, {})
    e, f, _ = zb(cl, zz, np.eye(3))
    e_o, f_o, _ = oracle.zbl_force(cl, zz)
    assert_energy(e, e_o)
    assert (np.abs(f - f_o) <= 1e-9 + 1e-10 * np.abs(f_o)).all()
    # EAM-Al and Fe-He crystals
    a = 4.05
    g = np.arange(7, dtype=float)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    al = a * (corner[:, None, :] + workloads._FCC_BASIS[None, :, :]).reshape(-1, 3)
    al = al + np.random.default_rng(2).normal(0, 0.05, al.shape)
    alp = rgpot_b200.EAMAlPot()
    alp.set_path(1)
    _with_options(alp, {})
    boxal = np.eye(3) * 7 * a
    e, f, _ = alp(np.mod(al, 7 * a), None, boxal)
    assert alp.stat("tile_used") == 1
    e_o, f_o = oracle.eam_al_force(np.mod(al, 7 * a), boxal)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    a = 2.87
    g = np.arange(9, dtype=float)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    fe = a * np.concatenate([corner, corner + 0.5]) + np.random.default_rng(8).normal(0, 0.04, (2 * 729, 3))
    Zf = np.full(len(fe), 26, np.int32)
    Zf[::41] = 2
    fh = rgpot_b200.FeHePot()
    fh.set_path(1)
    _with_options(fh, {})
    boxfe = np.eye(3) * 9 * a
    e, f, _ = fh(np.mod(fe, 9 * a), Zf, boxfe)
    assert fh.stat("tile_used") == 1
    e_o, f_o = oracle.fehe_force(np.mod(fe, 9 * a), Zf, boxfe)
    assert_energy(e, e_o)
    assert_forces(f, f_o)


def test_tile_dense_group_takes_the_slow_path_with_the_same_bits(oracle):
    """A group whose window exceeds the staged capacity is walked straight from global memory inside
    the kernel; the atom gets the SAME bits as on the fast path (same candidate order, same four
    partial sums, same fold)."""
    rng = np.random.default_rng(4)
    n = 6000
    pos, box = fluid(n, 0.8442, 9)
    blob = rng.random((600, 3)) * 1.2 + 4.0  # 600 atoms inside one small region (tiny well: forces stay finite)
    pos = np.vstack([pos, blob])
    z = np.ones(len(pos), np.int32)
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0e-6, 2.5, 0.3))
    pot.set_path(1)
    e_o, f_o, _ = oracle.lj_force(pos, box, 1.0e-6, 2.5, 0.3)
    _with_options(pot, {"tile_cap": 4000})  # everything fits: no slow path
    e_fast, f_fast, _ = pot(pos, z, box)
    assert pot.stat("tile_used") == 1 and pot.stat("slow_atoms") == 0
    assert_energy(e_fast, e_o)
    assert_forces(f_fast, f_o)
    _with_options(pot, {"tile_cap": 400})
    e, f, _ = pot(pos, z, box)
    assert pot.stat("tile_used") == 1 and pot.stat("slow_atoms") > 0
    assert np.array_equal(f, f_fast) and e == e_fast


@pytest.mark.parametrize("ulps", [0, 1, 2, 8, -1, -2, -8])
def test_tile_prefilter_keeps_pairs_at_the_cutoff(oracle, ulps):
    """Adversarial pairs at r^2 = rc^2 (1 +- k ulp) placed where the FP16 prefilter is weakest: atoms
    at group / brick corners and on opposite sides of the periodic boundary, inside an otherwise
    ordinary fluid.  A pair at or inside the cutoff must not be lost (the tile result must equal the
    oracle's, whose accept test is the reference's), a pair outside must not be kept."""
    rc = 2.5
    n = 4000
    pos, box = fluid(n, 0.8442, 123)
    L = box[0, 0]
    cells = int(np.floor(2 * L / (rc * (1 + 1e-6))))  # half-cutoff cells, like the engine's grid
    edge = L / cells
    rng = np.random.default_rng(5)
    extra = []
    # partners exactly rc apart along a random direction, first atom at a cell corner
    for trial in range(24):
        corner = np.array([rng.integers(0, cells), rng.integers(0, cells), rng.integers(0, cells)]) * edge
        p0 = corner + (1e-9 if trial % 2 else -1e-9)
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        r2_target = rc * rc
        for _ in range(abs(ulps)):
            r2_target = np.nextafter(r2_target, np.inf if ulps > 0 else 0.0)
        p1 = p0 + u * np.sqrt(r2_target)
        extra += [p0, p1]
    pos = np.vstack([pos, np.mod(np.array(extra), L)])
    # thin the fluid around the inserted atoms so forces stay moderate
    keep = np.ones(len(pos), bool)
    ins = pos[n:]
    for p in ins:
        d = pos[:n] - p
        d -= L * np.round(d / L)
        keep[:n] &= (d ** 2).sum(1) > 0.8
    pos = pos[keep]
    z = np.ones(len(pos), np.int32)
    e_o, f_o, npairs = oracle.lj_force(pos, box, 1.0, rc, 1.0)
    for opts in ({}, {"tile_group": 111}, {"tile_brick": 884}, {"fine": 1}):
        pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, rc, 1.0))
        pot.set_path(1)
        _with_options(pot, opts)
        e, f, _ = pot(pos, z, box)
        assert pot.stat("tile_used") == 1
        assert_energy(e, e_o, opts)
        assert_forces(f, f_o, opts)
        # (the shifted pair energy vanishes at the cutoff but the pair force does not -- 0.039 in reduced
        # units at rc = 2.5 -- so a lost or a spurious boundary pair fails the force comparison above)
