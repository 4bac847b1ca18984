"""Parity of the CUDA path against the CPU oracle and the committed golden fixtures.

Everything here calls the product through its C ABI (rgpot_b200 -> librgpot_b200.so) on a B200.
Tolerances (BASELINE.json north_star): neighbor pair sets bit-exact; FP64 energies within 1e-10
relative; forces within 1e-10 relative / 1e-9 eV/A absolute.
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
    assert f.shape == ref.shape
    err = np.abs(f - ref)
    assert (err <= F_ATOL + F_RTOL * np.abs(ref)).all(), (float(err.max()), what)


def canon(pairs, shifts):
    rec = np.column_stack([pairs.astype(np.int64), shifts.astype(np.int64)])
    if len(rec) == 0:
        return rec.reshape(0, 5)
    order = np.lexsort((rec[:, 4], rec[:, 3], rec[:, 2], rec[:, 1], rec[:, 0]))
    return rec[order]


@pytest.fixture(scope="module")
def lj():
    return rgpot_b200.LJPot()


@pytest.fixture(scope="module")
def ljc():
    return rgpot_b200.LJClusterPot()


@pytest.fixture(scope="module")
def cuh2():
    return rgpot_b200.CuH2Pot()


# ------------------------------------------------------------------------------------------- LJ
def test_lj_cluster_pins(ljc, golden):
    g = golden("lj13_cluster.json")
    pos, box = np.array(g["positions"]), np.array(g["box"]).reshape(3, 3)
    e, f, v = ljc(pos, np.ones(13, np.int32), box)
    assert e == pytest.approx(g["pin_energy"], rel=1e-9)  # LJClusterPotTest.cc:68
    assert np.sqrt((f ** 2).sum(1)).max() == pytest.approx(g["pin_max_force"], rel=1e-9)
    assert v == 0.0
    assert_energy(e, g["ref_cluster_energy"])
    assert_forces(f, np.array(g["ref_cluster_forces"]))
    assert not ljc.caps().periodic and ljc.caps().batched


def test_lj_periodic_equals_cluster_and_ignores_cell(lj, ljc, golden):
    g = golden("lj13_cluster.json")
    pos, box = np.array(g["positions"]), np.array(g["box"]).reshape(3, 3)
    z = np.ones(13, np.int32)
    e_c, f_c, _ = ljc(pos, z, box)
    e_p, f_p, _ = lj(pos, z, box)
    assert e_c == pytest.approx(e_p, rel=1e-12)  # LJClusterPotTest.cc:72-94
    assert np.abs(f_c - f_p).max() < 1e-12
    e_t, _, _ = ljc(pos, z, np.eye(3) * 5.0)  # :96-112
    assert e_t == pytest.approx(e_c, rel=1e-12)


@pytest.mark.parametrize("path", [1, 2])
def test_lj_reference_cases(golden, path):
    for name, c in golden("lj_reference_cases.json")["cases"].items():
        pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(c["u0"], c["cutoff"], c["psi"]))
        pot.set_path(path)
        pos, box = np.array(c["positions"]), np.array(c["box"]).reshape(3, 3)
        z = np.ones(len(pos), np.int32)
        e, f, _ = pot(pos, z, box)
        assert_energy(e, c["energy"])
        assert_forces(f, np.array(c["forces"]))
        pairs, shifts = pot.neighbors(pos, z, box, path=path)
        want = np.array(c["pairs"], dtype=np.int64).reshape(-1, 2)
        got = canon(pairs, shifts)[:, :2]
        assert np.array_equal(got, want[np.lexsort((want[:, 1], want[:, 0]))]), name


def test_lj_invariance_dimer(lj):
    # InvarianceTest.cc:25-89 (rc = 15 > L/2: nearest image only) and test_rgpot_pip_core.py:27-56
    box = np.eye(3) * 10.0
    pos = np.array([[1.0, 1.0, 1.0], [2.5, 1.0, 1.0]])
    z = np.ones(2, np.int32)
    e0, f0, _ = lj(pos, z, box)
    e1, _, _ = lj(pos + np.array([3.3, -2.1, 0.7]), z, box)
    assert e1 == pytest.approx(e0, abs=1e-12)
    rot = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    e2, _, _ = lj((pos - pos[0]) @ rot.T + pos[0], z, box)
    assert e2 == pytest.approx(e0, abs=1e-12)
    assert np.isfinite(e0) and np.abs(f0[0] + f0[1]).max() < 1e-9
    ef, ff, _ = rgpot_b200.evaluate_lj(pos, z, box)
    assert ef == pytest.approx(e0, abs=1e-12) and np.abs(ff - f0).max() < 1e-12
    n0 = lj.force_calls
    lj(pos, z, box)
    assert lj.force_calls == n0 + 1


def _random_fluid(n, rho, seed, wrapped):
    rng = np.random.default_rng(seed)
    m = int(np.ceil(n ** (1 / 3)))
    L = (n / rho) ** (1 / 3)
    g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] + 0.5) * (L / m)
    pos = g + rng.normal(0, 0.08, g.shape)
    if wrapped:
        pos = np.mod(pos, L)
    else:
        pos += rng.integers(-2, 3, size=g.shape) * L  # whole-box displacements: unwrapped input
    return pos, np.diag([L, L, L])


@pytest.mark.parametrize("n,wrapped,path", [(700, True, 1), (700, False, 1), (700, False, 2),
                                            (4000, True, 1), (4000, False, 1)])
def test_lj_fluid_vs_oracle(oracle, n, wrapped, path):
    pos, box = _random_fluid(n, 0.8442, n + int(wrapped), wrapped)
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
    pot.set_path(path)
    z = np.ones(n, np.int32)
    e, f, _ = pot(pos, z, box)
    e_o, f_o, _ = oracle.lj_force(pos, box, 1.0, 2.5, 1.0)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    pairs, shifts = pot.neighbors(pos, z, box, path=path)
    want = oracle.lj_pairs(pos, box, 2.5).astype(np.int64)
    assert np.array_equal(canon(pairs, shifts)[:, :2], want[np.lexsort((want[:, 1], want[:, 0]))])


def test_lj_cluster_large_cell_path(oracle):
    rng = np.random.default_rng(3)
    n = 3000
    pos = rng.random((n, 3)) * 22.0
    # relax overlaps a little so forces stay moderate
    pot = rgpot_b200.LJClusterPot(rgpot_b200.LJClusterConfig(1.0, 3.0, 1.0))
    pot.set_path(1)
    e, f, _ = pot(pos, np.ones(n, np.int32), np.eye(3) * 5.0)
    e_o, f_o, _ = oracle.lj_force(pos, np.eye(3) * 5.0, 1.0, 3.0, 1.0, cluster=True)
    assert abs(e - e_o) <= 1e-10 * abs(e_o)
    assert (np.abs(f - f_o) <= 1e-9 + 1e-10 * np.abs(f_o)).all()


def test_lj_trivial_inputs(lj, ljc):
    box = np.eye(3) * 10
    e, f, _ = lj(np.zeros((1, 3)), np.ones(1, np.int32), box)  # LJPot.cc:50-52
    assert e == 0.0 and np.array_equal(f, np.zeros((1, 3)))
    e, f, _ = lj(np.zeros((0, 3)), np.zeros(0, np.int32), box)
    assert e == 0.0 and f.shape == (0, 3)
    z = rgpot_b200.LJPot(rgpot_b200.LJConfig(cutoff=0.0))
    e, f, _ = z(np.array([[0, 0, 0], [1.0, 0, 0]]), np.ones(2, np.int32), box)
    assert e == 0.0 and not f.any()
    with pytest.raises(ValueError):
        lj(np.zeros((3, 2)), np.ones(3, np.int32), box)
    with pytest.raises(ValueError):
        lj(np.zeros((3, 3)), np.ones(2, np.int32), box)
    with pytest.raises(ValueError):
        lj(np.zeros((3, 3)), np.ones(3, np.int32), np.zeros(9))


# ----------------------------------------------------------------------------------------- CuH2
@pytest.mark.parametrize("path", [1, 2])
def test_cuh2_pin4(cuh2, golden, path):
    g = golden("cuh2_pin4.json")
    cuh2.set_path(path)
    e, f, v = cuh2(g["positions"], g["Z"], np.array(g["box"]).reshape(3, 3))
    cuh2.set_path(0)
    assert_energy(e, g["pin_energy"])  # CuH2PotTest.cc:24 (upstream tolerance 1e-6)
    assert_forces(f, np.array(g["pin_forces"]))
    assert v == 0.0


def test_cuh2_dimer(cuh2, golden):
    g = golden("cuh2_dimer.json")
    e, f, _ = cuh2(g["positions"], g["Z"], np.array(g["box"]).reshape(3, 3))
    assert e == pytest.approx(g["pin_energy"], rel=1e-6)  # tests/rpc_integ.py:76-78
    assert f[0, 0] == pytest.approx(g["pin_f0x"], rel=1e-6)
    assert f[1, 0] == pytest.approx(g["pin_f1x"], rel=1e-6)


@pytest.mark.parametrize("path", [1, 2])
def test_cuh2_slab218(cuh2, golden, path):
    g = golden("cuh2_slab218.json")
    pos, Z, box = np.array(g["positions"]), np.array(g["Z"], np.int32), np.array(g["box"]).reshape(3, 3)
    cuh2.set_path(path)
    e, f, _ = cuh2(pos, Z, box)
    pairs, shifts = cuh2.neighbors(pos, Z, box, path=path)
    cuh2.set_path(0)
    assert_energy(e, g["energy"])
    assert_forces(f, np.array(g["forces"]))
    assert len(pairs) == g["n_directed_pairs"]


@pytest.mark.parametrize("path", [1, 2])
def test_cuh2_pair_sets_match_vesin(cuh2, golden, path):
    for name, c in golden("vesin_pairsets.json")["cases"].items():
        pos, box = np.array(c["positions"]), np.array(c["box"]).reshape(3, 3)
        pairs, shifts = cuh2.neighbors(pos, np.full(len(pos), 29, np.int32), box, path=path)
        assert np.array_equal(canon(pairs, shifts), np.array(c["pairs"]).reshape(-1, 5)), name


@pytest.mark.parametrize("path", [1, 2])
def test_cuh2_physics_suite(cuh2, oracle, path):
    # test_cuh2.f90:71-137, multi-image box
    pos, Z, box = workloads.cu_slab_with_h2()
    cuh2.set_path(path)
    try:
        e, f0, _ = cuh2(pos, Z, box)
        e_o, f_o = oracle.cuh2_force(pos, Z, box)
        assert_energy(e, e_o)
        assert_forces(f0, f_o)
        assert np.abs(f0.sum(0)).max() < 1e-9
        e_t, _, _ = cuh2(pos + 0.37, Z, box)
        assert abs(e_t - e) < 1e-8
        e_r, f_r, _ = cuh2(pos[::-1], Z[::-1], box)
        assert abs(e_r - e) < 1e-8 and np.abs(f_r[::-1] - f0).max() < 1e-5
        delta = 1e-5
        for atom in [0, 4, 8, 13, 19, 26, 31, 32, 33]:
            for c in range(3):
                p = pos.copy()
                p[atom, c] += delta
                ep, _, _ = cuh2(p, Z, box)
                p[atom, c] -= 2 * delta
                em, _, _ = cuh2(p, Z, box)
                assert abs(-(ep - em) / (2 * delta) - f0[atom, c]) < 1e-5
    finally:
        cuh2.set_path(0)


def test_cuh2_errors(cuh2):
    pos, Z, box = workloads.cu_slab_with_h2()
    bad = Z.copy()
    bad[2] = 26
    with pytest.raises(rgpot_b200.PotentialError) as ei:
        cuh2(pos, bad, box)
    assert ei.value.status == 2
    assert str(ei.value) == ("CuH2 potential failed (status 2): rgpot_cuh2: atom 3 has atomic number 26; "
                             "this potential covers Cu (29) and H (1) only")
    p2 = pos.copy()
    p2[1] = p2[0]
    with pytest.raises(rgpot_b200.PotentialError) as ei:
        cuh2(p2, Z, box)
    assert ei.value.status == 3 and "two atoms occupy the same position" in str(ei.value)
    with pytest.raises(rgpot_b200.PotentialError) as ei:
        cuh2(pos, Z, np.zeros((3, 3)))
    assert ei.value.status == 1 and "not invertible" in str(ei.value)
    p3 = pos.copy()
    p3[5, 1] = np.nan
    with pytest.raises(rgpot_b200.PotentialError) as ei:
        cuh2(p3, Z, box)
    assert ei.value.status == 1 and "non-finite" in str(ei.value)
    e, f, _ = cuh2(np.zeros((0, 3)), np.zeros(0, np.int32), box)
    assert e == 0.0 and f.shape == (0, 3)
    # errors on the cell path too
    cuh2.set_path(1)
    try:
        with pytest.raises(rgpot_b200.PotentialError) as ei:
            cuh2(pos, bad, box)
        assert ei.value.status == 2 and "atom 3 has atomic number 26" in str(ei.value)
        with pytest.raises(rgpot_b200.PotentialError) as ei:
            cuh2(p2, Z, box)
        assert ei.value.status == 3
    finally:
        cuh2.set_path(0)


def test_cuh2_triclinic_slab_vs_oracle(cuh2, oracle):
    pos, Z, box = workloads.cu_slab_triclinic(cells=8, n_h2_side=3)  # 2066 atoms, sheared cell
    e, f, _ = cuh2(pos, Z, box)
    e_o, f_o, t = oracle.cuh2_force(pos, Z, box, return_table=True)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    pairs, shifts = cuh2.neighbors(pos, Z, box)
    assert np.array_equal(canon(pairs, shifts), oracle.canonical_pairs(t))
    # unwrapped input (atoms far outside the cell): same physics, per-pair shifts
    rng = np.random.default_rng(1)
    moved = pos + (rng.integers(-3, 4, size=pos.shape) @ box)
    e2, f2, _ = cuh2(moved, Z, box)
    e2_o, f2_o, t2 = oracle.cuh2_force(moved, Z, box, return_table=True)
    assert_energy(e2, e2_o)
    assert_forces(f2, f2_o)
    pairs, shifts = cuh2.neighbors(moved, Z, box)
    assert np.array_equal(canon(pairs, shifts), oracle.canonical_pairs(t2))


def test_cuh2_commensurate_shear(cuh2):
    a = 3.615
    cu = workloads._fcc(4, a)
    Z = np.full(len(cu), 29, np.int32)
    e0, _, _ = cuh2(cu, Z, np.diag([4 * a] * 3))
    e1, _, _ = cuh2(cu, Z, np.array([[4 * a, 0, 0], [a, 4 * a, 0], [a / 2, a / 2, 4 * a]]))
    assert e0 == pytest.approx(-906.183991113513, rel=1e-10)
    assert e1 == pytest.approx(e0, rel=1e-10)


# ---------------------------------------------------------------------------------------- batches
def test_force_batch_heterogeneous(cuh2, oracle):
    slab, Zs, bs = workloads.cuh2_slab218()
    small, Z2, b2 = workloads.cu_slab_with_h2()
    pin, Zp, bp = workloads.cuh2_pin4()
    big, Zb, bb = workloads.cu_slab_triclinic(cells=6, n_h2_side=2)  # 872 atoms -> cell path
    rng = np.random.default_rng(0)
    systems = [(slab + rng.normal(0, 0.05, slab.shape), Zs, bs) for _ in range(5)]
    systems += [(small, Z2, b2), (pin, Zp, bp), (big, Zb, bb), (np.zeros((0, 3)), np.zeros(0, np.int32), bs)]
    out = cuh2.forceBatch(systems)
    assert len(out) == len(systems)
    for (e, f, v), (p, z, b) in zip(out, systems):
        e_o, f_o = oracle.cuh2_force(p, z, b)
        assert_energy(e, e_o)
        assert_forces(f, f_o)
        assert v == 0.0


def test_force_batch_reports_per_system_status(cuh2):
    slab, Zs, bs = workloads.cuh2_slab218()
    bad = Zs.copy()
    bad[7] = 8
    systems = [(slab, Zs, bs), (slab, bad, bs), (slab, Zs, bs)]
    out, statuses = cuh2.forceBatch(systems, return_statuses=True)
    assert list(statuses) == [0, 2, 0]
    assert out[1][0] == 0.0 and not out[1][1].any()  # failed member is zeroed like the reference
    assert out[0][0] == pytest.approx(out[2][0], rel=1e-15)
    with pytest.raises(rgpot_b200.PotentialError) as ei:
        cuh2.forceBatch(systems)
    assert ei.value.status == 2 and "atom 8 has atomic number 8" in str(ei.value)


def test_force_batch_lj(oracle):
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
    systems = []
    for k, n in enumerate([13, 64, 200, 2500]):
        p, b = _random_fluid(n, 0.8, 50 + k, wrapped=False)
        systems.append((p, np.ones(n, np.int32), b))
    for (e, f, _), (p, z, b) in zip(pot.forceBatch(systems), systems):
        e_o, f_o, _ = oracle.lj_force(p, b, 1.0, 2.5, 1.0)
        assert_energy(e, e_o)
        assert_forces(f, f_o)


def test_results_are_bit_reproducible(cuh2):
    offsets, pos, Z, boxes = workloads.cuh2_batch(64)
    e1, f1 = cuh2.forceBatchPacked(offsets, pos, Z, boxes)
    e2, f2 = cuh2.forceBatchPacked(offsets, pos, Z, boxes)
    assert np.array_equal(e1, e2) and np.array_equal(f1, f2)
    big, Zb, bb = workloads.cu_slab_triclinic(cells=7, n_h2_side=2)
    a = cuh2(big, Zb, bb)
    b = cuh2(big, Zb, bb)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_fortran_dropin_symbol(golden):
    """rgpot_cuh2_force / rgpot_fortran_last_error with the reference's exact signature."""
    import ctypes as C
    from rgpot_b200 import _lib
    lib = _lib.lib()
    g = golden("cuh2_pin4.json")
    pos = np.array(g["positions"])
    Z = np.array(g["Z"], np.int32)
    box = np.array(g["box"])
    F = np.zeros_like(pos)
    e = C.c_double()
    st = lib.rgpot_cuh2_force(4, pos.ctypes.data, Z.ctypes.data, box.ctypes.data, F.ctypes.data, C.byref(e))
    assert st == 0
    assert_energy(e.value, g["pin_energy"])
    Z[0] = 3
    st = lib.rgpot_cuh2_force(4, pos.ctypes.data, Z.ctypes.data, box.ctypes.data, F.ctypes.data, C.byref(e))
    buf = C.create_string_buffer(512)
    n = lib.rgpot_fortran_last_error(buf, 512)
    assert st == 2 and n > 0 and b"atom 1 has atomic number 3" in buf.value
    assert e.value == 0.0 and not F.any()


# ---------------------------------------------------------------------------------------- Morse
def test_morse_pins(golden):
    """CppCore/tests/MorsePotTest.cc:62-126 (same pair search as LJ)."""
    pot = rgpot_b200.MorsePot()
    assert pot.energyShift() == pytest.approx(-3.5537997279178057e-05, abs=1e-14)
    two = np.array([[10.0, 10.0, 10.0], [10.0 + 2.8970, 10.0, 10.0]])
    e2, f2, _ = pot(two, np.full(2, 78, np.int32), np.eye(3) * 40.0)
    assert e2 == pytest.approx(-0.71016446200272088, abs=1e-12)
    assert np.abs(f2).max() < 1e-12
    g = golden("lj13_cluster.json")
    pos, box = np.array(g["positions"]), np.array(g["box"]).reshape(3, 3)
    e, f, _ = pot(pos, np.ones(13, np.int32), box)
    assert e == pytest.approx(7483.1166203812, rel=1e-9)
    assert f[0] == pytest.approx([-2056.95286055812, 333.678513150438, 1022.83140169994], rel=1e-9)
    assert np.sqrt((f ** 2).sum(1)).max() == pytest.approx(2480.28362300107, rel=1e-9)
    assert np.abs(f.sum(0)).max() < 1e-9
    deeper = rgpot_b200.MorsePot(rgpot_b200.MorseConfig(De=0.8))
    assert rgpot_b200.MorsePot().paramsKey() == pot.paramsKey() != deeper.paramsKey()
    assert pot.get_type() == rgpot_b200.PotType.Morse


@pytest.mark.parametrize("n,wrapped,path", [(300, False, 2), (3000, True, 1), (3000, False, 1)])
def test_morse_vs_oracle(oracle, n, wrapped, path):
    # a loose Pt-like lattice so the 9.5 A cutoff reaches several shells
    rng = np.random.default_rng(n + path)
    m = int(np.ceil(n ** (1 / 3)))
    L = 2.9 * m
    g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] + 0.5) * 2.9
    pos = g + rng.normal(0, 0.1, g.shape)
    pos = np.mod(pos, L) if wrapped else pos + rng.integers(-1, 2, size=g.shape) * L
    box = np.diag([L, L, L])
    pot = rgpot_b200.MorsePot()
    pot.set_path(path)
    e, f, _ = pot(pos, np.full(n, 78, np.int32), box)
    e_o, f_o, _ = oracle.morse_force(pos, box)
    assert_energy(e, e_o)
    assert_forces(f, f_o)


# ------------------------------------------------------------------------------------------ ZBL
def test_zbl_pins():
    """CppCore/tests/ZBLPotTest.cc:36-160 (free boundaries, species tables)."""
    pot = rgpot_b200.ZBLPot()
    dimer = np.array([[9.40, 9.35, 9.30], [10.60, 10.65, 10.70]])
    cell = np.eye(3) * 20.0
    e, f, _ = pot(dimer, [14, 79], cell)
    assert e == pytest.approx(0.38537731, abs=1e-8)  # LAMMPS pair_style zbl 2.0 2.5
    assert np.abs(f - np.array([[-2.37926, -2.57753, -2.7758], [2.37926, 2.57753, 2.7758]])).max() < 1e-5
    assert e == pytest.approx(0.38537730883, rel=1e-9)
    # tables follow the species set of each call (ZBLPotTest.cc:91-115)
    assert pot(dimer, [79, 79], cell)[0] == pytest.approx(0.961609764914, rel=1e-9)
    assert pot(dimer, [14, 14], cell)[0] == pytest.approx(0.167604056792, rel=1e-9)
    assert pot(dimer, [14, 79], cell)[0] == e
    # the cell is ignored (:117-131)
    assert pot(dimer, [14, 79], np.eye(3) * 3.0)[0] == e
    assert not pot.caps().periodic
    # switching zeroes energy and force at the cutoff (:57-89)
    e_cut, f_cut, _ = pot(np.array([[5.0, 5.0, 5.0], [7.5, 5.0, 5.0]]), [14, 79], cell)
    assert abs(e_cut) < 1e-12 and np.abs(f_cut).max() < 1e-12
    e_far, f_far, _ = pot(np.array([[5.0, 5.0, 5.0], [8.0, 5.0, 5.0]]), [14, 79], cell)
    assert e_far == 0.0 and not f_far.any()
    # invalid cutoff pairs (:133-146), fingerprint (:148-157)
    for bad in [(2.5, 2.0), (2.0, 2.0), (0.0, 2.5)]:
        with pytest.raises(ValueError):
            rgpot_b200.ZBLPot(rgpot_b200.ZBLConfig(*bad))
    wide = rgpot_b200.ZBLPot(rgpot_b200.ZBLConfig(2.0, 4.5))
    assert rgpot_b200.ZBLPot().paramsKey() == pot.paramsKey() != wide.paramsKey()


@pytest.mark.parametrize("n,path", [(200, 2), (2500, 1), (2500, 0)])
def test_zbl_vs_oracle(oracle, n, path):
    rng = np.random.default_rng(n + path)
    L = (n / 0.09) ** (1 / 3)  # a few neighbours inside 2.5 A
    pos = rng.random((n, 3)) * L
    Z = rng.choice([1, 8, 14, 29, 79], n).astype(np.int32)
    pot = rgpot_b200.ZBLPot()
    pot.set_path(path)
    e, f, _ = pot(pos, Z, np.eye(3) * 5.0)
    e_o, f_o, _ = oracle.zbl_force(pos, Z)
    assert_energy(e, e_o)
    assert (np.abs(f - f_o) <= 1e-9 + 1e-10 * np.abs(f_o)).all()
    # batches share one table over the union of the species
    half = n // 2
    out = pot.forceBatch([(pos[:half], Z[:half], np.eye(3)), (pos[half:], Z[half:], np.eye(3))])
    for (eb, fb, _), sl in zip(out, (slice(0, half), slice(half, n))):
        eo, fo, _ = oracle.zbl_force(pos[sl], Z[sl])
        assert_energy(eb, eo)
        assert (np.abs(fb - fo) <= 1e-9 + 1e-10 * np.abs(fo)).all()


def test_zbl_too_many_species():
    pot = rgpot_b200.ZBLPot()
    pos = np.random.default_rng(0).random((20, 3)) * 6
    with pytest.raises(rgpot_b200.PotentialError) as ei:
        pot(pos, np.arange(1, 21, dtype=np.int32), np.eye(3) * 6)
    assert ei.value.status == 103


# ------------------------------------------------------------------------- brick kernels (brick.cuh)
# The large-system path: half-cutoff cells (fine = 2), private survivor lists evaluated in rounds,
# oversized bricks cut in two.  The engine options shrink the capacities so that small systems walk
# every one of those branches; all variants must agree with the oracle AND bit for bit with each
# other (the summation order of an atom does not depend on the capacities).
BRICK_VARIANTS = [
    {},                                                   # automatic
    {"fine": 1},                                          # cutoff-sized cells, 27-cell stencil
    {"brick_list": 3},                                    # many search / physics rounds per atom
    {"brick_slack": -70},                                 # staged capacity far too small: bricks are cut
    {"brick": 111, "brick_threads": 64},                  # one cell per brick
    {"brick": 421, "brick_threads": 96, "brick_list": 7},
]


def _with_options(pot, opts):
    pot.set_option("fine_min", 0)
    pot.set_option("tile", 0)  # these are the brick kernels' tests (tests/test_gpu_tile.py has the tile kernels')
    for key in ("fine", "brick_list", "brick_slack", "brick", "brick_threads", "brick_tpa"):
        pot.set_option(key, opts.get(key, 0))


@pytest.mark.parametrize("wrapped", [True, False])
def test_brick_lj_variants_vs_oracle(oracle, wrapped):
    n = 4000
    pos, box = _random_fluid(n, 0.8442, 77, wrapped)
    z = np.ones(n, np.int32)
    e_o, f_o, _ = oracle.lj_force(pos, box, 1.0, 2.5, 1.0)
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
    pot.set_path(1)
    first = {}
    for opts in BRICK_VARIANTS:
        for tpa in (1, 2, 4):
            _with_options(pot, dict(opts, brick_tpa=tpa))
            e, f, _ = pot(pos, z, box)
            assert_energy(e, e_o, (opts, tpa))
            assert_forces(f, f_o, (opts, tpa))
            assert pot.stat("slow_atoms") == 0
            key = (tpa, opts.get("fine", 0))  # same threads per atom + same cell grid => same order
            if key in first:
                assert np.array_equal(f, first[key]), (opts, tpa)
            else:
                first[key] = f


def test_brick_cuh2_variants_vs_oracle(cuh2, oracle):
    pos, Z, box = workloads.cu_slab_triclinic(cells=8, n_h2_side=3)  # 2066 atoms, sheared cell
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    pot = rgpot_b200.CuH2Pot()
    pot.set_path(1)
    first = {}
    for opts in BRICK_VARIANTS:
        for tpa in (2, 4):
            _with_options(pot, dict(opts, brick_tpa=tpa))
            e, f, _ = pot(pos, Z, box)
            assert_energy(e, e_o, (opts, tpa))
            assert_forces(f, f_o, (opts, tpa))
            if "brick_slack" not in opts:
                # the slab's bulk is denser than its mean, so a first call may cut bricks or walk a
                # cell on the slow path; the engine remembers the halo populations it met and the
                # next call with the same signature is sized for them
                e, f, _ = pot(pos, Z, box)
                assert_energy(e, e_o, (opts, tpa))
                assert_forces(f, f_o, (opts, tpa))
                assert pot.stat("slow_atoms") == 0
                key = (tpa, opts.get("fine", 0))
                if key in first:
                    assert np.array_equal(f, first[key]), (opts, tpa)
                else:
                    first[key] = f


def test_brick_morse_zbl_fine_cells(oracle):
    n = 3000
    pos, box = _random_fluid(n, 0.041, 5, True)  # 2.9 A lattice: the 9.5 A cutoff reaches several shells
    z = np.ones(n, np.int32)
    pot = rgpot_b200.MorsePot()
    pot.set_path(1)
    _with_options(pot, {"brick_list": 5})
    e, f, _ = pot(pos, z, box)
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


def test_brick_dense_cell_falls_back_in_kernel(oracle):
    """A cell whose own stencil exceeds the staged capacity takes the per-atom gather walk inside
    the kernel (reported by the slow_atoms diagnostic); results still match the oracle."""
    rng = np.random.default_rng(4)
    n = 6000
    pos, box = _random_fluid(n, 0.8442, 9, True)
    # a dense blob: 600 atoms inside one small region (LJ with a tiny well so forces stay finite)
    blob = rng.random((600, 3)) * 1.2 + 4.0
    pos = np.vstack([pos, blob])
    z = np.ones(len(pos), np.int32)
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0e-6, 2.5, 0.3))
    pot.set_path(1)
    _with_options(pot, {"brick_slack": -60})
    e, f, _ = pot(pos, z, box)
    e_o, f_o, _ = oracle.lj_force(pos, box, 1.0e-6, 2.5, 0.3)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    assert pot.stat("slow_atoms") > 0


# ------------------------------------------------------- rgpot-core boundary (rgpot_b200_core.h)
def test_core_callback_cpu_tensors(cuh2, golden, oracle):
    """rgpot_b200_core_callback with kDLCPU DLPack tensors: CuH2PotTest.cc pin and the slab."""
    from rgpot_b200 import core
    g = golden("cuh2_pin4.json")
    pos, Z, box = np.array(g["positions"]), np.array(g["Z"], np.int32), np.array(g["box"])
    e, f, v = core.calculate(cuh2, pos, Z, box)
    assert e == pytest.approx(g["pin_energy"], rel=1e-10) and v == 0.0  # CuH2PotTest.cc:26
    assert np.abs(f - np.array(g["pin_forces"]).reshape(-1, 3)).max() < 1e-9
    e2, f2, _ = cuh2(pos, Z, box.reshape(3, 3))
    assert e == e2 and np.array_equal(f, f2)  # the same engine call underneath
    # LJ without atomic numbers (optional tensor, potential.hpp:219-222)
    lj = rgpot_b200.LJPot()
    c = golden("lj13_cluster.json")
    e, f, _ = core.calculate(lj, np.array(c["positions"]), None, np.array(c["box"]))
    assert e == pytest.approx(c["pin_energy"], rel=1e-9)


def test_core_callback_cuda_tensors(cuh2, oracle):
    """kDLCUDA tensors ("any device", rgpot.h:78-88): evaluated in place, forces come back on the GPU."""
    import torch

    from rgpot_b200 import core
    pos, Z, box = workloads.cu_slab_triclinic(cells=6, n_h2_side=2)
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    dev = torch.device("cuda", 0)
    e, f, _ = core.calculate(cuh2, torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev), box)
    assert f.is_cuda
    assert_energy(e, e_o)
    assert_forces(f.cpu().numpy(), f_o)


def test_core_callback_errors(cuh2):
    from rgpot_b200 import core
    pos = np.zeros((4, 3))
    with pytest.raises(core.CoreError) as ei:  # wrong shape
        core.calculate(cuh2, np.zeros((4, 2)), np.full(4, 29, np.int32), np.eye(3) * 20)
    assert ei.value.status == 1
    with pytest.raises(core.CoreError) as ei:  # Z of the wrong length
        core.calculate(cuh2, pos, np.full(3, 29, np.int32), np.eye(3) * 20)
    assert ei.value.status == 1
    with pytest.raises(core.CoreError) as ei:  # engine failure: foreign species -> status 2 message
        core.calculate(cuh2, np.random.default_rng(0).random((4, 3)) * 5, np.array([29, 29, 8, 1], np.int32),
                       np.eye(3) * 20)
    assert ei.value.status == 2 and "atom 3 has atomic number 8" in str(ei.value)


# ------------------------------------------------------------- EAM-Al
AL_FCC = np.array([[7.975, 7.975, 7.975], [10.0, 10.0, 7.975], [10.0, 7.975, 10.0], [7.975, 10.0, 10.0]])


def _al_crystal(cells, seed, rattle=0.08):
    a = 4.05
    pos = workloads._fcc(cells, a) + np.random.default_rng(seed).normal(0, rattle, (4 * cells ** 3, 3))
    return pos, np.eye(3) * cells * a


def test_eam_al_eon_pin_and_dropin_symbol():
    """FortranPotsTest.cc:135-144 through the class and through rgpot_eam_al_force, the symbol
    FortranPots.cc:26-28 binds."""
    import ctypes as C

    from rgpot_b200 import _lib
    pot = rgpot_b200.EAMAlPot()
    e, f, v = pot(AL_FCC, np.full(4, 13, np.int32), np.eye(3) * 20.0)
    assert e == pytest.approx(-5.217864, rel=1e-4) and v == 0.0
    assert np.sqrt((f ** 2).sum(1)).max() == pytest.approx(0.968647, rel=1e-3)
    assert np.abs(f.sum(0)).max() < 1e-9
    assert pot.get_type() == rgpot_b200.PotType.EAMAl
    lib = _lib.lib()
    F = np.zeros((4, 3))
    e2 = C.c_double()
    cell = np.ascontiguousarray(np.eye(3) * 20.0)
    st = lib.rgpot_eam_al_force(4, AL_FCC.ctypes.data, cell.ctypes.data, F.ctypes.data, C.byref(e2))
    assert st == 0 and e2.value == e and np.array_equal(F, f)


@pytest.mark.parametrize("cells,path,opts", [
    (2, 2, {}), (3, 2, {}), (3, 1, {}),            # boxes narrower than two cutoffs: several images per pair
    (4, 2, {}), (4, 1, {}), (6, 1, {}),
    (10, 1, {}), (10, 1, {"fine": 1}), (10, 1, {"brick_list": 5, "brick_slack": -50}),
])
def test_eam_al_vs_oracle(oracle, cells, path, opts):
    pos, box = _al_crystal(cells, cells + path)
    e_o, f_o = oracle.eam_al_force(pos, box)
    pot = rgpot_b200.EAMAlPot()
    pot.set_path(path)
    _with_options(pot, opts)
    e, f, _ = pot(pos, None, box)
    assert_energy(e, e_o, (cells, path, opts))
    assert_forces(f, f_o, (cells, path, opts))


def test_eam_al_batch_and_coexistence_with_cuh2(cuh2, oracle, golden):
    """EAM-Al and CuH2 handles alive together: each model has its own constant block."""
    al = rgpot_b200.EAMAlPot()
    systems = [(_al_crystal(4, s)[0], None, _al_crystal(4, s)[1]) for s in range(5)]
    out = al.forceBatch(systems)
    g = golden("cuh2_pin4.json")
    e_cu, _, _ = cuh2(np.array(g["positions"]), np.array(g["Z"], np.int32), np.array(g["box"]).reshape(3, 3))
    assert e_cu == pytest.approx(g["pin_energy"], rel=1e-10)
    for (e, f, _), (p, _, b) in zip(out, systems):
        e_o, f_o = oracle.eam_al_force(p, b)
        assert_energy(e, e_o)
        assert_forces(f, f_o)


# ------------------------------------------------- vesin-shaped neighbor lists
def _nl_canon(nl):
    rec = np.column_stack([nl["pairs"], nl["shifts"].astype(np.int64)])
    order = np.lexsort((rec[:, 4], rec[:, 3], rec[:, 2], rec[:, 1], rec[:, 0])) if len(rec) else np.zeros(0, int)
    return rec[order], nl["vectors"][order], nl["distances"][order]


@pytest.mark.parametrize("case", ["triclinic", "small_box", "unwrapped", "free"])
@pytest.mark.parametrize("full", [True, False])
def test_neighbor_list_is_bit_identical_to_vesin(oracle, case, full):
    """rgpot_b200_neighbor_list against the vendored vesin itself: the same pair set (self images
    included), bit-identical vectors and distances, full and half lists."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so (the vendored vesin) is not built")
    rng = np.random.default_rng(12)
    periodic, cutoff = True, 4.2
    if case == "triclinic":
        pos, _, box = workloads.cu_slab_triclinic(cells=5, n_h2_side=2)
    elif case == "small_box":  # box narrower than the cutoff: several images per pair, self images
        box = np.array([[3.1, 0.0, 0.0], [0.4, 3.3, 0.0], [0.2, -0.3, 3.6]])
        pos = rng.random((12, 3)) @ box
    elif case == "unwrapped":
        pos, _, box = workloads.cu_slab_triclinic(cells=4, n_h2_side=2)
        pos = pos + rng.integers(-3, 4, size=pos.shape) @ box
    else:
        periodic = False
        pos = rng.random((900, 3)) * 30.0 - 7.0
        box = np.eye(3)
    ours = rgpot_b200.neighbor_list(pos, box, cutoff, full=full, periodic=periodic)
    ref = oracle.vesin_raw(pos, box, cutoff, full=full, periodic=periodic)
    a, av, ad = _nl_canon(ours)
    b, bv, bd = _nl_canon(ref)
    assert len(a) == len(b) and len(a) > 0
    assert np.array_equal(a, b)
    assert np.array_equal(av, bv) and np.array_equal(ad, bd)
    if case == "small_box":
        assert (a[:, 0] == a[:, 1]).any()  # self images present
    # sorted output: ordered by the first index
    assert (np.diff(ours["pairs"][:, 0]) >= 0).all()


# ------------------------------------------------------------- FeHe
FE_BCC = np.array([[7.13, 7.13, 7.13], [8.565, 8.565, 8.565], [7.13, 7.13, 10.0], [8.565, 8.565, 11.435],
                   [7.13, 10.0, 7.13], [8.565, 11.435, 8.565], [7.13, 10.0, 10.0], [8.565, 11.435, 11.435],
                   [10.0, 7.13, 7.13], [11.435, 8.565, 8.565], [10.0, 7.13, 10.0], [11.435, 8.565, 11.435],
                   [10.0, 10.0, 7.13], [11.435, 11.435, 8.565], [10.0, 10.0, 10.0], [11.435, 11.435, 11.435]])


def _fehe_cell(n_side, n_he, seed, rattle=0.04, box=None):
    a = 2.8553
    base = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    cells = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    frac = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3) / n_side
    rng = np.random.default_rng(seed)
    H = np.eye(3) * n_side * a if box is None else box
    pos = frac @ H + rng.normal(0, rattle, frac.shape)
    Z = np.full(len(pos), 26, np.int32)
    Z[rng.choice(len(pos), size=n_he, replace=False)] = 2
    return pos, Z, H


def test_fehe_eon_pin_and_dropin_symbol():
    """FortranPotsTest.cc:146-155 through the class and through rgpot_fehe_force."""
    import ctypes as C

    from rgpot_b200 import _lib
    pot = rgpot_b200.FeHePot()
    Z = np.full(16, 26, np.int32)
    e, f, v = pot(FE_BCC, Z, np.eye(3) * 20.0)
    assert e == pytest.approx(-43.959774, rel=1e-4) and v == 0.0
    assert np.sqrt((f ** 2).sum(1)).max() == pytest.approx(1.245094, rel=1e-3)
    assert np.abs(f.sum(0)).max() < 1e-9
    assert pot.get_type() == rgpot_b200.PotType.FeHe
    F = np.zeros((16, 3))
    e2 = C.c_double()
    cell = np.ascontiguousarray(np.eye(3) * 20.0)
    st = _lib.lib().rgpot_fehe_force(16, FE_BCC.ctypes.data, Z.ctypes.data, cell.ctypes.data, F.ctypes.data, C.byref(e2))
    assert st == 0 and e2.value == e and np.array_equal(F, f)


@pytest.mark.parametrize("n_side,n_he,triclinic", [(2, 2, False), (3, 6, False), (4, 2, False), (6, 40, False),
                                                   (5, 25, True), (10, 200, False)])
def test_fehe_vs_oracle(oracle, n_side, n_he, triclinic):
    """All three pair channels and both density channels, boxes narrower than two cutoffs (several
    images per pair), a sheared cell, unwrapped input."""
    box = None
    if triclinic:
        a = 2.8553 * n_side
        box = np.array([[a, 0.0, 0.0], [0.2 * a, a, 0.0], [0.1 * a, -0.15 * a, a]])
    pos, Z, H = _fehe_cell(n_side, n_he, 100 + n_side, box=box)
    e_o, f_o = oracle.fehe_force(pos, Z, H)
    pot = rgpot_b200.FeHePot()
    e, f, _ = pot(pos, Z, H)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    moved = pos + np.random.default_rng(1).integers(-2, 3, size=pos.shape) @ H
    e2, f2, _ = pot(moved, Z, H)
    e2_o, f2_o = oracle.fehe_force(moved, Z, H)
    assert_energy(e2, e2_o)
    assert_forces(f2, f2_o)


@pytest.mark.parametrize("triclinic", [False, True])
def test_fehe_brick_kernels_large(oracle, triclinic):
    """21 296 atoms, 1 % helium: half-cutoff cells and the brick kernels (two density channels staged
    with the halo), against the oracle and against the gather kernels; a tenth of the atoms outside
    the cell."""
    n_side = 22
    box = None
    if triclinic:
        a = 2.8553 * n_side
        box = np.array([[a, 0.0, 0.0], [0.2 * a, a, 0.0], [0.1 * a, -0.15 * a, a]])
    pos, Z, H = _fehe_cell(n_side, 213, 7, box=box)
    assert len(pos) > 20000
    frac = pos @ np.linalg.inv(H)
    pos = (frac - np.floor(frac)) @ H  # wrapped, then a tenth of the atoms one cell vector outside
    pos = _scatter_outside(pos, H, np.random.default_rng(8))
    pot = rgpot_b200.FeHePot()
    e, f, _ = pot(pos, Z, H)
    assert pot.stat("max_halo") > 0 and pot.stat("pool_used") > 0  # bricks ran, lists were handed on
    e_o, f_o = oracle.fehe_force(pos, Z, H)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    pot.set_option("bricks", 0)
    e_g, f_g, _ = pot(pos, Z, H)
    assert_energy(e, e_g)
    assert_forces(f, f_g)


def test_fehe_errors_and_batch(oracle):
    pot = rgpot_b200.FeHePot()
    pos, Z, H = _fehe_cell(3, 4, 7)
    bad = Z.copy()
    bad[5] = 29
    with pytest.raises(rgpot_b200.PotentialError) as ei:
        pot(pos, bad, H)
    assert ei.value.status == 2
    assert "rgpot_fehe: unsupported atomic number 29 on atom 6; this potential covers iron (26) and helium (2) only" in str(ei.value)
    systems = [_fehe_cell(3, k + 1, 20 + k) for k in range(4)]
    out = pot.forceBatch(systems)
    for (e, f, _), (p, z, h) in zip(out, systems):
        e_o, f_o = oracle.fehe_force(p, z, h)
        assert_energy(e, e_o)
        assert_forces(f, f_o)


def test_zbl_device_entry_maps_species_on_the_device(oracle):
    """rgpot_b200_force_device for ZBL: species set from a presence bitmap, type indices mapped on the
    device; same numbers as the host-buffer call."""
    import torch
    rng = np.random.default_rng(8)
    n = 3000
    pos = rng.random((n, 3)) * (n / 0.09) ** (1 / 3)
    Z = rng.choice([1, 8, 14, 29, 79], n).astype(np.int32)
    pot = rgpot_b200.ZBLPot()
    e_h, f_h, _ = pot(pos, Z, np.eye(3))
    dev = torch.device("cuda", 0)
    d_pos, d_Z = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev)
    d_F = torch.empty((n, 3), dtype=torch.float64, device=dev)
    d_E = torch.empty(1, dtype=torch.float64, device=dev)
    pot.force_device(d_pos, d_Z, np.eye(3), d_F, d_E)
    assert float(d_E.item()) == e_h and np.array_equal(d_F.cpu().numpy(), f_h)
    e_o, f_o, _ = oracle.zbl_force(pos, Z)
    assert_energy(e_h, e_o)


# ------------------------------------------------ inhomogeneous large systems on the automatic path
def test_brick_auto_path_inhomogeneous_lj(oracle):
    """A dense droplet in a dilute gas, 24 000 atoms, periodic: half-cutoff cells and bricks are
    chosen automatically, the mean density badly underestimates the droplet (bricks are cut on the
    first call, sized from the feedback on the second); both calls match the oracle bit-identically
    to each other."""
    rng = np.random.default_rng(21)
    L = 120.0
    m = 28
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    g = g[((g - (m - 1) / 2) ** 2).sum(1) <= (m / 2) ** 2]  # a ball of a simple-cubic lattice, spacing 1.1
    drop = g * 1.1 + 40.0 + rng.normal(0, 0.05, g.shape)
    n_gas = 24000 - len(drop)
    gm = int(np.ceil(n_gas ** (1 / 3)))
    gas = (np.stack(np.meshgrid(*[np.arange(gm)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n_gas] + 0.5) * (L / gm)
    gas = gas[(((gas - 55.4) ** 2).sum(1)) > 20.0 ** 2]  # keep the gas out of the droplet
    pos = np.vstack([drop, gas])
    box = np.diag([L, L, L])
    z = np.ones(len(pos), np.int32)
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
    e1, f1, _ = pot(pos, z, box)
    e2, f2, _ = pot(pos, z, box)
    e_o, f_o, _ = oracle.lj_force(pos, box, 1.0, 2.5, 1.0)
    assert_energy(e1, e_o)
    assert_forces(f1, f_o)
    assert e1 == e2 and np.array_equal(f1, f2)
    assert pot.stat("slow_atoms") == 0


def test_brick_auto_path_nanoparticle_cuh2(oracle):
    """A 22 000-atom Cu nanoparticle with adsorbed H in a large triclinic cell (mostly vacuum)."""
    a = 3.615
    cells = 22
    cu = workloads._fcc(cells, a)
    c = cu.mean(0)
    cu = cu[((cu - c) ** 2).sum(1) <= (0.5 * cells * a) ** 2]
    rng = np.random.default_rng(22)
    cu = cu + rng.normal(0, 0.03, cu.shape)
    shell = rng.normal(size=(300, 3))
    h = c + shell / np.linalg.norm(shell, axis=1)[:, None] * (0.5 * cells * a + 1.6)
    pos = np.vstack([cu, h]) + 30.0
    Z = np.concatenate([np.full(len(cu), 29), np.full(len(h), 1)]).astype(np.int32)
    Lb = cells * a + 60.0
    box = np.array([[Lb, 0.0, 0.0], [0.15 * Lb, Lb, 0.0], [0.1 * Lb, -0.2 * Lb, Lb]])
    assert len(pos) > 20000
    pot = rgpot_b200.CuH2Pot()
    e1, f1, _ = pot(pos, Z, box)
    e2, f2, _ = pot(pos, Z, box)
    e_o, f_o = oracle.cuh2_force(pos, Z, box)
    assert_energy(e1, e_o)
    assert_forces(f1, f_o)
    assert e1 == e2 and np.array_equal(f1, f2)


def test_torch_interop_and_evaluate_batch(cuh2, golden, oracle):
    """CUDA tensors in, CUDA tensors out on torch's current stream (nothing
    crosses PCIe but the box), and the module-level batch function."""
    import torch

    s = golden("cuh2_slab218.json")
    pos, Z, box = np.array(s["positions"]), np.array(s["Z"], np.int32), np.array(s["box"]).reshape(3, 3)
    e_h, f_h, _ = cuh2(pos, Z, box)
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        d_pos, d_Z = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev)
        e_t, f_t = cuh2.force_torch(d_pos, d_Z, box)
        e_a, f_a = cuh2.force_torch(d_pos, d_Z.to(torch.int64), box, sync=False)  # other integer types are converted
        assert cuh2.sync_status() == 0
    assert f_t.is_cuda and f_t.shape == (len(pos), 3)
    assert_energy(e_t, e_h)
    assert_forces(f_t.cpu().numpy(), f_h)
    assert float(e_a.item()) == e_t and torch.equal(f_a, f_t)
    with pytest.raises(ValueError):
        cuh2.force_torch(torch.from_numpy(pos), d_Z, box)  # host tensor
    c13 = golden("lj13_cluster.json")
    lj_pos, lj_box = np.array(c13["positions"]), np.array(c13["box"]).reshape(3, 3)
    out = rgpot_b200.evaluate_batch([(pos, Z, box), (pos + 0.01, Z, box)])
    assert out[0][0] == pytest.approx(e_h, rel=1e-12) and len(out) == 2
    out_lj = rgpot_b200.evaluate_batch([(lj_pos, np.ones(len(lj_pos), np.int32), lj_box)], potential="LJ")
    e_o, f_o, _ = oracle.lj_force(lj_pos, lj_box, 1.0, 15.0, 1.0)
    assert_energy(out_lj[0][0], e_o)
    assert_forces(out_lj[0][1], f_o)


def _scatter_outside(pos, box, rng, far=0):
    """Atoms that left the cell, as they do in a molecular-dynamics run without re-wrapping: about a
    tenth moved by one cell vector (any direction), `far` atoms by two."""
    pos = pos.copy()
    n = len(pos)
    who = rng.choice(n, n // 10, replace=False)
    pos[who] += rng.integers(-1, 2, (len(who), 3)) @ box
    if far:
        pos[rng.choice(n, far, replace=False)] += np.array([2, 0, -2]) @ box
    return pos


@pytest.mark.parametrize("far", [0, 3])
def test_brick_atoms_outside_the_cell_lj(oracle, far):
    """The brick kernels take atoms up to one cell vector outside the cell (their wrap counts are
    folded into the image codes); further out, the gather kernels take over.  Same results, and the
    same as for the wrapped positions up to the reference's own rounding of r_j - r_i."""
    pos, Z, box = workloads.lj_fluid(24_000)
    rng = np.random.default_rng(31)
    out = _scatter_outside(pos, box, rng, far)
    pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
    e, f, _ = pot(out, Z, box)
    assert (pot.stat("max_halo") > 0) == (far == 0)  # which kernels did the work
    e_o, f_o, _ = oracle.lj_force(out, box, 1.0, 2.5, 1.0)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    pot.set_option("bricks", 0)  # gather kernels, literal fold
    e_g, f_g, _ = pot(out, Z, box)
    assert_energy(e, e_g)
    assert_forces(f, f_g)


@pytest.mark.parametrize("far", [0, 2])
def test_brick_atoms_outside_the_cell_cuh2(oracle, far):
    pos, Z, box = workloads.cu_slab_triclinic(cells=18)
    assert len(pos) > 20000
    rng = np.random.default_rng(32)
    out = _scatter_outside(pos, box, rng, far)
    pot = rgpot_b200.CuH2Pot()
    e, f, _ = pot(out, Z, box)
    assert (pot.stat("max_halo") > 0) == (far == 0)
    e_o, f_o = oracle.cuh2_force(out, Z, box)
    assert_energy(e, e_o)
    assert_forces(f, f_o)
    pot.set_option("bricks", 0)  # gather kernels (vesin's per-pair shift form)
    e_g, f_g, _ = pot(out, Z, box)
    assert_energy(e, e_g)
    assert_forces(f, f_g)


def test_cache_keys_bit_identical_to_reference_hash(cuh2, oracle, golden):
    """rgpot_b200_cache_keys == Potential::cacheKey (XXH3 of positions, species,
    box, type, paramsKey) for every input-length class: empty systems, 1..12 atoms (short forms of
    both arrays), 61 atoms (long form for the species array starts at 61), 218, 1000 (several
    1024-byte blocks), and the committed xxhash vectors."""
    rng = np.random.default_rng(41)
    box = np.diag([15.3456, 21.702, 100.0])
    systems = []
    for n in [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 42, 43, 60, 61, 64, 128, 218, 219, 1000, 2731]:
        pos = rng.random((n, 3)) * 20.0
        Z = np.where(rng.random(n) < 0.1, 1, 29).astype(np.int32)
        systems.append((pos, Z, box * (1.0 + 0.01 * n)))
    keys = cuh2.cacheKeys(systems)
    for (pos, Z, b), k in zip(systems, keys.tolist()):
        assert k == oracle.cache_key(pos, Z, b, rgpot_b200.PotType.CuH2.value, cuh2.paramsKey()), len(pos)
    for c in golden("xxh3_vectors.json")["cache_keys"]:  # digests of libxxhash itself
        class Fixed(type(cuh2)):
            def paramsKey(self):
                return c["params_key"]
        pot = Fixed()
        k = pot.cacheKeys([(np.array(c["positions"]).reshape(-1, 3), np.array(c["Z"], np.int32),
                            np.array(c["box"]).reshape(3, 3))])
        assert int(k[0]) == int(c["key"])
    lj = rgpot_b200.LJPot()
    assert lj.cacheKeys(systems[5:6])[0] != keys[5]  # type and parameter fingerprint enter the key


def test_force_batch_cached_compacts_misses(cuh2, golden):
    """Potential::forceBatch with a cache (Potential.hpp:251-306): hits are served, only the misses
    reach the kernel, results are written back."""
    s = golden("cuh2_slab218.json")
    pos, Z, box = np.array(s["positions"]), np.array(s["Z"], np.int32), np.array(s["box"]).reshape(3, 3)
    rng = np.random.default_rng(42)
    band = [(pos + rng.normal(0, 0.02, pos.shape), Z, box) for _ in range(6)]
    cache = {}
    first, missed = cuh2.forceBatchCached(band, cache)
    assert missed == 6 and len(cache) == 6
    band[2] = (band[2][0] + 0.01, Z, box)  # one image moved
    calls = cuh2.force_calls
    again, missed = cuh2.forceBatchCached(band, cache)
    assert missed == 1 and cuh2.force_calls == calls + 1
    direct = cuh2.forceBatch(band)
    for a, d in zip(again, direct):
        assert a[0] == d[0] and np.array_equal(a[1], d[1])
