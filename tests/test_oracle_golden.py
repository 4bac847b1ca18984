"""The oracle against every golden vector the reference's own tests hold for this path.

Not a GPU test: this pins the CHECKER (oracle/), so that parity of the CUDA path against the
oracle means parity against the reference.
"""
import numpy as np
import pytest

from rgpot_b200 import workloads


def test_lj_cluster_pins(oracle, golden):
    # CppCore/tests/LJClusterPotTest.cc:55-69
    g = golden("lj13_cluster.json")
    pos, box = np.array(g["positions"]), np.array(g["box"]).reshape(3, 3)
    e, f, npairs = oracle.lj_force(pos, box, cluster=True)
    assert npairs == 78
    assert e == pytest.approx(g["pin_energy"], rel=1e-9)
    assert e == pytest.approx(g["pin_eon_energy"], rel=1e-4)
    assert np.sqrt((f ** 2).sum(1)).max() == pytest.approx(g["pin_max_force"], rel=1e-9)
    # against the reference object itself: identical scan order, so bit-exact
    assert e == g["ref_cluster_energy"]
    assert np.array_equal(f, np.array(g["ref_cluster_forces"]))


def test_lj_periodic_equals_cluster(oracle, golden):
    # LJClusterPotTest.cc:72-94
    g = golden("lj13_cluster.json")
    pos, box = np.array(g["positions"]), np.array(g["box"]).reshape(3, 3)
    e_c, f_c, _ = oracle.lj_force(pos, box, cluster=True)
    e_p, f_p, _ = oracle.lj_force(pos, box, cluster=False)
    assert e_c == pytest.approx(e_p, rel=1e-12)
    assert np.abs(f_c - f_p).max() < 1e-12
    assert e_p == g["ref_periodic_energy"]


def test_lj_cluster_ignores_cell(oracle, golden):
    # LJClusterPotTest.cc:96-112
    g = golden("lj13_cluster.json")
    pos, box = np.array(g["positions"]), np.array(g["box"]).reshape(3, 3)
    e_wide, _, _ = oracle.lj_force(pos, box, cluster=True)
    e_tight, _, _ = oracle.lj_force(pos, np.eye(3) * 5.0, cluster=True)
    assert e_tight == pytest.approx(e_wide, rel=1e-12)


def test_lj_reference_cases(oracle, golden):
    # outputs of the unmodified LJPot.cc (oracle/_ref) on periodic boxes incl. rc > L/2
    for name, c in golden("lj_reference_cases.json")["cases"].items():
        pos, box = np.array(c["positions"]), np.array(c["box"]).reshape(3, 3)
        e, f, npairs = oracle.lj_force(pos, box, c["u0"], c["cutoff"], c["psi"])
        assert e == c["energy"], name
        assert np.array_equal(f, np.array(c["forces"])), name
        assert npairs == len(c["pairs"])


def test_lj_invariance_dimer(oracle):
    # CppCore/tests/InvarianceTest.cc:25-89: dimer r = 1.5, 10 A box, rc = 15 > L/2
    box = np.eye(3) * 10.0
    pos = np.array([[1.0, 1.0, 1.0], [2.5, 1.0, 1.0]])
    e0, f0, _ = oracle.lj_force(pos, box)
    e1, _, _ = oracle.lj_force(pos + np.array([3.3, -2.1, 0.7]), box)
    assert e1 == pytest.approx(e0, abs=1e-12)
    rot = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    e2, _, _ = oracle.lj_force((pos - pos[0]) @ rot.T + pos[0], box)
    assert e2 == pytest.approx(e0, abs=1e-12)
    assert np.abs(f0[0] + f0[1]).max() < 1e-9  # tests/test_rgpot_pip_core.py:27-56


def test_cuh2_pin4(oracle, golden):
    # CppCore/tests/CuH2PotTest.cc:11-42 -- 17-digit energy and 12 force components
    g = golden("cuh2_pin4.json")
    e, f = oracle.cuh2_force(g["positions"], g["Z"], np.array(g["box"]).reshape(3, 3))
    assert abs(e - g["pin_energy"]) <= 2e-15 * abs(g["pin_energy"])
    assert np.abs(f - np.array(g["pin_forces"])).max() < 1e-13


def test_cuh2_dimer(oracle, golden):
    # tests/rpc_integ.py:52-78
    g = golden("cuh2_dimer.json")
    e, f = oracle.cuh2_force(g["positions"], g["Z"], np.array(g["box"]).reshape(3, 3))
    assert e == pytest.approx(g["pin_energy"], rel=g["pin_rel_tol"])
    assert f[0, 0] == pytest.approx(g["pin_f0x"], rel=g["pin_rel_tol"])
    assert f[1, 0] == pytest.approx(g["pin_f1x"], rel=g["pin_rel_tol"])


def test_cuh2_slab218_fixture(oracle, golden):
    g = golden("cuh2_slab218.json")
    e, f, t = oracle.cuh2_force(g["positions"], g["Z"], np.array(g["box"]).reshape(3, 3), return_table=True)
    assert e == pytest.approx(g["energy"], rel=1e-13)
    assert np.abs(f - np.array(g["forces"])).max() < 1e-11
    assert int(t["row"][-1]) == g["n_directed_pairs"] == 11954


def test_cuh2_physics_suite(oracle):
    # CppCore/rgpot/fortran/test_cuh2.f90:71-137 in the multi-image box (7.23 A < 2 * 6.1 A)
    pos, Z, box = workloads.cu_slab_with_h2()
    e, f0, t = oracle.cuh2_force(pos, Z, box, return_table=True)
    assert int(t["row"][-1]) == 1874  # multi-image pairs present
    assert e == pytest.approx(-107.3299063315978, rel=1e-12)
    assert np.abs(f0.sum(0)).max() < 1e-9
    e_t, _ = oracle.cuh2_force(pos + 0.37, Z, box)
    assert abs(e_t - e) < 1e-8
    e_r, f_r = oracle.cuh2_force(pos[::-1], Z[::-1], box)
    assert abs(e_r - e) < 1e-8
    assert np.abs(f_r[::-1] - f0).max() < 1e-5
    delta = 1e-5
    for atom in [0, 4, 8, 13, 19, 26, 31, 32, 33]:
        for c in range(3):
            p = pos.copy()
            p[atom, c] += delta
            ep, _ = oracle.cuh2_force(p, Z, box)
            p[atom, c] -= 2 * delta
            em, _ = oracle.cuh2_force(p, Z, box)
            assert abs(-(ep - em) / (2 * delta) - f0[atom, c]) < 1e-5


def test_cuh2_errors(oracle):
    pos, Z, box = workloads.cu_slab_with_h2()
    bad = Z.copy()
    bad[2] = 26
    with pytest.raises(oracle.CuH2Error) as ei:
        oracle.cuh2_force(pos, bad, box)
    assert ei.value.status == 2
    assert "atom 3 has atomic number 26; this potential covers Cu (29) and H (1) only" in ei.value.message
    p2 = pos.copy()
    p2[1] = p2[0]
    with pytest.raises(oracle.CuH2Error) as ei:
        oracle.cuh2_force(p2, Z, box)
    assert ei.value.status == 3 and "two atoms occupy the same position" in ei.value.message
    with pytest.raises(oracle.CuH2Error) as ei:
        oracle.cuh2_force(pos, Z, np.zeros((3, 3)))
    assert ei.value.status == 1 and "not invertible" in ei.value.message
    e, f = oracle.cuh2_force(np.zeros((0, 3)), np.zeros(0, np.int32), box)
    assert e == 0.0 and f.shape == (0, 3)


def test_vesin_pairset_fixtures(oracle, golden):
    # pair sets produced by the vendored vesin itself (triclinic, atoms outside the cell, images)
    for name, c in golden("vesin_pairsets.json")["cases"].items():
        t = oracle.nlist(c["positions"], np.array(c["box"]).reshape(3, 3), c["cutoff"])
        assert np.array_equal(oracle.canonical_pairs(t), np.array(c["pairs"]).reshape(-1, 5)), name
        assert t["n_self_images"] == c["n_self_images"]


def test_triclinic_commensurate_shear(oracle):
    # a lattice-commensurate shear leaves a perfect crystal's energy unchanged
    a = 3.615
    cu = workloads._fcc(4, a)
    Z = np.full(len(cu), 29, dtype=np.int32)
    e0, _ = oracle.cuh2_force(cu, Z, np.diag([4 * a] * 3))
    e1, _ = oracle.cuh2_force(cu, Z, np.array([[4 * a, 0, 0], [a, 4 * a, 0], [0, 0, 4 * a]]))
    e2, _ = oracle.cuh2_force(cu, Z, np.array([[4 * a, 0, 0], [a, 4 * a, 0], [a / 2, a / 2, 4 * a]]))
    assert e0 == pytest.approx(-906.183991113513, rel=1e-12)
    assert e1 == pytest.approx(e0, rel=1e-11) and e2 == pytest.approx(e0, rel=1e-11)


def test_morse_pins(oracle, golden):
    # CppCore/tests/MorsePotTest.cc:62-117
    e2, f2, _ = oracle.morse_force([[10.0, 10.0, 10.0], [10.0 + 2.8970, 10.0, 10.0]], np.eye(3) * 40.0)
    assert e2 == pytest.approx(-0.71016446200272088, abs=1e-12)
    assert np.abs(f2).max() < 1e-12
    g = golden("lj13_cluster.json")
    pos, box = np.array(g["positions"]), np.array(g["box"]).reshape(3, 3)
    e, f, n = oracle.morse_force(pos, box)
    assert n == 78
    assert e == pytest.approx(7483.1166203812, rel=1e-9)
    assert f[0] == pytest.approx([-2056.95286055812, 333.678513150438, 1022.83140169994], rel=1e-9)
    assert np.sqrt((f ** 2).sum(1)).max() == pytest.approx(2480.28362300107, rel=1e-9)
    assert np.abs(f.sum(0)).max() < 1e-9


def test_zbl_pins(oracle):
    # CppCore/tests/ZBLPotTest.cc:36-142: LAMMPS pair_style zbl 2.0 2.5 values
    dimer = np.array([[9.40, 9.35, 9.30], [10.60, 10.65, 10.70]])
    e, f, n = oracle.zbl_force(dimer, [14, 79])
    assert n == 1
    assert e == pytest.approx(0.38537731, abs=1e-8)
    assert np.abs(f - np.array([[-2.37926, -2.57753, -2.7758], [2.37926, 2.57753, 2.7758]])).max() < 1e-5
    assert e == pytest.approx(0.38537730883, rel=1e-9)
    assert oracle.zbl_force(dimer, [79, 79])[0] == pytest.approx(0.961609764914, rel=1e-9)
    assert oracle.zbl_force(dimer, [14, 14])[0] == pytest.approx(0.167604056792, rel=1e-9)
    e_cut, f_cut, _ = oracle.zbl_force([[5.0, 5.0, 5.0], [7.5, 5.0, 5.0]], [14, 79])
    assert abs(e_cut) < 1e-12 and np.abs(f_cut).max() < 1e-12
    e_far, f_far, n_far = oracle.zbl_force([[5.0, 5.0, 5.0], [8.0, 5.0, 5.0]], [14, 79])
    assert n_far == 0 and e_far == 0.0 and not f_far.any()


# ------------------------------------------------------------------------------------- EAM-Al
AL_FCC = np.array([[7.975, 7.975, 7.975], [10.0, 10.0, 7.975], [10.0, 7.975, 10.0], [7.975, 10.0, 10.0]])
CUBE20 = np.eye(3) * 20.0


def test_eam_al_matches_the_eon_reference(oracle):
    # CppCore/tests/FortranPotsTest.cc:135-144 (eOn EAMAlTest fixture al_fcc in a 20 A cube)
    e, f = oracle.eam_al_force(AL_FCC, CUBE20)
    assert e == pytest.approx(-5.217864, rel=1e-4)
    assert np.sqrt((f ** 2).sum(1)).max() == pytest.approx(0.968647, rel=1e-3)
    assert np.abs(f.sum(0)).max() < 1e-6  # requireNoNetForce


def test_eam_al_physics_checks(oracle):
    # the checks of CppCore/rgpot/fortran/test_eam_al.f90, re-run against the restatement:
    # translation invariance, vanishing net force, analytic forces against central differences,
    # every pair term zero from the truncation radius on
    rng = np.random.default_rng(5)
    a = 4.05
    cell = np.eye(3) * 3 * a
    pos = (workloads._fcc(3, a) + rng.normal(0, 0.08, (108, 3)))
    e0, f0 = oracle.eam_al_force(pos, cell)
    e1, f1 = oracle.eam_al_force(pos + np.array([0.37, -1.21, 2.6]), cell)
    assert e1 == pytest.approx(e0, rel=1e-12) and np.abs(f1 - f0).max() < 1e-10
    assert np.abs(f0.sum(0)).max() < 1e-10
    h = 1e-5
    for (i, c) in [(0, 0), (17, 1), (55, 2)]:
        p, m = pos.copy(), pos.copy()
        p[i, c] += h
        m[i, c] -= h
        fd = -(oracle.eam_al_force(p, cell)[0] - oracle.eam_al_force(m, cell)[0]) / (2 * h)
        assert fd == pytest.approx(f0[i, c], abs=2e-6)
    rt = np.sqrt(0.9999) * 7.1
    for which in range(4):
        assert oracle.eam_al_scalar(which, rt) == 0.0 and oracle.eam_al_scalar(which, 7.1) == 0.0
        assert oracle.eam_al_scalar(which, np.nextafter(rt, 0.0)) != 0.0
    # F(rho) = g rho + sum c_k rho^k and its derivative (rgpot_eam_al.f90:189-213)
    g, c = -0.60647886749203, [4.4572157051836, 193.21775368064, -1173.6502684704, 4203.3500116196,
                               -8785.1680280827, 10632.994102532, -6921.0410455328, 1875.2698365752]
    for rho in (0.05, 0.4, 1.0):
        assert oracle.eam_al_scalar(4, rho) == pytest.approx(g * rho + sum(ck * rho ** (k + 1) for k, ck in enumerate(c)), rel=1e-13)
        assert oracle.eam_al_scalar(5, rho) == pytest.approx(g + sum((k + 1) * ck * rho ** k for k, ck in enumerate(c)), rel=1e-13)


# ------------------------------------------------------------------------------------- FeHe
FE_BCC = np.array([[7.13, 7.13, 7.13], [8.565, 8.565, 8.565], [7.13, 7.13, 10.0], [8.565, 8.565, 11.435],
                   [7.13, 10.0, 7.13], [8.565, 11.435, 8.565], [7.13, 10.0, 10.0], [8.565, 11.435, 11.435],
                   [10.0, 7.13, 7.13], [11.435, 8.565, 8.565], [10.0, 7.13, 10.0], [11.435, 8.565, 11.435],
                   [10.0, 10.0, 7.13], [11.435, 11.435, 8.565], [10.0, 10.0, 10.0], [11.435, 11.435, 11.435]])


def fehe_test_cell(n_side=4, rattle=0.03, seed=3):
    """test_fehe.f90: bcc iron, lattice 2.8553, two substitutional He on a <111>/2 bond, rattled."""
    a = 2.8553
    base = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    cells = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pos = ((cells[:, None, :] + base[None, :, :]).reshape(-1, 3)) * a
    Z = np.full(len(pos), 26, np.int32)
    Z[:2] = 2
    pos = pos + np.random.default_rng(seed).normal(0, rattle, pos.shape)
    return pos, Z, np.eye(3) * n_side * a


def test_fehe_matches_the_eon_reference(oracle):
    # CppCore/tests/FortranPotsTest.cc:146-155
    e, f = oracle.fehe_force(FE_BCC, np.full(16, 26, np.int32), CUBE20)
    assert e == pytest.approx(-43.959774, rel=1e-4)
    assert np.sqrt((f ** 2).sum(1)).max() == pytest.approx(1.245094, rel=1e-3)
    assert np.abs(f.sum(0)).max() < 1e-6


def test_fehe_physics_checks(oracle):
    # CppCore/rgpot/fortran/test_fehe.f90 re-run against the restatement
    pos, Z, cell = fehe_test_cell()
    e0, f0 = oracle.fehe_force(pos, Z, cell)
    e1, f1 = oracle.fehe_force(pos + 0.37, Z, cell)
    assert e1 == pytest.approx(e0, abs=1e-8) and np.abs(f0.sum(0)).max() < 1e-9
    h = 1e-5
    for (i, c) in [(0, 0), (0, 1), (1, 2), (2, 0), (8, 0), (32, 1), (42, 2), (16, 0), (99, 2)]:
        p, m = pos.copy(), pos.copy()
        p[i, c] += h
        m[i, c] -= h
        fd = -(oracle.fehe_force(p, Z, cell)[0] - oracle.fehe_force(m, Z, cell)[0]) / (2 * h)
        assert fd == pytest.approx(f0[i, c], abs=5e-6)
    assert np.abs(f0[:2]).max() > 1e-3  # the helium atoms feel force through all channels
    bad = Z.copy()
    bad[5] = 29
    with pytest.raises(oracle.FeHeError) as ei:
        oracle.fehe_force(pos, bad, cell)
    assert ei.value.status == 2 and "unsupported atomic number 29 on atom 6" in ei.value.message
    # every term is continuous across its branch points and vanishes beyond its cutoff
    for which, rc in [(0, 5.3), (2, 3.9028), (6, 4.2), (8, 4.10), (4, 5.4)]:
        assert oracle.fehe_scalar(which, np.nextafter(rc, np.inf)) == 0.0
    for r in (1.0, 2.05):
        assert oracle.fehe_scalar(0, np.nextafter(r, 0)) == pytest.approx(oracle.fehe_scalar(0, r), rel=2e-3)


# ------------------------------------------------------------------ result-cache key
def test_xxh3_restatement_matches_xxhash_vectors(oracle, golden):
    """oracle/xxh3_oracle.c against digests of libxxhash (tests/golden/make_xxh3_golden.py): every
    length class of XXH3_64bits, and Potential::cacheKey assembled as Potential.hpp:324-336 does."""
    g = golden("xxh3_vectors.json")

    def pattern(n):
        return bytes(((i * 131 + n * 31 + (i >> 8) * 7) & 0xFF) for i in range(n))

    assert len(g["digests"]) > 260
    for n, digest in g["digests"]:
        assert oracle.xxh3_64(pattern(n)) == int(digest), n
    for c in g["cache_keys"]:
        key = oracle.cache_key(np.array(c["positions"]).reshape(-1, 3), np.array(c["Z"], np.int32),
                               np.array(c["box"]).reshape(3, 3), c["pot_type"], c["params_key"])
        assert key == int(c["key"])
