"""The C restatement against the reference's own compiled objects (oracle/_ref/libref.so).

Skipped where libref.so is absent; it is built only where /root/reference exists.
"""
import numpy as np
import pytest

from rgpot_b200 import workloads


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    oracle.ref_lib()
    return oracle


@pytest.mark.parametrize("n,L,rc,seed", [(64, 9.0, 2.5, 1), (200, 7.0, 3.0, 2), (50, 5.0, 4.0, 3),
                                         (300, 30.0, 15.0, 4)])
def test_lj_restatement_is_bit_exact(ref, n, L, rc, seed):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3)) * L * 1.3 - 0.15 * L
    box = np.diag([L, 1.07 * L, 0.93 * L])
    e_r, f_r = ref.ref_lj_force(pos, box, 1.3, rc, 0.9)
    e_o, f_o, _ = ref.lj_force(pos, box, 1.3, rc, 0.9)
    assert e_o == e_r
    assert np.array_equal(f_o, f_r)
    e_r, f_r = ref.ref_lj_force(pos, box, 1.3, rc, 0.9, cluster=True)
    e_o, f_o, _ = ref.lj_force(pos, box, 1.3, rc, 0.9, cluster=True)
    assert e_o == e_r and np.array_equal(f_o, f_r)


def test_lj_second_sighting_uses_cached_list(ref):
    # PairListCache: 1st call scans, 2nd captures, 3rd replays the list; all agree with the scan
    rng = np.random.default_rng(11)
    pos = rng.random((120, 3)) * 12.0
    box = np.eye(3) * 12.0
    e_o, f_o, _ = ref.lj_force(pos, box, 1.0, 3.0, 1.0)
    for _ in range(3):
        e_r, f_r = ref.ref_lj_force(pos, box, 1.0, 3.0, 1.0)
        assert e_r == pytest.approx(e_o, rel=1e-13)
        assert np.abs(f_r - f_o).max() < 1e-9 * max(1.0, np.abs(f_o).max())


@pytest.mark.parametrize("kind", ["ortho", "triclinic", "tiny", "slab"])
def test_nlist_restatement_matches_vesin(ref, kind):
    rng = np.random.default_rng(5)
    if kind == "ortho":
        box, n = np.diag([13.0, 12.5, 14.0]), 150
    elif kind == "triclinic":
        box, n = np.array([[12.0, 0, 0], [3.0, 11.0, 0], [-2.0, 1.5, 13.0]]), 150
    elif kind == "tiny":
        box, n = np.array([[3.1, 0, 0], [0.4, 3.3, 0], [0.2, -0.3, 3.0]]), 5
    else:
        pos, Z, box = workloads.cu_slab_with_h2(3)
    if kind != "slab":
        pos = (rng.random((n, 3)) * 2.0 - 0.5) @ box
    t_ref = ref.nlist(pos, box, 6.1, use_ref=True)
    t_orc = ref.nlist(pos, box, 6.1, use_ref=False)
    assert np.array_equal(ref.canonical_pairs(t_ref), ref.canonical_pairs(t_orc))
    assert t_ref["n_self_images"] == t_orc["n_self_images"]


def test_cuh2_over_real_vesin(ref):
    pos, Z, box = workloads.cu_slab_triclinic(cells=5, n_h2_side=2)
    e_r, f_r = ref.cuh2_force(pos, Z, box, use_ref=True)
    e_o, f_o = ref.cuh2_force(pos, Z, box)
    assert e_o == pytest.approx(e_r, rel=1e-12)
    assert np.abs(f_o - f_r).max() < 1e-10


def test_params_key_matches_reference_scheme(ref):
    # LJClusterPotTest.cc:114-126
    lib = ref.ref_lib()
    k0 = lib.ref_lj_params_key(1.0, 15.0, 1.0, 0)
    assert k0 == lib.ref_lj_params_key(1.0, 15.0, 1.0, 1)
    assert k0 != lib.ref_lj_params_key(1.0, 20.0, 1.0, 0)


def test_morse_restatement_is_bit_exact(ref):
    rng = np.random.default_rng(9)
    for n, L, rc in [(80, 14.0, 5.0), (40, 8.0, 9.5), (150, 25.0, 9.5)]:
        pos = rng.random((n, 3)) * L * 1.2 - 0.1 * L
        box = np.diag([L, 1.1 * L, 0.9 * L])
        e_r, f_r = ref.ref_morse_force(pos, box, 0.7102, 1.6047, 2.8970, rc)
        e_o, f_o, _ = ref.morse_force(pos, box, 0.7102, 1.6047, 2.8970, rc)
        assert e_o == e_r and np.array_equal(f_o, f_r)


def test_zbl_restatement_is_bit_exact(ref):
    rng = np.random.default_rng(13)
    for n, L, species in [(60, 7.0, [14, 79]), (120, 9.0, [1, 8, 14, 29, 79]), (30, 4.0, [6])]:
        pos = rng.random((n, 3)) * L
        Z = rng.choice(species, n).astype(np.int32)
        e_r, f_r = ref.ref_zbl_force(pos, Z, np.eye(3) * 3.0)  # the cell is ignored
        e_o, f_o, _ = ref.zbl_force(pos, Z)
        assert e_o == e_r and np.array_equal(f_o, f_r)


def test_linear_lj_comparator_agrees_with_the_reference_object(ref):
    """The "reference components" comparator (vendored vesin half list + the LJPot.cc pair lambda) used
    for the full-size LJ check and as `cpu_baseline_linear`, against the reference's own LJPot object on
    a system its O(N^2) scan can afford: same pair count, energy to 1e-12, forces to 1e-10 (the two sum
    in different orders).  Box wider than two cutoffs, atoms outside the cell included."""
    pos, Z, box = workloads.lj_fluid(6000)
    pos = pos + np.random.default_rng(3).integers(-1, 2, size=pos.shape) @ box * (np.arange(len(pos)) % 9 == 0)[:, None]
    e_r, f_r = ref.ref_lj_force(pos, box, 1.0, 2.5, 1.0)
    e_c, f_c, m = ref.lj_vesin_force(pos, box, 1.0, 2.5, 1.0, n_threads=2)
    assert m == ref.lj_force(pos, box, 1.0, 2.5, 1.0)[2]
    assert abs(e_c - e_r) <= 1e-12 * abs(e_r)
    assert np.abs(f_c - f_r).max() <= 1e-10 * max(1.0, np.abs(f_r).max())
