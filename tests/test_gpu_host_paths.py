"""The reference-shaped host entry points on a B200: ForceBatch with separately allocated pageable
buffers (CppCore/rgpot/Potential.hpp:225-240, ForceStructs.hpp:40-54), pageable single systems
(Potential.hpp:165-211), the one-launch path of single small systems, homogeneous batches that share
one copy of the atomic numbers, and batches cut across several devices by ONE engine call.

All of them must give the bits the packed / pinned / single-device entries give: the kernels and the
summation orders are the same, only the way the bytes travel differs.
"""
import numpy as np
import pytest

import rgpot_b200
from rgpot_b200 import workloads

pytestmark = pytest.mark.gpu


def _n_gpus():
    from rgpot_b200 import _lib
    return int(_lib.lib().rgpot_b200_available())


def _systems(n_cfg, seed=7):
    offsets, pos, Z, boxes = workloads.cuh2_batch(n_cfg, seed=seed)
    systems = []
    for s in range(n_cfg):
        a, b = int(offsets[s]), int(offsets[s + 1])
        # separate allocations, like a caller's per-image buffers
        systems.append((pos[a:b].copy(), Z[a:b].copy(), boxes[s].reshape(3, 3).copy()))
    return offsets, pos, Z, boxes, systems


def test_force_batch_pointer_entry_matches_packed(oracle):
    n_cfg = 3000  # several pipeline chunks
    offsets, pos, Z, boxes, systems = _systems(n_cfg)
    pot = rgpot_b200.CuH2Pot()
    e_packed, f_packed = pot.forceBatchPacked(offsets, pos, Z, boxes)  # pageable packed arrays
    fb = pot.prepareForceBatch(systems)
    assert pot.forceBatchPrepared(fb) == 0
    assert np.array_equal(fb["energies"], e_packed)
    assert np.array_equal(np.concatenate(fb["forces"]), f_packed)
    assert (fb["statuses"] == 0).all()
    # pinned packed arrays: the DMA-from-user-memory route, same bits
    h_pos = rgpot_b200.pinned_empty(pos.shape)
    h_Z = rgpot_b200.pinned_empty(Z.shape, np.int32)
    h_boxes = rgpot_b200.pinned_empty(boxes.shape)
    h_F = rgpot_b200.pinned_empty(pos.shape)
    h_E = rgpot_b200.pinned_empty((n_cfg,))
    h_pos[:], h_Z[:], h_boxes[:] = pos, Z, boxes
    pot.forceBatchPacked(offsets, h_pos, h_Z, h_boxes, h_F, h_E)
    assert np.array_equal(h_E, e_packed) and np.array_equal(h_F, f_packed)
    # against the oracle on a sample
    for s in (0, 1499, 2999):
        a, b = int(offsets[s]), int(offsets[s + 1])
        e_o, f_o = oracle.cuh2_force(pos[a:b], Z[a:b], boxes[s].reshape(3, 3))
        assert abs(e_packed[s] - e_o) <= 1e-10 * abs(e_o)
        assert np.abs(f_packed[a:b] - f_o).max() <= 1e-9


def test_shared_species_copy_is_only_used_when_every_system_matches(oracle):
    """A homogeneous batch uploads ONE copy of the atomic numbers; a batch with one odd member must not."""
    n_cfg = 2500
    offsets, pos, Z, boxes, systems = _systems(n_cfg, seed=11)
    pot = rgpot_b200.CuH2Pot()
    fb = pot.prepareForceBatch(systems)
    pot.forceBatchPrepared(fb)
    e_ref = fb["energies"].copy()
    f_ref = [f.copy() for f in fb["forces"]]
    # swap the species of two atoms of ONE system in the middle of a chunk (a Cu atom becomes H ...)
    odd = 1300
    p, z, b = systems[odd]
    z2 = z.copy()
    z2[5], z2[-1] = z2[-1], z2[5]
    assert (z2 != z).any()
    systems2 = list(systems)
    systems2[odd] = (p, z2, b)
    fb2 = pot.prepareForceBatch(systems2)
    pot.forceBatchPrepared(fb2)
    e_o, f_o = oracle.cuh2_force(p, z2, b)
    assert abs(fb2["energies"][odd] - e_o) <= 1e-10 * abs(e_o)
    assert np.abs(fb2["forces"][odd] - f_o).max() <= 1e-9
    assert abs(fb2["energies"][odd] - e_ref[odd]) > 1e-6  # it really is another system
    for s in (0, odd - 1, odd + 1, n_cfg - 1):  # its neighbours are untouched
        assert fb2["energies"][s] == e_ref[s]
        assert np.array_equal(fb2["forces"][s], f_ref[s])
    # the same through the packed entry
    Z2 = Z.copy()
    a = int(offsets[odd])
    Z2[a:a + len(z2)] = z2
    e_p, f_p = pot.forceBatchPacked(offsets, pos, Z2, boxes)
    assert np.array_equal(e_p, fb2["energies"])


def test_statuses_and_messages_across_chunks():
    n_cfg = 2600
    offsets, pos, Z, boxes, systems = _systems(n_cfg, seed=3)
    bad = 2100
    p, z, b = systems[bad]
    zb = z.copy()
    zb[7] = 79  # gold: status 2, one-based atom index in the message (rgpot_cuh2.f90:481-486)
    systems[bad] = (p, zb, b)
    pot = rgpot_b200.CuH2Pot()
    fb = pot.prepareForceBatch(systems)
    st = pot.forceBatchPrepared(fb, check=False)
    assert st == 2
    assert fb["statuses"][bad] == 2 and (np.delete(fb["statuses"], bad) == 0).all()
    assert fb["energies"][bad] == 0.0 and not fb["forces"][bad].any()
    assert "system 2100: rgpot_cuh2: atom 8 has atomic number 79" in pot._last_error()
    assert fb["energies"][bad - 1] != 0.0


def test_pageable_single_system_matches_pinned():
    """Large single systems from pageable arrays travel through the pinned ring: same bits."""
    pos, Z, box = workloads.lj_fluid(300_000)
    lj = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
    e_page, f_page, _ = lj(pos, Z, box)  # numpy (pageable) in, numpy out
    h_pos = rgpot_b200.pinned_empty(pos.shape)
    h_F = rgpot_b200.pinned_empty(pos.shape)
    h_pos[:] = pos
    e_pin, f_pin, _ = lj(h_pos, Z, box, out=h_F)
    assert e_page == e_pin
    assert np.array_equal(f_page, f_pin)
    p, z, b = workloads.cu_slab_triclinic(cells=24, n_h2_side=8)
    cu = rgpot_b200.CuH2Pot()
    e1, f1, _ = cu(p, z, b)
    hp = rgpot_b200.pinned_empty(p.shape)
    hz = rgpot_b200.pinned_empty(z.shape, np.int32)
    hf = rgpot_b200.pinned_empty(p.shape)
    hp[:], hz[:] = p, z
    e2, f2, _ = cu(hp, hz, b, out=hf)
    assert e1 == e2 and np.array_equal(f1, f2)


def test_small_single_fast_path(oracle, golden):
    """configs[0] / configs[1]: one launch, inputs and outputs in mapped page-locked memory."""
    c, zc, cell = workloads.lj13_cluster()
    ljc = rgpot_b200.LJClusterPot()
    l0 = ljc.launch_count()
    e, f, _ = ljc(c, zc, cell)
    assert ljc.launch_count() - l0 == 1
    assert abs(e - (-39.9653513626179)) <= 1e-9 * 39.97
    p4, z4, b4 = workloads.cuh2_pin4()
    cu = rgpot_b200.CuH2Pot()
    l0 = cu.launch_count()
    e4, f4, _ = cu(p4, z4, b4)  # (a 7.2 A cell: several images per pair -- may take the general path)
    e_o, f_o = oracle.cuh2_force(p4, z4, b4)
    assert abs(e4 - e_o) <= 1e-10 * abs(e_o) and np.abs(f4 - f_o).max() <= 1e-9
    p, z, b = workloads.cuh2_slab218()
    l0 = cu.launch_count()
    e218, f218, _ = cu(p, z, b)
    assert cu.launch_count() - l0 == 1  # ONE cooperative launch over fourteen blocks (single.cuh)
    e_o, f_o = oracle.cuh2_force(p, z, b)
    assert abs(e218 - e_o) <= 1e-10 * abs(e_o) and np.abs(f218 - f_o).max() <= 1e-9
    cu.set_option("one", 0)  # the batch kernel on one block: one launch, same tolerance
    l0 = cu.launch_count()
    e1b, f1b, _ = cu(p, z, b)
    assert cu.launch_count() - l0 == 1
    assert abs(e1b - e_o) <= 1e-10 * abs(e_o) and np.abs(f1b - f_o).max() <= 1e-9
    cu.set_option("one", 1)
    # a repeated call reuses the block; a different size regrows it; errors still come back
    for n_side in (2, 3):
        pp, zz, bb = workloads.cu_slab_with_h2(cells=n_side)
        ee, ff, _ = cu(pp, zz, bb)
        e_o, f_o = oracle.cuh2_force(pp, zz, bb)
        assert abs(ee - e_o) <= 1e-10 * abs(e_o) and np.abs(ff - f_o).max() <= 1e-9
    zbad = z.copy()
    zbad[3] = 47
    with pytest.raises(rgpot_b200.PotentialError, match=r"status 2\): rgpot_cuh2: atom 4 has atomic number 47"):
        cu(p, zbad, b)
    pnan = p.copy()
    pnan[10, 1] = np.nan
    with pytest.raises(rgpot_b200.PotentialError, match="status 1"):
        cu(pnan, z, b)
    e_again, f_again, _ = cu(p, z, b)
    assert e_again == e218 and np.array_equal(f_again, f218)


def test_fehe_and_al_failures_have_their_own_status():
    """a non-finite coordinate / singular box under FeHePot and EAMAlPot is status 1 with
    zeroed outputs (rgpot_fehe.f90:595 propagates the table's status), not a silent garbage result."""
    a = 2.87
    g = np.arange(6, dtype=float)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    pos = a * np.concatenate([corner, corner + 0.5])
    Z = np.full(len(pos), 26, dtype=np.int32)
    Z[::37] = 2
    box = np.eye(3) * 6 * a
    for cls, z in ((rgpot_b200.FeHePot, Z), (rgpot_b200.EAMAlPot, None)):
        pot = cls()
        e, f, _ = pot(pos, z, box)
        assert np.isfinite(e) and np.isfinite(f).all()
        bad = pos.copy()
        bad[17, 2] = np.inf
        with pytest.raises(rgpot_b200.PotentialError, match="status 1"):
            pot(bad, z, box)
        sing = box.copy()
        sing[2] = sing[1]
        with pytest.raises(rgpot_b200.PotentialError, match="status 1"):
            pot(pos, z, sing)


@pytest.mark.skipif("_n_gpus() < 2")
def test_one_call_cuts_the_batch_across_devices(oracle):
    """RgpotB200Config.devices: a single force_batch call drives every listed GPU; results are the
    single-device bits (same kernel per system), statuses and messages keep batch-wide indices."""
    n_dev = min(_n_gpus(), 8)
    n_cfg = 4096 + 37
    offsets, pos, Z, boxes, systems = _systems(n_cfg, seed=5)
    one = rgpot_b200.CuH2Pot(device=0)
    many = rgpot_b200.CuH2Pot(devices=list(range(n_dev)))
    fb1 = one.prepareForceBatch(systems)
    one.forceBatchPrepared(fb1)
    fbn = many.prepareForceBatch(systems)
    many.forceBatchPrepared(fbn)
    assert np.array_equal(fb1["energies"], fbn["energies"])
    assert all(np.array_equal(a, b) for a, b in zip(fb1["forces"], fbn["forces"]))
    e_p, f_p = many.forceBatchPacked(offsets, pos, Z, boxes)
    assert np.array_equal(e_p, fb1["energies"]) and np.array_equal(f_p, np.concatenate(fb1["forces"]))
    # heterogeneous sizes: the cut is by atoms, every system still lands somewhere
    rng = np.random.default_rng(0)
    mixed = []
    for s in range(600):
        p, z, b = systems[s]
        keep = int(rng.integers(150, 219))
        mixed.append((p[:keep].copy(), z[:keep].copy(), b))
    r1 = one.forceBatch(mixed)
    rn = many.forceBatch(mixed)
    for (e1, f1, _), (en, fn, _) in zip(r1, rn):
        assert e1 == en and np.array_equal(f1, fn)
    # a failing member on a later device: batch-wide index in the message
    bad = n_cfg - 10
    p, z, b = systems[bad]
    zb = z.copy()
    zb[0] = 8
    systems[bad] = (p, zb, b)
    fbb = many.prepareForceBatch(systems)
    assert many.forceBatchPrepared(fbb, check=False) == 2
    assert fbb["statuses"][bad] == 2 and fbb["statuses"].sum() == 2
    assert f"system {bad}: rgpot_cuh2: atom 1 has atomic number 8" in many._last_error()
    # single systems on a multi-device handle run on its first device
    p, z, b = workloads.cuh2_slab218()
    e_a, f_a, _ = many(p, z, b)
    e_b, f_b, _ = one(p, z, b)
    assert e_a == e_b and np.array_equal(f_a, f_b)
