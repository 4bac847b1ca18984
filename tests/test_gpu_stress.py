"""Stress of the host pipeline and of handle re-entrancy (caps(): PerInstance -- Potential.hpp / pot_caps.hpp:20-29:
one handle per thread may be driven concurrently): batches of random sizes through the pageable pointer entry, the
packed entry and the single-system entries, interleaved, from three threads on three handles at once; every result must
be bit-identical to the one a quiet handle produced for the same input.  A race in the chunk pipeline, the staging
ring, the helper teams or the mapped single-system blocks shows up here as a mismatch or a hang (pytest-timeout).
"""
import threading

import numpy as np
import pytest

import rgpot_b200
from rgpot_b200 import workloads

pytestmark = pytest.mark.gpu


def _reference_results(pot, offsets, pos, Z, boxes, sizes):
    out = {}
    for n_cfg in sorted(set(sizes)):
        a = int(offsets[n_cfg])
        out[n_cfg] = pot.forceBatchPacked(offsets[: n_cfg + 1], pos[:a], Z[:a], boxes[:n_cfg])
    return out


@pytest.mark.timeout(300)
def test_three_threads_three_handles_random_batches():
    n_max = 2600
    offsets, pos, Z, boxes = workloads.cuh2_batch(n_max, seed=4242)
    rng = np.random.default_rng(7)
    sizes = [int(x) for x in rng.choice([1, 2, 3, 17, 64, 255, 256, 257, 1000, 2047, 2048, 2049, n_max], size=40)]
    quiet = rgpot_b200.CuH2Pot()
    ref = _reference_results(quiet, offsets, pos, Z, boxes, sizes)
    slab = workloads.cuh2_slab218()
    e_slab, f_slab, _ = quiet(*slab)
    pin4 = workloads.cuh2_pin4()
    e_pin, f_pin, _ = quiet(*pin4)
    big = workloads.cu_slab_triclinic(cells=8, n_h2_side=2)
    e_big, f_big, _ = quiet(*big)
    errors = []

    def worker(tid):
        try:
            pot = rgpot_b200.CuH2Pot(host_threads=3 + tid)
            order = list(sizes)
            np.random.default_rng(100 + tid).shuffle(order)
            for k, n_cfg in enumerate(order):
                a = int(offsets[n_cfg])
                E_ref, F_ref = ref[n_cfg]
                if k % 3 == 0:  # the ForceBatch shape: separately allocated pageable systems
                    systems = [(pos[int(offsets[s]):int(offsets[s + 1])].copy(), Z[int(offsets[s]):int(offsets[s + 1])].copy(),
                                boxes[s].reshape(3, 3).copy()) for s in range(n_cfg)]
                    res = pot.forceBatch(systems)
                    E = np.array([r[0] for r in res])
                    F = np.concatenate([r[1] for r in res])
                elif k % 3 == 1:  # packed pageable arrays
                    E, F = pot.forceBatchPacked(offsets[: n_cfg + 1], pos[:a].copy(), Z[:a].copy(), boxes[:n_cfg].copy())
                else:  # packed page-locked arrays
                    h_pos = rgpot_b200.pinned_empty((a, 3))
                    h_F = rgpot_b200.pinned_empty((a, 3))
                    h_E = rgpot_b200.pinned_empty((n_cfg,))
                    h_pos[:] = pos[:a]
                    E, F = pot.forceBatchPacked(offsets[: n_cfg + 1], h_pos, Z[:a], boxes[:n_cfg], h_F, h_E)
                    E, F = np.array(E), np.array(F)
                if not (np.array_equal(E, E_ref) and np.array_equal(F, F_ref)):
                    errors.append((tid, k, n_cfg, "batch mismatch", float(np.abs(F - F_ref).max())))
                # single-system entries in between (mapped blocks, cooperative kernel, cell-list path)
                e, f, _ = pot(*slab)
                if not (e == e_slab and np.array_equal(f, f_slab)):
                    errors.append((tid, k, "slab218 mismatch"))
                e, f, _ = pot(*pin4)
                if not (e == e_pin and np.array_equal(f, f_pin)):
                    errors.append((tid, k, "pin4 mismatch"))
                if k % 5 == 0:
                    e, f, _ = pot(*big)
                    if not (e == e_big and np.array_equal(f, f_big)):
                        errors.append((tid, k, "cell-list slab mismatch"))
        except Exception as exc:  # noqa: BLE001
            errors.append((tid, "exception", repr(exc)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
