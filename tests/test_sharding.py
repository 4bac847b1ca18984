"""Host logic of the multi-GPU batch path: partitioning and the host-side gather (gloo, CPU)."""
import os
import socket

import numpy as np
import pytest

from rgpot_b200 import sharding


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        for counts in (np.full(65536, 218), rng.integers(1, 400, size=1000), np.array([5]), np.zeros(0, int),
                       np.array([1000, 1, 1, 1, 1, 1])):
            ranges = sharding.shard_ranges(counts, world)
            assert len(ranges) == world
            assert ranges[0][0] == 0 and ranges[-1][1] == len(counts)
            for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
                assert e0 == b1 and b0 <= e0
            if len(counts) >= 64 * world:
                loads = [counts[b:e].sum() for b, e in ranges]
                assert max(loads) - min(loads) <= 2 * counts.max()
    assert sharding.shard_ranges(np.full(65536, 218), 8) == [(k * 8192, (k + 1) * 8192) for k in range(8)]
    with pytest.raises(ValueError):
        sharding.shard_ranges([1, 2], 0)


def test_local_slice_matches_offsets():
    offsets = np.concatenate([[0], np.cumsum([3, 5, 2, 7, 4, 4])])
    seen = []
    for r in range(3):
        b, e, a0, a1 = sharding.local_slice(offsets, r, 3)
        assert a0 == offsets[b] and a1 == offsets[e]
        seen += list(range(b, e))
    assert seen == list(range(6))


class _FakePot:
    """Stands in for an engine handle: the sharding logic must not care what evaluates."""

    def forceBatchPacked(self, offsets, positions, atom_types, boxes):
        offsets = np.asarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        energies = np.array([positions[offsets[s]:offsets[s + 1]].sum() for s in range(n)])
        return energies, -np.asarray(positions)


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)  # same batch on every rank
    counts = rng.integers(1, 30, size=37)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    pos = rng.normal(size=(int(offsets[-1]), 3))
    boxes = np.tile(np.eye(3).reshape(9), (len(counts), 1))
    b, e, en, fr = sharding.evaluate_sharded(_FakePot(), offsets, pos, None, boxes, rank, world)
    energies, forces = sharding.gather_to_root(en, fr, len(counts), offsets)
    if rank == 0:
        want = np.array([pos[offsets[s]:offsets[s + 1]].sum() for s in range(len(counts))])
        ok = np.allclose(energies, want) and np.array_equal(forces, -pos)
        open(os.path.join(tmpdir, "result"), "w").write("ok" if ok else "bad")
    else:
        assert energies is None and forces is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluate_and_gather_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "result").read() == "ok"


def test_gather_without_process_group_is_identity():
    e, f = sharding.gather_to_root(np.arange(3.0), np.ones((5, 3)), 3, [0, 1, 3, 5])
    assert list(e) == [0.0, 1.0, 2.0] and f.shape == (5, 3)
