"""Sharding of a batch of independent configurations over the GPUs of one box.

Systems of a ``ForceBatch`` share nothing (CppCore/rgpot/ForceStructs.hpp:42-48,
Potential.hpp:231-233), so the batch is cut into contiguous ranges of the batch index balanced by
atom count, one range per rank (one process per GPU).  There is NO data-path collective: each
rank evaluates its own range; the only exchange is a host-side gather of the per-system energies
(and forces, if the caller wants them on one rank).
"""
from __future__ import annotations

import numpy as np


def shard_ranges(atoms_per_system, world_size: int):
    """Contiguous [begin, end) ranges of the batch index, one per rank, balanced by sum(atoms).

    Every system lands in exactly one range; ranges may be empty when there are fewer systems than
    ranks.  Deterministic, so all ranks compute the same cut without talking to each other.
    """
    counts = np.asarray(atoms_per_system, dtype=np.int64)
    n = len(counts)
    if world_size < 1:
        raise ValueError("world_size must be positive")
    prefix = np.concatenate([[0], np.cumsum(counts)])
    total = int(prefix[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        # first index whose prefix reaches the target, never moving backwards
        k = int(np.searchsorted(prefix, target, side="left"))
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def local_slice(offsets, rank: int, world_size: int):
    """(sys_begin, sys_end, atom_begin, atom_end) of this rank for a packed batch."""
    offsets = np.asarray(offsets, dtype=np.int64)
    b, e = shard_ranges(np.diff(offsets), world_size)[rank]
    return b, e, int(offsets[b]), int(offsets[e])


def evaluate_sharded(pot, offsets, positions, atom_types, boxes, rank: int, world_size: int):
    """Evaluate this rank's range of a packed batch with ``pot.forceBatchPacked``.

    Returns (sys_begin, sys_end, energies_local, forces_local).
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    b, e, a0, a1 = local_slice(offsets, rank, world_size)
    if e == b:
        return b, e, np.zeros(0), np.zeros((0, 3))
    rel = (offsets[b:e + 1] - a0).astype(np.uint64)
    energies, forces = pot.forceBatchPacked(rel, positions[a0:a1],
                                            None if atom_types is None else atom_types[a0:a1],
                                            np.asarray(boxes).reshape(-1, 9)[b:e])
    return b, e, energies, forces


def gather_to_root(local_energies, local_forces, n_systems: int, offsets, root: int = 0):
    """Host-side gather of the per-rank results onto ``root`` (torch.distributed, any backend).

    Returns (energies, forces) on root, (None, None) elsewhere.  Works without an initialised
    process group (single process): the local result is the result.
    """
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local_energies), np.asarray(local_forces)
    world = dist.get_world_size()
    rank = dist.get_rank()
    payload = (np.asarray(local_energies), np.asarray(local_forces))
    gathered = [None] * world if rank == root else None
    dist.gather_object(payload, gathered, dst=root)
    if rank != root:
        return None, None
    energies = np.concatenate([g[0] for g in gathered])
    forces = np.concatenate([g[1].reshape(-1, 3) for g in gathered])
    assert len(energies) == n_systems and len(forces) == int(np.asarray(offsets)[-1])
    return energies, forces
