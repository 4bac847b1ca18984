#!/usr/bin/env python
"""Device-resident throughput of the widened rows: Morse, ZBL, EAM-Al, FeHe, and the
vesin-shaped neighbor list.  One JSON line per workload (atom-evals/s, CUDA events on the launching
stream, 256 MB L2 flush between timed calls); `python tools/bench_widened.py > profiles/rNN_widened.json`.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgpot_b200  # noqa: E402
from rgpot_b200 import workloads  # noqa: E402


def fcc(cells, a, seed, rattle):
    pos = workloads._fcc(cells, a) + np.random.default_rng(seed).normal(0, rattle, (4 * cells ** 3, 3))
    return pos, np.eye(3) * cells * a


def bcc_fehe(n_side, frac_he, seed):
    a = 2.8553
    base = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    cells = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pos = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3) * a
    rng = np.random.default_rng(seed)
    pos = pos + rng.normal(0, 0.04, pos.shape)
    Z = np.full(len(pos), 26, np.int32)
    Z[rng.choice(len(pos), size=int(frac_he * len(pos)), replace=False)] = 2
    return pos, Z, np.eye(3) * n_side * a


def timed(pot, pos, Z, box, steps=5):
    dev = torch.device("cuda", 0)
    n = len(pos)
    d_pos = torch.from_numpy(np.ascontiguousarray(pos)).to(dev)
    d_Z = torch.from_numpy(np.ascontiguousarray(Z, dtype=np.int32)).to(dev)
    d_F = torch.empty((n, 3), dtype=torch.float64, device=dev)
    d_E = torch.empty(1, dtype=torch.float64, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        pot.force_device(d_pos, d_Z, box, d_F, d_E, ts.cuda_stream)
    ms = []
    for _ in range(steps):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        pot.force_device(d_pos, d_Z, box, d_F, d_E, ts.cuda_stream, sync=False)
        t1.record()
        torch.cuda.synchronize()
        ms.append(t0.elapsed_time(t1))
    assert pot.sync_status() == 0
    return float(np.median(ms)), float(d_E.item())


def cpu_baselines():
    """bench.py's cpu_baseline leg (the C restatements of the reference kernels; Morse / ZBL with the
    reference's O(N^2) pair scan, the EAMs over the restated vesin cell list) on a smaller sample of
    each recipe, one process, before this process touches CUDA.  atom-evals/s."""
    from bench import cpu_leg_rate
    res = {}
    pos, box = fcc(12, 4.1, 1, 0.08)
    res["morse_fcc_256k"] = {"value": cpu_leg_rate("morse_force", (pos, box), len(pos)), "sample": f"{len(pos)}-atom crystal"}
    rng = np.random.default_rng(2)
    n = 8000
    p2 = rng.random((n, 3)) * (n / 0.09) ** (1 / 3)
    z2 = rng.choice([1, 8, 14, 29, 79], n).astype(np.int32)
    res["zbl_gas_500k"] = {"value": cpu_leg_rate("zbl_force", (p2, z2), n), "sample": f"{n}-atom gas (O(N^2) scan)"}
    pos3, box3 = fcc(12, 4.05, 3, 0.08)
    res["eam_al_fcc_256k"] = {"value": cpu_leg_rate("eam_al_force", (pos3, box3), len(pos3)), "sample": f"{len(pos3)}-atom crystal"}
    pos4, z4, box4 = bcc_fehe(14, 0.01, 4)
    res["fehe_bcc_128k"] = {"value": cpu_leg_rate("fehe_force", (pos4, z4, box4), len(pos4)), "sample": f"{len(pos4)}-atom crystal"}
    for v in res.values():
        v.update({"unit": "atom-evals/s", "cores": 1, "kind": "port"})  # the C restatements of oracle/
    return res


def main():
    cpu = cpu_baselines()
    out = []
    # Morse: a loose Pt-like fcc crystal, 9.5 A cutoff (MorsePot defaults)
    pos, box = fcc(40, 4.1, 1, 0.08)
    ms, e = timed(rgpot_b200.MorsePot(), pos, np.full(len(pos), 78), box)
    out.append({"workload": "morse_fcc_256k", "kernel": "k_brick<PAIR>", "n_atoms": len(pos), "ms_per_step": ms,
                "value": len(pos) / ms * 1e3, "energy": e})
    # ZBL: a random gas, 2.5 A cutoff, five species, free boundaries
    rng = np.random.default_rng(2)
    n = 500_000
    pos = rng.random((n, 3)) * (n / 0.09) ** (1 / 3)
    Z = rng.choice([1, 8, 14, 29, 79], n).astype(np.int32)
    ms, e = timed(rgpot_b200.ZBLPot(), pos, Z, np.eye(3))
    out.append({"workload": "zbl_gas_500k", "kernel": "k_brick<PAIR>", "n_atoms": n, "ms_per_step": ms,
                "value": n / ms * 1e3, "energy": e})
    # EAM-Al: fcc aluminium 40^3 cells, 7.1 A cutoff
    pos, box = fcc(40, 4.05, 3, 0.08)
    ms, e = timed(rgpot_b200.EAMAlPot(), pos, np.full(len(pos), 13), box)
    out.append({"workload": "eam_al_fcc_256k", "kernel": "k_brick<AL_DENSITY> + k_brick<AL_FORCE>",
                "n_atoms": len(pos), "ms_per_step": ms, "value": len(pos) / ms * 1e3, "energy": e})
    # FeHe: bcc iron 40^3 cells with 1 % substitutional helium
    pos, Z, box = bcc_fehe(40, 0.01, 4)
    ms, e = timed(rgpot_b200.FeHePot(), pos, Z, box)
    out.append({"workload": "fehe_bcc_128k", "kernel": "k_brick<FEHE_DENSITY> + k_brick<FEHE_FORCE>", "n_atoms": len(pos),
                "ms_per_step": ms, "value": len(pos) / ms * 1e3, "energy": e})
    # neighbor list: the Cu slab, 6.1 A, full list, unsorted (host-buffer call: two passes + D2H of the list)
    pos, Zs, box = workloads.cu_slab_triclinic(cells=20)
    rgpot_b200.neighbor_list(pos, box, 6.1, full=True, sorted=False)
    t0 = time.perf_counter()
    nl = rgpot_b200.neighbor_list(pos, box, 6.1, full=True, sorted=False)
    dt = time.perf_counter() - t0
    out.append({"workload": "neighbor_list_cu_slab_32k_6.1A_full", "kernel": "k_emit_pairs<GEO_IMAGES> (two passes)",
                "n_atoms": len(pos), "entries": int(len(nl["distances"])), "seconds_host_call": dt,
                "value": len(pos) / dt, "note": "host call incl. H2D, two emit passes, D2H of 44 B per entry"})
    # result-cache keys (XXH3) of a 65 536 x 218-atom batch, device-resident: HBM-bound byte work
    from bench import cpu_leg_rate
    nsys, na = 65536, 218
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    d_pos = torch.rand((nsys * na, 3), dtype=torch.float64, device=dev, generator=g) * 20.0
    d_Z = torch.full((nsys * na,), 29, dtype=torch.int32, device=dev)
    d_boxes = torch.eye(3, dtype=torch.float64, device=dev).repeat(nsys, 1).contiguous() * 25.0
    d_keys = torch.zeros(nsys, dtype=torch.int64, device=dev)
    offsets = np.arange(nsys + 1, dtype=np.uint64) * na
    pot = rgpot_b200.CuH2Pot()
    ts = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        pot.cache_keys_device(offsets, d_pos, d_Z, d_boxes, d_keys, ts.cuda_stream)
    ms = []
    for _ in range(5):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        pot.cache_keys_device(offsets, d_pos, d_Z, d_boxes, d_keys, ts.cuda_stream)
        t1.record()
        torch.cuda.synchronize()
        ms.append(t0.elapsed_time(t1))
    nbytes = nsys * (na * 28 + 72)
    hp, hz, hb = d_pos[:na].cpu().numpy(), d_Z[:na].cpu().numpy(), d_boxes[:3].cpu().numpy()
    cpu_rate = cpu_leg_rate("cache_key", (hp, hz, hb, 1, pot.paramsKey()), na * 28 + 72, reps=2000)
    out.append({"workload": "cache_keys_65536x218", "kernel": "k_cache_keys", "bytes": nbytes,
                "ms_per_step": float(np.median(ms)), "GB_per_s": nbytes / float(np.median(ms)) / 1e6,
                "value": nsys * na / float(np.median(ms)) * 1e3,
                "cpu_baseline": {"value": cpu_rate / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                 "sample": "2000 keys of one 218-atom system through oracle/xxh3_oracle.c (ctypes call included)"}})
    for rec in out:
        rec["unit"] = "atom-evals/s"
        if rec["workload"] in cpu:
            rec["cpu_baseline"] = cpu[rec["workload"]]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
