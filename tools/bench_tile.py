#!/usr/bin/env python
"""Device-resident timing of the large single systems (LJ fluid 1M, Cu slab 256k) for kernel tuning:
CUDA events on the launching stream, L2 flushed between iterations, one JSON line per variant.

    python tools/bench_tile.py [--what lj,slab] [--variants "tile=0;tile=1;tile=1,tile_brick=844"]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rgpot_b200  # noqa: E402
from rgpot_b200 import workloads  # noqa: E402

ALL_KEYS = ("tile", "tile_group", "tile_brick", "tile_cap", "tile_wcap", "fine", "handoff", "brick", "brick_tpa")


def run(name, variants, steps):
    import torch
    if name == "lj":
        pos, Z, box = workloads.lj_fluid()
        pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
    else:
        pos, Z, box = workloads.cu_slab_triclinic()
        pot = rgpot_b200.CuH2Pot()
    n = len(pos)
    dev = torch.device("cuda", 0)
    d_pos = torch.from_numpy(pos).to(dev)
    d_Z = torch.from_numpy(Z).to(dev)
    d_F = torch.empty((n, 3), dtype=torch.float64, device=dev)
    d_E = torch.empty(1, dtype=torch.float64, device=dev)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ref = None
    for var in variants:
        opts = dict(kv.split("=") for kv in var.split(",") if kv)
        defaults = {"tile": 1, "handoff": 1}
        for key in ALL_KEYS:
            pot.set_option(key, int(opts.get(key, defaults.get(key, 0))))
        for _ in range(3):
            pot.force_device(d_pos, d_Z, box, d_F, d_E, stream)
        l0 = pot.launch_count()
        total_ms = 0.0
        for _ in range(steps):
            flush.zero_()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            pot.force_device(d_pos, d_Z, box, d_F, d_E, stream, sync=False)
            t1.record()
            torch.cuda.synchronize()
            total_ms += t0.elapsed_time(t1)
        assert pot.sync_status() == 0
        e = float(d_E.item())
        f = d_F.cpu().numpy()
        if ref is None:
            ref = (e, f)
        print(json.dumps({"what": name, "variant": var, "n_atoms": n, "ms": total_ms / steps,
                          "atom_evals_per_s": n * steps / (total_ms * 1e-3),
                          "launches": (pot.launch_count() - l0) / steps, "energy": e,
                          "de_rel": abs(e - ref[0]) / abs(ref[0]), "df_max": float(np.abs(f - ref[1]).max()),
                          "max_halo": pot.stat("max_halo"), "cap": pot.stat("cap_cand"), "slow": pot.stat("slow_atoms"),
                          "pool_used": pot.stat("pool_used")}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="lj,slab")
    ap.add_argument("--variants", default="tile=0;tile=1")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    for name in args.what.split(","):
        run(name, args.variants.split(";"), args.steps)


if __name__ == "__main__":
    main()
