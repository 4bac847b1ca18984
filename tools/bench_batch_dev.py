#!/usr/bin/env python
"""Device-resident timing of the CuH2 batch kernel (k_eam_small) for kernel tuning: CUDA events on the
launching stream, one JSON line.    python tools/bench_batch_dev.py [--configs 16384] [--steps 5]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rgpot_b200  # noqa: E402
from rgpot_b200 import workloads  # noqa: E402


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--options", default="", help="engine options, e.g. half=0")
    args = ap.parse_args()
    offsets, pos, Z, boxes = workloads.cuh2_batch(args.configs, seed=1000)
    n_atoms = int(offsets[-1])
    pot = rgpot_b200.CuH2Pot(device=0)
    for kv in args.options.split(","):
        if kv:
            k, v = kv.split("=")
            pot.set_option(k, int(v))
    dev = torch.device("cuda", 0)
    d_pos = torch.from_numpy(pos).to(dev)
    d_Z = torch.from_numpy(Z).to(dev)
    d_boxes = torch.from_numpy(np.ascontiguousarray(boxes)).to(dev)
    d_F = torch.empty((n_atoms, 3), dtype=torch.float64, device=dev)
    d_E = torch.empty(args.configs, dtype=torch.float64, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    for _ in range(3):
        pot.force_batch_device(offsets, d_pos, d_Z, d_boxes, d_F, d_E, ts.cuda_stream)
    assert pot.sync_status() == 0
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        pot.force_batch_device(offsets, d_pos, d_Z, d_boxes, d_F, d_E, ts.cuda_stream)
    t1.record()
    torch.cuda.synchronize()
    assert pot.sync_status() == 0
    ms = t0.elapsed_time(t1) / args.steps
    print(json.dumps({"what": "cuh2_batch_dev", "configs": args.configs, "ms": ms, "atom_evals_per_s": n_atoms / (ms * 1e-3),
                      "options": args.options, "e_sum": float(d_E.sum().item()), "f_abs_sum": float(d_F.abs().sum().item())}))


if __name__ == "__main__":
    main()
