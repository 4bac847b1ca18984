#!/usr/bin/env python
"""Per-call wall times of the host-buffer single-system entry (pageable and page-locked arrays) on the 256k Cu slab
and the 1M LJ fluid, for several helper-thread counts: shows the PCIe link waking up (the first 10-20 calls after
an idle period run at a quarter of the speed) and what the staging of pageable arrays costs.

    python tools/diag_slab.py [--threads 1,4,14]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgpot_b200  # noqa: E402
from rgpot_b200 import workloads  # noqa: E402


def percall(pot, pos, Z, box, F, n=40):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        pot(pos, Z, box, out=F)
        ts.append((time.perf_counter() - t0) * 1e3)
    return ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", default="1,4,14")
    args = ap.parse_args()
    cases = [("cu_slab_256k", workloads.cu_slab_triclinic(cells=40), lambda t: rgpot_b200.CuH2Pot(device=0, host_threads=t)),
             ("lj_fluid_1m", workloads.lj_fluid(1000000),
              lambda t: rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0), device=0, host_threads=t))]
    for name, (pos, Z, box), make in cases:
        n = len(pos)
        h_pos = rgpot_b200.pinned_empty((n, 3))
        h_Z = rgpot_b200.pinned_empty((n,), np.int32)
        h_F = rgpot_b200.pinned_empty((n, 3))
        h_pos[:] = pos
        h_Z[:] = Z
        for t in [int(x) for x in args.threads.split(",")]:
            pot = make(t)
            F = np.zeros_like(pos)
            first = percall(pot, pos, Z, box, F, 40)
            page = percall(pot, pos, Z, box, F, 40)
            pin = percall(pot, h_pos, h_Z, box, h_F, 40)
            print(json.dumps({"case": name, "host_threads": t, "first_calls_ms": [round(x, 2) for x in first[:24]],
                              "pageable_ms_median": float(np.median(page)), "pinned_ms_median": float(np.median(pin)),
                              "ratio": float(np.median(pin) / np.median(page))}), flush=True)
            del pot


if __name__ == "__main__":
    main()
