#!/usr/bin/env python
"""A few single-system calls of each small case (for an ncu launch list: kernel durations of the
one-launch path), and the per-call latency through the C-ABI when run plain.

    python tools/small_calls.py [--reps 2000]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import rgpot_b200  # noqa: E402
from rgpot_b200 import _lib, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=2000)
    ap.add_argument("--stamps", action="store_true", help="stage stamps (variants/lib_stamps.so build)")
    args = ap.parse_args()
    cases = [
        ("lj13", rgpot_b200.LJClusterPot(), workloads.lj13_cluster()),
        ("pin4", rgpot_b200.CuH2Pot(), workloads.cuh2_pin4()),
        ("slab218", rgpot_b200.CuH2Pot(), workloads.cuh2_slab218()),
        ("lj500", rgpot_b200.LJPot(), bench.potbench_lj500()),
        ("cuh2_130", rgpot_b200.CuH2Pot(), bench.potbench_cuh2_130()),
    ]
    lib = _lib.lib()
    for name, pot, (pos, Z, box) in cases:
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        Z = np.ascontiguousarray(Z, dtype=np.int32)
        box = np.ascontiguousarray(box, dtype=np.float64)
        F = np.zeros_like(pos)
        e, v = C.c_double(0.0), C.c_double(0.0)
        a = (pot._h, len(pos), pos.ctypes.data, Z.ctypes.data, F.ctypes.data, C.byref(e), C.byref(v), box.ctypes.data)
        for _ in range(20):
            assert lib.rgpot_b200_force(*a) == 0
        best = 1e30
        for _ in range(5 if args.reps > 50 else 1):
            t0 = time.perf_counter()
            for _ in range(args.reps):
                lib.rgpot_b200_force(*a)
            best = min(best, (time.perf_counter() - t0) / args.reps * 1e6)
        rec = {"case": name, "n": len(pos), "us_per_call": best, "energy": e.value}
        if args.stamps:
            acc = np.zeros(12)
            buf = (C.c_ulonglong * 12)()
            for _ in range(32):
                lib.rgpot_b200_force(*a)
                lib.rgpot_b200_debug_stamps(buf, 12)
                v = np.array(list(buf), dtype=np.float64)
                acc += v - v[0]
            rec["stamps_ns_from_start"] = [int(x) for x in acc / 32]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
