#!/usr/bin/env python
"""Latency of the single-configuration cases (BASELINE configs[0] and [1]; SURVEY 8(d): "throughput is
launch latency; report microseconds per evaluation"): the public host-buffer call, pageable numpy
arrays in and out, one call at a time -- beside the reference CPU path for the same inputs.

    python tools/bench_latency.py > profiles/rNN_latency.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rgpot_b200  # noqa: E402
from bench import cpu_leg_rate  # noqa: E402  (bench.py's cpu_baseline leg)


def load(name):
    with open(os.path.join(ROOT, "tests", "golden", name)) as fh:
        return json.load(fh)


def per_call(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    out = []
    c13 = load("lj13_cluster.json")
    p13, b13 = np.array(c13["positions"]), np.array(c13["box"]).reshape(3, 3)
    z13 = np.ones(len(p13), np.int32)
    pin = load("cuh2_pin4.json")
    p4, z4, b4 = np.array(pin["positions"]), np.array(pin["Z"], np.int32), np.array(pin["box"]).reshape(3, 3)
    slab = load("cuh2_slab218.json")
    p218, z218, b218 = np.array(slab["positions"]), np.array(slab["Z"], np.int32), np.array(slab["box"]).reshape(3, 3)
    lj, ljc, cu = rgpot_b200.LJPot(), rgpot_b200.LJClusterPot(), rgpot_b200.CuH2Pot()
    cases = [
        ("lj13_cluster LJPot (configs[0])", 13, lambda: lj(p13, z13, b13), ("lj_force", (p13, b13, 1.0, 15.0, 1.0))),
        ("lj13_cluster LJClusterPot", 13, lambda: ljc(p13, z13, b13), ("lj_force", (p13, b13, 1.0, 15.0, 1.0, True))),
        ("cuh2_pin4 CuH2Pot (configs[1])", 4, lambda: cu(p4, z4, b4), ("cuh2_force", (p4, z4, b4))),
        ("cuh2_slab218 CuH2Pot (configs[1])", 218, lambda: cu(p218, z218, b218), ("cuh2_force", (p218, z218, b218))),
    ]
    for name, n, gpu, (cpu_fn, cpu_args) in cases:
        us_gpu = per_call(gpu, 2000)
        us_cpu = 1e6 / cpu_leg_rate(cpu_fn, cpu_args, 1, reps=200)
        out.append({"case": name, "n_atoms": n, "us_per_eval": us_gpu, "atom_evals_per_s": n / us_gpu * 1e6,
                    "cpu_us_per_eval": us_cpu, "cpu_kind": "oracle port / reference LJ objects through ctypes, 1 thread"})
    for rec in out:
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
