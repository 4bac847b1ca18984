#!/usr/bin/env python
"""Timing of the reference-shaped host entries ("time -- then fix -- the reference-shaped
entries"): one JSON line per measurement.

    python tools/bench_host_paths.py [--configs 65536] [--threads 2,4,8,12] [--skip-big]

  batch     CuH2 batch through (a) the packed entry on pinned arrays, (b) the packed entry on pageable
            arrays, (c) rgpot_b200_force_batch on 65 536 separately allocated pageable systems -- the
            ForceBatch a C++ caller hands to Potential::forceBatch -- for several helper-thread counts,
            and, with more than one GPU, the same single call cut across all of them
  single    LJ 1M / Cu slab 256k from pinned and from pageable arrays
  latency   microseconds per call of the small single systems (BASELINE configs[0] / [1] and the
            reference's own PotBench.cc workloads)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rgpot_b200  # noqa: E402
from rgpot_b200 import _lib, workloads  # noqa: E402


def out(**kw):
    print(json.dumps(kw), flush=True)


def timeit(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def bench_batch(n_cfg, threads, n_gpus, reps=4):
    offsets, pos, Z, boxes = workloads.cuh2_batch(n_cfg, seed=1000)
    n_atoms = int(offsets[-1])
    systems = []
    for s in range(n_cfg):
        a, b = int(offsets[s]), int(offsets[s + 1])
        systems.append((pos[a:b].copy(), Z[a:b].copy(), boxes[s].reshape(3, 3).copy()))
    h_pos = rgpot_b200.pinned_empty(pos.shape)
    h_Z = rgpot_b200.pinned_empty(Z.shape, np.int32)
    h_boxes = rgpot_b200.pinned_empty(boxes.shape)
    h_F = rgpot_b200.pinned_empty(pos.shape)
    h_E = rgpot_b200.pinned_empty((n_cfg,))
    h_pos[:], h_Z[:], h_boxes[:] = pos, Z, boxes
    F_page = np.zeros_like(pos)
    E_page = np.zeros(n_cfg)
    ref_E = None
    for dev_list in ([0], list(range(n_gpus)) if n_gpus > 1 else None):
        if dev_list is None:
            continue
        for t in threads:
            pot = rgpot_b200.CuH2Pot(devices=dev_list, host_threads=t) if len(dev_list) > 1 else \
                rgpot_b200.CuH2Pot(device=0, host_threads=t)
            fb = pot.prepareForceBatch(systems)
            dt_pin = timeit(lambda: pot.forceBatchPacked(offsets, h_pos, h_Z, h_boxes, h_F, h_E), reps)
            if ref_E is None:
                ref_E = h_E.copy()
            assert np.array_equal(h_E, ref_E)
            dt_page = timeit(lambda: pot.forceBatchPacked(offsets, pos, Z, boxes, F_page, E_page), reps)
            assert np.array_equal(E_page, ref_E)
            dt_fb = timeit(lambda: pot.forceBatchPrepared(fb), reps)
            assert np.array_equal(fb["energies"], ref_E)
            out(what="cuh2_batch", configs=n_cfg, devices=len(dev_list), host_threads=t,
                cpus=len(os.sched_getaffinity(0)),
                packed_pinned_ms=dt_pin * 1e3, packed_pageable_ms=dt_page * 1e3, force_batch_ms=dt_fb * 1e3,
                packed_pinned=n_atoms / dt_pin, packed_pageable=n_atoms / dt_page, force_batch=n_atoms / dt_fb)
            del pot
        h2d = pos.nbytes + boxes.nbytes
        d2h = pos.nbytes + n_cfg * 8
        gbs = rgpot_b200.measure_copy_gbs(dev_list, h2d // len(dev_list), d2h // len(dev_list))
        out(what="copy_ceiling", devices=len(dev_list), gb_per_s=gbs,
            batch_ms_at_ceiling=(h2d + d2h) / gbs / 1e6 if gbs else None,
            atom_evals_per_s_at_ceiling=n_atoms / ((h2d + d2h) / gbs / 1e9) if gbs else None)


def bench_single():
    for name in ("lj_fluid_1m", "cuh2_slab_256k"):
        if name == "lj_fluid_1m":
            pos, Z, box = workloads.lj_fluid()
            pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0))
        else:
            pos, Z, box = workloads.cu_slab_triclinic()
            pot = rgpot_b200.CuH2Pot()
        n = len(pos)
        h_pos = rgpot_b200.pinned_empty(pos.shape)
        h_Z = rgpot_b200.pinned_empty(Z.shape, np.int32)
        h_F = rgpot_b200.pinned_empty(pos.shape)
        h_pos[:], h_Z[:] = pos, Z
        F = np.zeros_like(pos)
        dt_pin = timeit(lambda: pot(h_pos, h_Z, box, out=h_F), 5, warm=2)
        dt_page = timeit(lambda: pot(pos, Z, box, out=F), 5, warm=2)
        assert np.array_equal(F, h_F)
        out(what=name, n_atoms=n, pinned_ms=dt_pin * 1e3, pageable_ms=dt_page * 1e3, pinned=n / dt_pin,
            pageable=n / dt_page, ratio=dt_pin / dt_page)


def potbench_lj500():
    """CppCore/tests/PotBench.cc:128-193: 500 atoms, fcc 5^3 cells in a 32 A cube, rc = 15 (rc ~ L/2)."""
    a = 32.0 / 5.0
    pos = workloads._fcc(5, a) + workloads._sin_jitter(500, 0.05)
    return np.ascontiguousarray(pos), np.ones(500, np.int32), np.diag([32.0, 32.0, 32.0])


def potbench_cuh2_130():
    """CppCore/tests/PotBench.cc:262-321: 128 Cu (4 x 4 x 2 fcc cells) + 2 H, box 14.46 x 14.46 x 30."""
    a = 3.615
    g = [np.arange(4.0), np.arange(4.0), np.arange(2.0)]
    ix, iy, iz = np.meshgrid(*g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    cu = a * (corner[:, None, :] + workloads._FCC_BASIS[None, :, :]).reshape(-1, 3)
    top = cu[:, 2].max()
    h = np.array([[7.0 - 0.37, 7.0, top + 2.0], [7.0 + 0.37, 7.0, top + 2.0]])
    pos = np.concatenate([cu, h]) + workloads._sin_jitter(130, 0.02)
    Z = np.concatenate([np.full(128, 29, np.int32), np.ones(2, np.int32)])
    return np.ascontiguousarray(pos), Z, np.diag([14.46, 14.46, 30.0])


def bench_latency():
    cases = {
        "lj13_cluster (configs[0])": (rgpot_b200.LJClusterPot(), workloads.lj13_cluster()),
        "lj13 periodic LJPot": (rgpot_b200.LJPot(), workloads.lj13_cluster()),
        "cuh2_pin4 (configs[1])": (rgpot_b200.CuH2Pot(), workloads.cuh2_pin4()),
        "cuh2_slab218 (configs[1])": (rgpot_b200.CuH2Pot(), workloads.cuh2_slab218()),
        "PotBench LJ 500 atoms rc=15": (rgpot_b200.LJPot(), potbench_lj500()),
        "PotBench CuH2 128 Cu + 2 H": (rgpot_b200.CuH2Pot(), potbench_cuh2_130()),
    }
    lib = _lib.lib()
    import ctypes as C
    for name, (pot, (pos, Z, box)) in cases.items():
        pos = np.ascontiguousarray(pos)
        Z = np.ascontiguousarray(Z, dtype=np.int32)
        box = np.ascontiguousarray(box)
        F = np.zeros_like(pos)
        e, v = C.c_double(0.0), C.c_double(0.0)

        def raw():  # the C ABI call itself, no numpy marshalling
            return lib.rgpot_b200_force(pot._h, len(pos), pos.ctypes.data, Z.ctypes.data, F.ctypes.data,
                                        C.byref(e), C.byref(v), box.ctypes.data)
        assert raw() == 0
        l0 = pot.launch_count()
        dt = timeit(raw, 2000, warm=50)
        launches = (pot.launch_count() - l0) / 2050
        out(what="latency", case=name, n_atoms=len(pos), us_per_call=dt * 1e6, launches_per_call=launches,
            energy=e.value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, default=65536)
    ap.add_argument("--threads", default="4,8,12")
    ap.add_argument("--skip-big", action="store_true")
    ap.add_argument("--skip-batch", action="store_true")
    args = ap.parse_args()
    n_gpus = int(_lib.lib().rgpot_b200_available())
    out(what="host", cpus=len(os.sched_getaffinity(0)), gpus=n_gpus)
    bench_latency()
    if not args.skip_big:
        bench_single()
    if not args.skip_batch:
        bench_batch(args.configs, [int(t) for t in args.threads.split(",")], n_gpus)


if __name__ == "__main__":
    main()
