#!/usr/bin/env python
"""Threads per atom of the brick kernels (option brick_tpa) on the large-system workloads: ms per step for 1 / 2 / 4."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_widened as bw  # noqa: E402
import rgpot_b200  # noqa: E402
from rgpot_b200 import workloads  # noqa: E402


def main():
    cases = []
    pos, box = bw.fcc(40, 4.05, 3, 0.08)
    cases.append(("eam_al_fcc_256k", rgpot_b200.EAMAlPot, pos, np.full(len(pos), 13), box))
    pos, Z, box = bw.bcc_fehe(40, 0.01, 4)
    cases.append(("fehe_bcc_128k", rgpot_b200.FeHePot, pos, Z, box))
    pos, box = bw.fcc(40, 4.1, 1, 0.08)
    cases.append(("morse_fcc_256k", rgpot_b200.MorsePot, pos, np.full(len(pos), 78), box))
    pos, Z, box = workloads.cu_slab_triclinic(cells=40)
    cases.append(("cu_slab_256k", rgpot_b200.CuH2Pot, pos, Z, box))
    pos, Z, box = workloads.cu_slab_triclinic(cells=20)
    cases.append(("cu_slab_32k", rgpot_b200.CuH2Pot, pos, Z, box))
    pos, Z, box = workloads.lj_fluid(1000000)
    cases.append(("lj_fluid_1m", lambda: rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0)), pos, Z, box))
    for name, make, pos, Z, box in cases:
        for tpa in (0, 1, 2, 4):
            pot = make()
            pot.set_option("brick_tpa", tpa)
            ms, e = bw.timed(pot, pos, Z, box, steps=6)
            print(json.dumps({"workload": name, "brick_tpa": tpa, "ms": ms, "atom_evals_per_s": len(pos) / ms * 1e3, "energy": e}), flush=True)


if __name__ == "__main__":
    main()
