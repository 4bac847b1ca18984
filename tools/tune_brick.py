#!/usr/bin/env python
"""Time the large-system path (device-resident) under different brick-kernel settings.

    python tools/tune_brick.py lj|slab [brick:tpa:threads:fine:list:slack:nohandoff ...]   e.g.  lj 0 222:2:256 slab 0:0:0:0:0:0:1
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgpot_b200  # noqa: E402
from rgpot_b200 import workloads  # noqa: E402


def main():
    which = sys.argv[1]
    settings = sys.argv[2:] or ["0:0:0"]
    if which == "lj":
        pos, Z, box = workloads.lj_fluid(1_000_000)
        pot = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0), device=0)
    else:
        pos, Z, box = workloads.cu_slab_triclinic(cells=40)
        pot = rgpot_b200.CuH2Pot(device=0)
    n = len(pos)
    dev = torch.device("cuda", 0)
    d_pos, d_Z = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev)
    d_F = torch.empty((n, 3), dtype=torch.float64, device=dev)
    d_E = torch.empty(1, dtype=torch.float64, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref_F = None
    for st in settings:
        b, tpa, thr, fine, lst, slack, nohand = (int(x) for x in (st.split(":") + ["0"] * 7)[:7])
        pot.set_option("handoff", 0 if nohand else 1)
        pot.set_option("fine", fine)
        pot.set_option("brick_list", lst)
        pot.set_option("brick_slack", slack)
        pot.set_option("brick", b)
        pot.set_option("brick_tpa", tpa)
        pot.set_option("brick_threads", thr)
        for _ in range(3):
            pot.force_device(d_pos, d_Z, box, d_F, d_E, ts.cuda_stream)
        ms = []
        for _ in range(7):
            flush.zero_()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            pot.force_device(d_pos, d_Z, box, d_F, d_E, ts.cuda_stream, sync=False)
            t1.record()
            torch.cuda.synchronize()
            ms.append(t0.elapsed_time(t1))
        assert pot.sync_status() == 0
        F = d_F.cpu().numpy()
        if ref_F is None:
            ref_F = F
        print(f"{which} {st:>18}  min {min(ms):.4f} ms  med {float(np.median(ms)):.4f} ms  "
              f"{n / np.median(ms) / 1e6:.1f} M atom-evals/ms... E={float(d_E.item()):.12e}  "
              f"max|dF| vs first {float(np.abs(F - ref_F).max()):.2e}  slow_atoms {pot.stat('slow_atoms')} pool {pot.stat('pool_used')}", flush=True)
        prof = [pot.stat(f"prof{k}") for k in range(4)]
        if sum(prof) > 0:  # an -DRB_BRICK_TIMING build (RGPOT_B200_ENGINE=...): warp-clock shares of the phases
            tot = float(sum(prof))
            print("    warp clocks: setup %.1f%%  stage %.1f%%  search %.1f%%  physics %.1f%%" %
                  tuple(100.0 * p / tot for p in prof), flush=True)


if __name__ == "__main__":
    main()
