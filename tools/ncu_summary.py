#!/usr/bin/env python
"""Condense an .ncu-rep into the few counters DESIGN.md / bench.py argue with (per kernel)."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "avg_active_lanes",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__sass_average_branch_targets_threads_uniform.pct": "branch_uniform_pct",
    "sm__sass_thread_inst_executed_op_fp64_pred_on.sum": "fp64_thread_inst",
}


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                try:
                    rec[WANT[h]] = float(v)
                except ValueError:
                    rec[WANT[h]] = v
                rec[WANT[h] + "_unit"] = u
        if "dfma" in rec:
            rec["fp64_flops"] = 2 * rec.get("dfma", 0) + rec.get("dadd", 0) + rec.get("dmul", 0)
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
