#!/bin/bash
set -u
M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"
python tools/bench_tile.py --what slab --variants tile=0 --steps 1 > gpurun_out/plain_k_brick_eam.log 2>&1 &&
ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_brick -s 6 -c 2 \
    -o gpurun_out/r02_k_brick_eam python tools/bench_tile.py --what slab --variants tile=0 --steps 1 > gpurun_out/ncu_k_brick_eam.log 2>&1
tail -1 gpurun_out/ncu_k_brick_eam.log
