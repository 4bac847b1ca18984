#!/bin/bash
# ncu capture of the half-list batch kernel (round 2, third pass): plain run first, then under ncu;
# summaries into profiles/ by tools/ncu_summary.py / tools/ncu_lines.py afterwards.
set -u
M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"
python tools/bench_batch_dev.py --configs 16384 --steps 1 > gpurun_out/plain_k_eam_small_half.log 2>&1 &&
ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_eam_small -s 3 -c 1 \
    -o gpurun_out/r02c_k_eam_small python tools/bench_batch_dev.py --configs 16384 --steps 1 > gpurun_out/ncu_k_eam_small_half.log 2>&1
tail -1 gpurun_out/ncu_k_eam_small_half.log
python bench.py --workload cuh2_batch --steps 2 --warmup 3 --configs 16384 --no-extras --cpu-seconds 1 > gpurun_out/plain_launch_half.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_cuh2_batch.csv \
    python bench.py --workload cuh2_batch --steps 2 --warmup 3 --configs 16384 --no-extras --cpu-seconds 1 > gpurun_out/ncu_launch_half.log 2>&1
# the small single-system kernels: launch list and stage stamps
python tools/small_calls.py --reps 200 > gpurun_out/plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_small.csv \
    python tools/small_calls.py --reps 3 > gpurun_out/ncu_small.log 2>&1
RGPOT_B200_ENGINE=$PWD/variants/lib_stamps.so python tools/small_calls.py --stamps --reps 300 > gpurun_out/r02c_small_stamps.jsonl 2>&1
./tools/latency_probe2 > gpurun_out/r02c_latency_probe2.jsonl 2>&1
ls -la gpurun_out/r02c_*
