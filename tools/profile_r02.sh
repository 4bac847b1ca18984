#!/bin/bash
# ncu captures of the dominant kernels (round 2): each command first runs plain, then under ncu.
# Reports land in gpurun_out/; tools/ncu_summary.py / ncu_lines.py condense them into profiles/.
set -u
M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"
run() {  # name, kernel regex, skip, command...
  local name=$1 re=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --metrics $M --clock-control none --import-source on -k regex:$re -s $skip -c 1 \
      -o gpurun_out/r02_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
run k_eam_small k_eam_small 3 python tools/bench_batch_dev.py --configs 16384 --steps 1
run k_brick_lj 'k_brick' 3 python tools/bench_tile.py --what lj --variants tile=0 --steps 1
run k_brick_eam_density 'k_brick<\(int\)1' 3 python tools/bench_tile.py --what slab --variants tile=0 --steps 1
run k_brick_eam_force 'k_brick<\(int\)2' 3 python tools/bench_tile.py --what slab --variants tile=0 --steps 1
run k_dfma_peak k_dfma_peak 2 python -c "import rgpot_b200; print(rgpot_b200.measure_fp64_tflops(0))"
# launch lists of the three workloads (cold-cache, serialised: compare SHARES)
for w in cuh2_batch lj_fluid_1m cuh2_slab_256k; do
  python bench.py --workload $w --steps 2 --warmup 3 --configs 16384 --no-extras --cpu-seconds 1 > gpurun_out/plain_launch_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --configs 16384 --no-extras --cpu-seconds 1 > gpurun_out/ncu_launch_$w.log 2>&1
done
ls -la gpurun_out/r02_*
