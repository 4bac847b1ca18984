#!/bin/bash
# Run on the GPU box (via gpurun): launch lists + one full capture per dominant kernel.
# Every ncu run is preceded by the identical plain command exiting 0 (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
for W in cuh2_batch lj_fluid_1m cuh2_slab_256k; do
  EXTRA=""; [ "$W" = cuh2_batch ] && EXTRA="--configs 16384 --no-extras --cpu-seconds 1"
  CMD="python bench.py --steps 2 --warmup 3 --workload $W $EXTRA"
  $CMD > gpurun_out/${TAG}_plain_$W.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file gpurun_out/${TAG}_launches_$W.csv $CMD > /dev/null 2>&1
done
CMD="python bench.py --steps 2 --warmup 3 --workload cuh2_batch --configs 16384 --no-extras --cpu-seconds 1"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_eam_small -s 3 -c 1 \
    -o gpurun_out/${TAG}_k_eam_small $CMD > /dev/null 2>&1
CMD="python bench.py --steps 2 --warmup 3 --workload lj_fluid_1m"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_brick -s 3 -c 1 \
    -o gpurun_out/${TAG}_k_brick_lj $CMD > /dev/null 2>&1
CMD="python bench.py --steps 2 --warmup 3 --workload cuh2_slab_256k"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_brick -s 6 -c 2 \
    -o gpurun_out/${TAG}_k_brick_eam $CMD > /dev/null 2>&1
ls -la gpurun_out | tail -12
