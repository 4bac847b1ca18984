M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"
python tools/bench_batch_dev.py --configs 16384 --steps 1 > gpurun_out/plain_q.log 2>&1 &&
ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_eam_small -s 3 -c 1 -o gpurun_out/r02d_k_eam_small python tools/bench_batch_dev.py --configs 16384 --steps 1 > gpurun_out/ncu_q.log 2>&1
tail -1 gpurun_out/ncu_q.log
