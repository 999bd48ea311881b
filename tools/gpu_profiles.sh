#!/bin/bash
# ncu evidence for the round: launch list of the bench + full captures of the FP64 fused kernel and the TF32 kernel
set -x
SHORT="python bench.py --steps 2 --warmup 1 --points 4e6 --e2e-points 1e6 --no-cpu-baseline"
$SHORT > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
$SHORT > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_predict_full -s 1 -c 1 -o gpurun_out/prof_full_final -f $SHORT > gpurun_out/ncu_full.log 2>&1
python tools/prof_tf32.py > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_predict_tf32 -s 1 -c 1 -o gpurun_out/prof_tf32 -f python tools/prof_tf32.py > gpurun_out/ncu_tf32.log 2>&1
tail -n 2 gpurun_out/ncu_full.log; tail -n 2 gpurun_out/ncu_tf32.log
