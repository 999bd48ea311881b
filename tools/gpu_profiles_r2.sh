#!/bin/bash
# ncu --set full captures of the round-2 kernels: WHATS="proj sym small train" tools/gpu_profiles_r2.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
for W in ${WHATS:-proj}; do
  case $W in proj) K=k_project;; train) K=k_train_eval;; *) K=k_predict_full;; esac
  WHAT=$W python tools/prof_r2.py > gpurun_out/plain_$W.log 2>&1 && \
  WHAT=$W ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_${TAG}_$W -f python tools/prof_r2.py > gpurun_out/ncu_$W.log 2>&1
  tail -1 gpurun_out/ncu_$W.log
done
