#!/bin/bash
# ncu captures of the mean + gradient kernel and of the fused kernel's Hessian variant
set -x
python tools/prof_hess.py > gpurun_out/plain_h.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_predict_mean2|k_predict_full" -s 2 -c 2 -o gpurun_out/prof_hess -f python tools/prof_hess.py > gpurun_out/ncu_hess.log 2>&1
tail -2 gpurun_out/ncu_hess.log
