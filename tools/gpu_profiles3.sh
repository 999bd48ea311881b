#!/bin/bash
# ncu capture of the large-M variance kernel
set -x
python tools/prof_large.py > gpurun_out/plain_l.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_var_large" -s 1 -c 1 -o gpurun_out/prof_large -f python tools/prof_large.py > gpurun_out/ncu_large.log 2>&1
tail -2 gpurun_out/ncu_large.log
