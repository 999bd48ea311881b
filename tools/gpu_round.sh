#!/bin/bash
# One GPU-box session: smoke, tests, bench (both arms), ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_ref.json
if [ "${NGPU:-1}" -gt 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NGPU --steps 5 --warmup 3 2> gpurun_out/bench_n$NGPU.err | tee gpurun_out/bench_n$NGPU.json
  tail -5 gpurun_out/bench_n$NGPU.err
fi
if [ -n "$DO_NCU" ]; then
SHORT="python bench.py --steps 2 --warmup 1 --points 4e6 --e2e-points 1e6 --no-cpu-baseline"
$SHORT > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
$SHORT > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_predict_full -s 1 -c 2 -o gpurun_out/prof_full -f $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out
