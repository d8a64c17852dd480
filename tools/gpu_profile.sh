#!/bin/bash
# Runs on the GPU box (via gpurun): ncu launch list of the bench command and full captures of the two
# dominant kernels. Outputs under gpurun_out/<tag>_*; numbers printed under ncu are never bench values.
tag=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
for k in k_types k_eval; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k} -c 1 -f -o gpurun_out/${tag}_${k} \
      python tools/dbg.py asteroid1024 > gpurun_out/${tag}_ncu_${k}.log 2>&1
done
ls -la gpurun_out | tail -8
