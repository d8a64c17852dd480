#!/bin/bash
# Runs on the GPU box: ncu launch lists of one generate+mesh step at 1024^3 and 2048^3 (tools/probe_scale.py).
mkdir -p gpurun_out
for hi in 1024 2048; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/scale_${hi}_launches.csv \
      python tools/probe_scale.py $hi 1 time-only > gpurun_out/scale_${hi}_launches.log 2>&1
done
ls -la gpurun_out | tail -5
