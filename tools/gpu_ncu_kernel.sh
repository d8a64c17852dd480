#!/bin/bash
# one `ncu --set full` capture of a named kernel while a script runs: tools/gpu_ncu_kernel.sh <tag> <kernel regex> <skip> <count> <command...>
tag=$1; kern=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kern" --launch-skip "$skip" -c "$count" \
    -o gpurun_out/${tag} -f "$@" > gpurun_out/${tag}.log 2>&1
tail -5 gpurun_out/${tag}.log
ls -la gpurun_out/${tag}.ncu-rep
