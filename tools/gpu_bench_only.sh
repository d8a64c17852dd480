#!/bin/bash
# Runs on the GPU box (via gpurun): bench only, prints per-kernel times.
tag=${1:-b}
shift
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"])
    print(d["kernel_ms_per_step"])
except Exception as e:
    print("bench failed:", e)
PY
tail -5 gpurun_out/${tag}_bench.err
