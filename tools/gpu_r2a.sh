#!/bin/bash
# round 2, first call: full-size parity tests (new) + N=1 bench line with the parity block
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_baseline_sizes.py -x -q 2>&1 | tail -25 > gpurun_out/r2a_sizes.log
tail -25 gpurun_out/r2a_sizes.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2a_bench.json"))
print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "parity", d["parity"])
print(d["kernel_ms_per_step"]); print(d["cpu_baseline"])
PY
