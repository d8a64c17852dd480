#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2f}
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_baseline_sizes.py tests/test_meta.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -6 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "parity", d["parity"]["digest_ok"])
print({k: round(v,3) for k,v in d["kernel_ms_per_step"].items()})
PY
