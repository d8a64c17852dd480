#!/bin/bash
# Runs on the GPU box: bench.py once per library variant (tools/build_variant.sh), prints per-kernel times.
# usage: tools/gpu_ab.sh TAG VARIANT...   ("base" = the in-tree library)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset IMPACT_VOXEL_CUDA_LIB; else export IMPACT_VOXEL_CUDA_LIB=$PWD/impact_b200/csrc/_build/var_$v/libimpact_voxel_cuda.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${v}_bench.json 2> gpurun_out/${tag}_${v}_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${v}_bench.json"))
    print("$v", "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(x,3) for k,x in d["kernel_ms_per_step"].items()})
except Exception as e:
    print("$v bench failed:", e)
PY
done
