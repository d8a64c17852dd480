#!/bin/bash
# Runs on the GPU box: tools/dbg_streamed.py once per library variant.
for v in "$@"; do
  if [ "$v" = base ]; then unset IMPACT_VOXEL_CUDA_LIB; else export IMPACT_VOXEL_CUDA_LIB=$PWD/impact_b200/csrc/_build/var_$v/libimpact_voxel_cuda.so; fi
  echo "== $v"; timeout 300 python tools/dbg_streamed.py 2>&1 | tail -2 | cut -c1-100
done
