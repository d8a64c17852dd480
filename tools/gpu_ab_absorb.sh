#!/bin/bash
# Runs on the GPU box: tools/probe_absorb.py once per library variant (tools/build_variant.sh); "base" = the in-tree library
for v in "$@"; do
  if [ "$v" = base ]; then unset IMPACT_VOXEL_CUDA_LIB; else export IMPACT_VOXEL_CUDA_LIB=$PWD/impact_b200/csrc/_build/var_$v/libimpact_voxel_cuda.so; fi
  echo "== $v"; timeout 120 python tools/probe_absorb.py 2>&1 | sed -n '4,12p' | awk '{print $4, $5, $14, $15}' | tr '\n' ';'; echo
done
