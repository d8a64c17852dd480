#!/bin/bash
# Builds an experimental variant of the CUDA library next to the real one:
#   tools/build_variant.sh NAME "-DIVX_TYPES_CTAS=4 ..."  →  impact_b200/csrc/_build/var_NAME/libimpact_voxel_cuda.so
# Select it at run time with IMPACT_VOXEL_CUDA_LIB=<path> (impact_b200/_lib.py). For A/B timing on the GPU box only.
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../impact_b200/csrc"
out=_build/var_$name
mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -DIVX_FMAD_OFF --expt-relaxed-constexpr -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math $flags"
pids=()
for f in generate types util derive mesh modify extract halo split inertia query comm mesh_sync probes regions render_buffers api; do $NV -c $f.cu -o $out/$f.o & pids+=($!); done
for f in program geometry meta; do $NV -x cu -c $f.cpp -o $out/$f.o & pids+=($!); done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libimpact_voxel_cuda.so $out/*.o
echo built $out/libimpact_voxel_cuda.so
