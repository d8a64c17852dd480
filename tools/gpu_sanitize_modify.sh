#!/bin/bash
# Runs on the GPU box: compute-sanitizer (memcheck, racecheck) over the tests that drive the modification path
# (absorb_impl: k_absorb_apply, k_boundary_prep_box, k_boundary_apply<box>, k_occupied_ranges) and the synced mesh.
mkdir -p gpurun_out
tag=${1:-r2m}
sel='(test_capsule_absorption_is_bit_exact and sphere) or (test_absorption_and_dirty_remesh_are_bit_exact and sphere) or test_absorb_until_it_splits or test_absorption_updates_the_moments_bit_for_bit or (test_mutual and identity) or (test_synced_mesh_follows and asteroid_like) or test_small_fragment_is_repacked or (test_ragged_random_grids and 11) or (test_connected_regions_match_the_oracle and two_spheres)'
files="tests/test_gpu_parity.py tests/test_gpu_extraction.py tests/test_gpu_inertia.py tests/test_gpu_mutual_absorption.py tests/test_gpu_synced_mesh.py tests/test_gpu_split_detection.py"
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 30 \
      python -m pytest $files -m gpu -q -k "$sel" > gpurun_out/sanitize_${tag}_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|Race reported|Invalid|Hazard" gpurun_out/sanitize_${tag}_$tool.log | cut -c1-220 | sort | uniq -c | sort -rn | head -12
done
