#!/bin/bash
# Runs on the GPU box: compute-sanitizer (memcheck, then racecheck) over a few small parity tests.
mkdir -p gpurun_out
sel='test_generated_object_and_mesh_are_bit_exact and (asteroid_like or mid_noise or zoo or box_types) or test_capsule_absorption_is_bit_exact and sphere or test_streamed_generation and box_types or test_absorption_and_dirty_remesh_are_bit_exact and sphere or test_small_fragment_is_repacked or test_absorb_until_it_splits or test_split_off_sphere'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_extraction.py -m gpu -x -q -k "$sel" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error|Race reported" gpurun_out/sanitize_$tool.log | cut -c1-220 | tail -6
done
