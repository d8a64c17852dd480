#!/bin/bash
# Runs on the GPU box: compute-sanitizer (memcheck, then racecheck) over small parity tests that together launch every
# kernel of the library (generation, types, derive, mesh, absorb incl. range modes and inertial updates, ingest of
# generated chunks, moments, surface queries and contacts, split detection incl. the device-side global pass, extraction, synced mesh, collision
# probes, the block probes of the meta compiler and the
# slab protocol; the peer-memory communicator is left out: its on-device waits need two streams running concurrently, which
# the sanitizer serialises). Summaries go to gpurun_out/sanitize_<tool>.log (copy into profiles/).
mkdir -p gpurun_out
tag=${1:-r2}
sel='(test_generated_object_and_mesh_are_bit_exact and (asteroid_like or mid_noise or zoo or box_types)) or (test_capsule_absorption_is_bit_exact and sphere) or (test_streamed_generation and box_types) or (test_absorption_and_dirty_remesh_are_bit_exact and sphere) or test_small_fragment_is_repacked or test_absorb_until_it_splits or test_split_off_sphere or (test_voxel_type_generators_are_bit_exact) or test_absorption_updates_the_moments_bit_for_bit or test_moments_from_scratch or (test_mutual and identity) or (test_gpu_surface_voxels and sphere) or (test_gpu_sphere_contacts and sphere) or (test_gpu_plane_and_capsule_contacts and sphere) or (test_fixtures and box) or (test_random_voxel_grids) or (test_synced_mesh_follows and asteroid_like) or (test_slabs_in_one_process and zoo and 2) or test_debris or (test_probes_of_all_chunks_match_the_oracle and (tiny2 or asteroid_like)) or test_probes_follow_the_synced_mesh or (test_ragged_random_grids and 11) or (test_connected_regions_match_the_oracle and (two_spheres or noisy_debris)) or (test_probing_kinds and RayTranslation and ShapeBoundary) or (test_mutual_contacts_match_the_oracle and (rotated_extents or deep)) or test_chunks_too_large_for_the_shared_memory_lists'
files="tests/test_gpu_parity.py tests/test_gpu_extraction.py tests/test_gpu_inertia.py tests/test_gpu_mutual_absorption.py tests/test_surface_voxels.py tests/test_gpu_generated_chunks.py tests/test_gpu_synced_mesh.py tests/test_gpu_slabs.py tests/test_gpu_split_detection.py tests/test_collision_probes.py tests/test_meta_native.py tests/test_mutual_contacts.py"
for tool in memcheck racecheck; do
  timeout 2400 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 30 \
      python -m pytest $files -m gpu -q -k "$sel" > gpurun_out/sanitize_${tag}_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|Race reported|Invalid|Hazard" gpurun_out/sanitize_${tag}_$tool.log | cut -c1-200 | sort | uniq -c | sort -rn | head -12
done
