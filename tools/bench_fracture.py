#!/usr/bin/env python
"""BASELINE config 5: repeated spherical absorption on the 1024^3 asteroid with dirty-chunk remesh.

    python tools/bench_fracture.py [--workload asteroid1024] [--steps 32] [--cpu-steps 4]

Geometry of the reference's `update_mesh` bench (engine/src/benchmark/benchmarks/voxel_object.rs:343-362):
an absorbing sphere of radius 0.15 R starts on the bounding sphere along the (1,1,1) diagonal and moves
inward by one absorber radius per step; after each step the invalidated chunks are re-meshed
(`sync_with_voxel_object`) and, every step, connected regions are re-resolved when the library offers it.
Prints one JSON line (per-step device times from CUDA events inside the library, host wall time per step,
dirty chunks / touched voxels per step) and, with --cpu-steps > 0, the CPU restatement (oracle/) timed on
the same first steps after generating the same object on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def absorber_path(shape, steps):
    R = 0.5 * float(max(shape))
    radius = np.float32(0.15 * R)
    start = (0.5 * np.asarray(shape, np.float64) - R / np.sqrt(3.0)).astype(np.float32)
    d = np.float32(1.0 / np.sqrt(3.0))
    return [(start + np.float32(s) * radius * d).astype(np.float32) for s in range(steps)], float(radius)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="asteroid1024")
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--cpu-steps", type=int, default=0)
    ap.add_argument("--split", action="store_true", help="also resolve connected regions and split off disconnected ones after every step")
    ap.add_argument("--synced-mesh", action="store_true",
                    help="keep the object's mesh on the device and patch it in place (ivx_object_mesh_sync = "
                         "VoxelObjectMesh::sync_with_voxel_object with RangeAllocator placement) instead of returning the "
                         "re-meshed chunks as a compact patch (ivx_object_remesh_dirty)")
    ap.add_argument("--inertial", action="store_true",
                    help="attach the inertial-property updater to every absorption (ivx_object_absorb_sphere_inertial)")
    args = ap.parse_args()

    import bench
    from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

    graph, types, desc = bench.make_workload(args.workload)
    ctx = Context(0)
    gen = ctx.build_generator(graph)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, types))
    info = obj.info()
    shape = info["grid_shape"]
    VoxelObjectMesh.create(obj)
    centers, radius = absorber_path(shape, args.steps)
    influence = radius + 2.0  # absorption.rs:170-179: influence radius = radius + 2 voxel extents

    densities = np.float32([1.0, 2.7, 0.3, 5.5])
    moments = obj.inertial_moments(densities).copy() if args.inertial else None
    moments_start = None if moments is None else moments.copy()

    def step(obj, c, fragments, moments):
        t0 = time.perf_counter()
        if args.inertial:
            st = obj.absorb_sphere_inertial(c, radius, influence, densities, moments)
        else:
            st = obj.absorb_sphere(c, radius, influence)
        n_dirty = st["dirty_chunks"]  # size of invalidated_mesh_chunk_indices, part of the call's statistics
        t1 = time.perf_counter()
        patch = VoxelObjectMesh.sync(obj) if args.synced_mesh else VoxelObjectMesh.sync_with_voxel_object(obj)
        ctx.synchronize()
        t2 = time.perf_counter()
        regions = None
        n_extracted = n_discarded = 0
        if args.split:
            # handle_voxel_object_after_removing_voxels: resolve, then split off disconnected regions while there
            # are any (every extraction re-resolves); the fragments' meshes are created like any new object's
            while True:
                xi, frag = obj.extract_any_disconnected_region()
                if regions is None:
                    regions = xi["n_regions_before"]
                if not xi["found_two"]:
                    break
                if xi["extracted"]:
                    n_extracted += 1
                    VoxelObjectMesh.create(frag)
                    fragments.append(frag)
                else:
                    n_discarded += 1
            if n_extracted or n_discarded:
                VoxelObjectMesh.sync(obj) if args.synced_mesh else VoxelObjectMesh.sync_with_voxel_object(obj)
            ctx.synchronize()
        t3 = time.perf_counter()
        return {"absorb_ms": 1e3 * (t1 - t0), "remesh_ms": 1e3 * (t2 - t1), "split_ms": 1e3 * (t3 - t2),
                "extracted": n_extracted, "discarded": n_discarded,
                "touched_chunks": st["touched_chunks"], "touched_voxels": st["touched_voxels"],
                "emptied_voxels": st["emptied_voxels"], "dirty_chunks": n_dirty,
                "remeshed_submeshes": patch.n_submeshes, "patch_vertices": patch.n_vertices,
                "regions": regions}

    # warm-up on a scratch copy of the object: the first three steps of the sequence, untimed (kernels loaded, pools grown)
    scratch = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, types))
    VoxelObjectMesh.create(scratch)
    scratch_moments = None if moments is None else moments.copy()
    scratch_fragments = []
    for c in centers[:3]:
        step(scratch, c, scratch_fragments, scratch_moments)
    for f in scratch_fragments:
        f.free()
    scratch.free()
    ctx.synchronize()

    ctx.profile_enable(True)
    ctx.profile_reset()
    per_step = []
    fragments = []  # extracted objects stay alive, like the entities the engine spawns for them
    launches0 = ctx.kernel_launch_count
    t_all = time.perf_counter()
    for c in centers:
        per_step.append(step(obj, c, fragments, moments))
    wall = time.perf_counter() - t_all
    prof = ctx.profile_get()
    launches = ctx.kernel_launch_count - launches0
    mean = lambda k: float(np.mean([s[k] for s in per_step]))
    touched = sum(s["touched_voxels"] for s in per_step)
    out = {
        "metric": "fracture_step_ms", "workload": args.workload, "description": desc, "grid_shape": list(shape),
        "steps": args.steps, "absorber_radius_voxels": radius,
        "remesh": "ivx_object_mesh_sync (mesh patched in place on the device)" if args.synced_mesh else "ivx_object_remesh_dirty (compact patch)",
        "ms_per_step": 1e3 * wall / args.steps, "absorb_ms": mean("absorb_ms"), "remesh_ms": mean("remesh_ms"),
        # steps that touch the object (the absorber leaves it after ~13 steps), medians: without the one-off device
        # allocations of the first steps (pool growth, first buffers of a new size)
        "active_steps": len([x for x in per_step if x["touched_chunks"]]),
        "median_active_absorb_ms": float(np.median([x["absorb_ms"] for x in per_step if x["touched_chunks"]] or [0.0])),
        "median_active_remesh_ms": float(np.median([x["remesh_ms"] for x in per_step if x["touched_chunks"]] or [0.0])),
        "median_active_split_ms": float(np.median([x["split_ms"] for x in per_step if x["touched_chunks"]] or [0.0])) if args.split else None,
        "split_ms": mean("split_ms") if args.split else None,
        "fragments_extracted": sum(s["extracted"] for s in per_step), "fragments_dropped": sum(s["discarded"] for s in per_step),
        "split_note": "resolve connected regions + extract_any_disconnected_region until one region is left + mesh the fragments",
        "dirty_chunks_per_step": mean("dirty_chunks"), "touched_chunks_per_step": mean("touched_chunks"),
        "touched_voxels_per_step": mean("touched_voxels"), "emptied_voxels_total": sum(s["emptied_voxels"] for s in per_step),
        "touched_voxels_per_s": touched / wall,
        # modification moves 6 B per voxel of the touched chunks (3 read + 3 written), SURVEY 8d
        "absorb_kernel_ms_per_step": prof["absorb"][0] / max(1, args.steps),
        "absorb_kernel_gbs": (6.0 * 4096 * sum(s["touched_chunks"] for s in per_step)) / max(prof["absorb"][0] * 1e-3, 1e-12) / 1e9,
        "mesh_kernel_ms_per_step": (prof["mesh_count"][0] + prof["mesh_emit"][0]) / max(1, args.steps),
        "gpu_launches": int(launches), "timing": "host wall clock around synchronous C-ABI calls, after three untimed warm-up steps on a scratch copy of the object; kernel times from CUDA events",
        "last_step": per_step[-1],
        "per_step_ms": [[round(x["absorb_ms"], 3), round(x["remesh_ms"], 3), round(x["split_ms"], 3), x["dirty_chunks"]] for x in per_step],
    }
    if args.inertial:
        scratch = obj.inertial_moments(densities)
        out["inertial"] = {
            "note": "VoxelObjectInertialPropertyUpdater attached to every absorption (bit-exact ordered subtraction); "
                    "absorb_ms includes it",
            "ordered_sum_kernel_ms_per_step": prof["moments_sum"][0] / max(1, args.steps),
            "mass_before": float(moments_start[0]), "mass_after_incremental": float(moments[0]),
            "mass_after_from_scratch": float(scratch[0]),
            "max_rel_dev_incremental_vs_scratch": float(np.max(np.abs(moments - scratch) / np.maximum(np.abs(scratch), 1e-30))),
        }

    if args.cpu_steps > 0:
        from oracle import oracle_lib as O

        threads = os.cpu_count() or 1
        ogen = O.Generator(graph.nodes(), graph.root_node_id)
        t0 = time.perf_counter()
        oobj = O.Object.generate(O.VoxelGenerator(ogen, 1.0, types), threads)
        oobj.mesh(threads)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        n_dirty = 0
        for c in centers[: args.cpu_steps]:
            oobj.absorb_sphere(c, radius, influence)
            dirty = oobj.dirty()
            n_dirty += len(dirty)
            cc = oobj.info()["chunk_counts"]
            for lin in dirty:
                oobj.mesh_chunk(int(lin // (cc[1] * cc[2])), int((lin // cc[2]) % cc[1]), int(lin % cc[2]))
            oobj.clear_dirty()
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"kind": "port", "cores": 1, "steps": args.cpu_steps, "ms_per_step": 1e3 * dt / args.cpu_steps,
                               "generate_and_mesh_s": t_gen, "generate_threads": threads,
                               "note": "oracle absorb + per-chunk remesh of the dirty set, single thread (the "
                                       "reference's absorb / sync_with_voxel_object are single-threaded too)"}
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
