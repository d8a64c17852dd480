#!/usr/bin/env python
"""The rows around the hot path, measured on the 1024^3 asteroid beside the CPU restatement (oracle/, 16 threads where it
threads): meta-graph compile, collision probes, connected regions. One JSON object on stdout.

    python tools/bench_aux.py [--workload asteroid1024] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def best_of(fn, n=5):
    ts = []
    out = None
    for _ in range(n):
        t0 = time.perf_counter()
        out = fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    return min(ts), float(np.median(ts)), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="asteroid1024")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import bench
    from impact_b200 import meta as M
    from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

    ctx = Context(0)
    out = {"workload": args.workload}

    # ---- meta-graph compile (MetaSDFGraph::build_in): the library against its Python mirror, same atomic graph ----
    nodes = M.asteroid_meta_nodes()
    z = np.load(os.path.join(M.DATA_DIR, "asteroid_1024_seed0.npz"))
    scale = float(z["scale_factor"])
    M.compile_meta_nodes(nodes, scale, 0, ctx)  # first call: kernels loaded, pools filled
    t_min, t_med, g = best_of(lambda: M.compile_meta_nodes(nodes, scale, 0, ctx))
    t0 = time.perf_counter()
    g_py = M.MetaCompiler(nodes, scale, 0, ctx).build()
    t_py = 1e3 * (time.perf_counter() - t0)
    out["meta_compile"] = {"ivx_meta_compile_ms": round(t_med, 3), "best_ms": round(t_min, 3), "python_mirror_ms": round(t_py, 1),
                           "atomic_nodes": len(g), "identical": bool(np.array_equal(g.nodes(), g_py.nodes())),
                           "note": "asteroid.vgen.ron (29 meta nodes, three ray-cast crater passes: 440 instances probed on "
                                   "the device per pass), scale factor of the 1024^3 workload"}

    # ---- the object, its mesh, its probes ----
    graph, types, _ = bench.make_workload(args.workload)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    mesh = VoxelObjectMesh.create(obj)
    mesh.collision_probes()
    ctx.synchronize()
    t_min, t_med, pr = best_of(lambda: mesh.collision_probes())
    info = ctx._lib  # noqa: F841
    out["collision_probes"] = {"gpu_ms_incl_download": round(t_med, 3), "best_ms": round(t_min, 3), "points": int(len(pr["points"])),
                               "chunks": int(len(pr["ranges"])), "submeshes": int(mesh.n_submeshes),
                               "vertices": int(mesh.n_vertices), "log2_block_size": pr["log2_block_size"]}
    r0 = obj.resolve_connected_regions()
    t_min, t_med, r = best_of(lambda: obj.resolve_connected_regions())
    out["connected_regions_unchanged_object"] = {"gpu_ms": round(t_med, 3), "best_ms": round(t_min, 3),
                                                 "relabelled_chunks_first_resolve": r0["n_relabelled_chunks"],
                                                 "local_regions": r["n_local_regions"], "connections": r["n_connections"],
                                                 "regions": r["n_regions"]}

    # ---- contacts between two copies of the object pushed into each other ----
    from impact_b200 import voxel as V
    other = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    mesh_other = VoxelObjectMesh.create(other)
    mesh_other.collision_probes()
    dens = np.float32([1.0, 2.7, 0.3, 5.5])
    mom_a, mom_b = obj.inertial_moments(dens).copy(), other.inertial_moments(dens).copy()
    shape = np.float64(obj.info()["grid_shape"])
    q = np.float64([0.0, np.sin(0.35), 0.0, np.cos(0.35)])  # B turned 40 degrees about y and shifted by 0.8 of the width
    centre = 0.5 * shape

    def rot(qq, v):
        b, w = qq[:3], qq[3]
        return v * (w * w - b @ b) + b * (2.0 * (v @ b)) + np.cross(b, v) * (2.0 * w)

    world_to_a = np.float32([0, 0, 0, 1, *centre])
    world_to_b = np.float32([*q, *(centre - rot(q, np.float64([0.8 * shape[0], 0.05 * shape[1], 0.0])))])
    qa_inv = np.float64([0, 0, 0, 1])
    qb_inv = np.float64([-q[0], -q[1], -q[2], q[3]])
    # transform_from_b_to_a = world_to_a * world_to_b.inverted()
    t_inv = -rot(qb_inv, np.float64(world_to_b[4:]))
    b_to_a = np.float32([*qb_inv, *(t_inv + centre)])
    ranges = V.intersection_voxel_ranges(obj.info()["occupied_voxel_ranges"], 1.0, other.info()["occupied_voxel_ranges"], 1.0,
                                         b_to_a[:4], b_to_a[4:])
    if ranges is not None:
        V.mutual_contacts(obj, other, world_to_a, world_to_b, ranges[0], ranges[1], mom_a, mom_b)
        t_min, t_med, (c_ab, c_ba) = best_of(lambda: V.mutual_contacts(obj, other, world_to_a, world_to_b, ranges[0], ranges[1], mom_a, mom_b))
        out["mutual_contacts"] = {"gpu_ms_incl_download": round(t_med, 3), "best_ms": round(t_min, 3), "contacts_a_in_b": int(len(c_ab)),
                                  "contacts_b_in_a": int(len(c_ba)),
                                  "note": "two copies of the object, the second turned 40 degrees about y and shifted by 0.8 of the width"}
    else:
        out["mutual_contacts"] = {"note": "the chosen pose does not intersect"}

    # ---- the renderer's mesh buffers (VoxelMeshGPUBuffers): creation, and a sync after absorb + mesh sync ----
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_fracture import absorber_path
    robj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    VoxelObjectMesh.create(robj)
    ctx.synchronize()
    t0 = time.perf_counter()
    bufs = V.VoxelMeshGPUBuffers.for_voxel_object(robj)
    ctx.synchronize()
    t_create = 1e3 * (time.perf_counter() - t0)
    created_bytes = bufs.bytes_copied
    centers, radius = absorber_path(robj.info()["grid_shape"], 10)
    ts, moved, n_ranges = [], [], []
    for c in centers:
        st = robj.absorb_sphere(c, radius, radius + 2.0)
        VoxelObjectMesh.sync(robj)
        ctx.synchronize()
        t0 = time.perf_counter()
        bufs.sync_with_voxel_object()
        ctx.synchronize()
        if st["touched_chunks"] > 100:
            ts.append(1e3 * (time.perf_counter() - t0))
            moved.append(bufs.bytes_copied)
            n_ranges.append(bufs.n_updated_ranges)
    out["render_buffers"] = {"create_ms": round(t_create, 3), "create_bytes": int(created_bytes),
                             "sync_ms_median": round(float(np.median(ts)), 3), "sync_bytes_median": int(np.median(moved)),
                             "updated_ranges_median": int(np.median(n_ranges)), "recreated_in_last_sync": bufs.recreated,
                             "note": "five exportable device allocations (one file descriptor each); a sync copies the updated "
                                     "ranges and the submesh table device to device, after an absorption step of config 5"}
    bufs.close()
    robj.free()

    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_lib as O
        O.build()
        threads = min(16, os.cpu_count() or 1)
        ogen = O.Generator(graph.nodes(), graph.root_node_id)
        t0 = time.perf_counter()
        ocpu = O.Object.generate(O.VoxelGenerator(ogen, 1.0, types), threads)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        omesh = ocpu.mesh(threads)
        t_mesh = time.perf_counter() - t0
        t0 = time.perf_counter()
        opr = O.CollisionProbes(ocpu, omesh)
        t_pr = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        sd = ocpu.split_detection()
        t_sd = 1e3 * (time.perf_counter() - t0)
        same = bool(len(opr.points) == len(pr["points"]) and
                    np.array_equal(opr.points.view(np.uint32), pr["points"].view(np.uint32)))
        if ranges is not None:
            t0 = time.perf_counter()
            o_ab, o_ba = O.mutual_contacts(ocpu, opr, mom_a, world_to_a, ocpu, opr, mom_b, world_to_b, ranges[0], ranges[1])
            out["mutual_contacts"]["cpu_ms_1_thread"] = round(1e3 * (time.perf_counter() - t0), 2)
            out["mutual_contacts"]["identical_to_cpu"] = bool(
                len(o_ab) == len(c_ab) and len(o_ba) == len(c_ba) and
                np.array_equal(o_ab["position"].view(np.uint32), c_ab["position"].view(np.uint32)) and
                np.array_equal(o_ba["normal"].view(np.uint32), c_ba["normal"].view(np.uint32)) and
                np.array_equal(o_ab["depth"].view(np.uint32), c_ab["depth"].view(np.uint32)))
        out["cpu_port"] = {"threads": threads, "generate_s": round(t_gen, 2), "mesh_s": round(t_mesh, 2),
                           "collision_probes_ms_1_thread": round(t_pr, 2), "probes_identical_to_gpu": same,
                           "split_detection_first_ms_1_thread": round(t_sd, 1), "regions": sd["n_regions"],
                           "regions_identical_to_gpu": bool(sd["n_regions"] == r["n_regions"])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
