#!/usr/bin/env python
"""The reference's own criterion benchmarks for this path, same shapes and parameters, through the C ABI on one GPU and
(with --cpu) through the CPU oracle beside it:

    engine/src/benchmark/benchmarks/generation.rs:24-178   generate_box ... generate_object_from_complex_graph
    engine/src/benchmark/benchmarks/voxel_object.rs:101-450 compute_all_derived_state ... obtain_mutual_voxel_object_contacts

Each entry: median wall time of the call a user of the reference would swap in (synchronised on both sides, after two
untimed repetitions), what the call covers when that differs from the reference's benchmark body (the library derives
the internal state inside generation; results come back to host memory where the reference hands them to a closure),
and the CPU restatement on ONE core (the reference's benchmarks are single-threaded except the asteroid, which uses 8
workers — so does the oracle there). No published numbers exist for these benchmarks. One JSON document.

    python tools/bench_reference_suite.py [--cpu] [--reps 7]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()

    from impact_b200 import meta as M
    from impact_b200 import voxel as V
    from impact_b200.graph import SDFGraph, VoxelTypeGenerator
    from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject, VoxelObjectMesh, compile_program_host

    O = None
    if args.cpu:
        from oracle import oracle_lib as O  # the CPU baseline of this bench (test infrastructure, not on the product path)

        O.build()

    ctx = Context(0)
    SAME = VoxelTypeGenerator.same(0)
    GRADIENT = VoxelTypeGenerator.gradient_noise([0, 1, 2, 3], 0.02, 1.0, 0)
    out = {"suite": "engine/src/benchmark/benchmarks/{generation,voxel_object}.rs", "reps": args.reps,
           "cpu": "oracle (C++ restatement of the reference), 1 core unless stated" if args.cpu else None, "benchmarks": {}}

    def timed(fn, reps=None, setup=None):
        ts, res = [], None
        for r in range((reps or args.reps) + 2):
            state = setup() if setup else None
            ctx.synchronize()
            t0 = time.perf_counter()
            res = fn(state) if setup else fn()
            ctx.synchronize()
            if r >= 2:
                ts.append(1e3 * (time.perf_counter() - t0))
        return float(np.median(ts)), res

    def cpu_timed(fn, reps=3, setup=None):
        ts, res = [], None
        for _ in range(reps):
            state = setup() if setup else None
            t0 = time.perf_counter()
            res = fn(state) if setup else fn()
            ts.append(1e3 * (time.perf_counter() - t0))
        return float(np.median(ts)), res

    def record(name, ref, gpu_ms, covers, cpu_ms=None, **extra):
        e = {"reference": ref, "gpu_ms": round(gpu_ms, 4), "covers": covers}
        if cpu_ms is not None:
            e["cpu_ms"] = round(cpu_ms, 3)
            e["cpu_over_gpu"] = round(cpu_ms / gpu_ms, 1)
        e.update(extra)
        out["benchmarks"][name] = e
        print(name, e, file=sys.stderr, flush=True)

    def cpu_generator(graph, types):
        return O.VoxelGenerator(O.Generator(graph.nodes(), graph.root_node_id), 1.0, types)

    # ---------------- generation.rs ----------------
    def g_box():
        g = SDFGraph()
        g.box([80.0] * 3)
        return g

    def g_sphere_union():
        g = SDFGraph()
        s1 = g.sphere(60.0)
        s2 = g.translation(g.sphere(60.0), [50.0, 0.0, 0.0])
        g.union(s1, s2, 1.0)
        return g

    def g_complex():
        g = SDFGraph()
        s = g.translation(g.sphere(60.0), [50.0, 0.0, 0.0])
        b = g.rotation_from_axis_angle(g.scaling(g.box([50.0, 60.0, 70.0]), 0.9), [0.0, 1.0, 0.0], 10.0)
        g.union(s, b, 1.0)
        return g

    def g_noise():
        g = SDFGraph()
        g.multifractal_noise(g.sphere(80.0), 8, 0.02, 2.0, 0.6, 4.0, 0)
        return g

    gen_cases = [("generate_box", "generation.rs:24-38", g_box(), SAME),
                 ("generate_sphere_union", "generation.rs:40-59", g_sphere_union(), SAME),
                 ("generate_complex_object", "generation.rs:61-85", g_complex(), SAME),
                 ("generate_object_with_multifractal_noise", "generation.rs:87-103", g_noise(), SAME),
                 ("generate_box_with_gradient_noise_voxel_types", "generation.rs:105-129", g_box(), GRADIENT)]
    for name, ref, graph, types in gen_cases:
        vg = SDFVoxelGenerator(1.0, ctx.build_generator(graph), types)

        def run():
            o = VoxelObject.generate(vg)
            shape = o.info()["grid_shape"]
            o.free()
            return shape

        ms, shape = timed(run)
        cpu = None
        if O:
            cvg = cpu_generator(graph, types)
            cpu, _ = cpu_timed(lambda: O.Object.generate(cvg, 1))
        record(name, ref, ms, "ivx_object_generate: generate_without_derived_state AND compute_all_derived_state "
               "(the CPU time likewise)", cpu, grid_shape=[int(x) for x in shape])

    nodes = M.asteroid_meta_nodes()
    M.compile_meta_nodes(nodes, 1.0, 0, ctx)
    ms, atomic = timed(lambda: M.compile_meta_nodes(nodes, 1.0, 0, ctx))
    cpu = None
    if O:
        cpu, mirror = cpu_timed(lambda: M.MetaCompiler(nodes, 1.0, 0, ctx).build(), reps=1)
    record("compile_complex_meta_graph", "generation.rs:131-139", ms,
           "ivx_meta_compile of asteroid.vgen.ron (scale 1, seed 0); the comparison is the readable Python mirror of "
           "the same compile, not a CPU port", cpu, atomic_nodes=len(atomic))
    ms, _ = timed(lambda: compile_program_host(atomic))
    record("build_complex_atomic_graph", "generation.rs:141-151", ms, "ivx_program_compile_host (SDFGraph::build_in, host only)")
    vg = SDFVoxelGenerator(1.0, ctx.build_generator(atomic), SAME)

    def run_asteroid():
        o = VoxelObject.generate(vg)
        shape = o.info()["grid_shape"]
        o.free()
        return shape

    ms, shape = timed(run_asteroid)
    cpu = None
    if O:
        cvg = cpu_generator(atomic, SAME)
        cpu, _ = cpu_timed(lambda: O.Object.generate(cvg, 8))
    record("generate_object_from_complex_graph", "generation.rs:153-178", ms,
           "ivx_object_generate incl. derived state; CPU: 8 worker threads like the reference's thread pool", cpu,
           grid_shape=[int(x) for x in shape])

    # ---------------- voxel_object.rs: a sphere of radius 100 (202^3) ----------------
    def sphere_graph(r):
        g = SDFGraph()
        g.sphere(r)
        return g

    R = 100.0
    sg = sphere_graph(R)
    svg = SDFVoxelGenerator(1.0, ctx.build_generator(sg), SAME)
    obj = VoxelObject.generate(svg)
    info = obj.info()
    occ = np.asarray(info["occupied_voxel_ranges"], np.float64)  # (3, 2): per axis [begin, end)
    lo, hi = occ[:, 0], occ[:, 1]
    center = 0.5 * (lo + hi)
    diag = np.ones(3) / np.sqrt(3.0)
    absorber_center = (center - R * diag).astype(np.float32)
    absorber_radius = 0.15 * R
    dens = np.ones(256, np.float32)
    cobj = None
    if O:
        csvg = cpu_generator(sg, SAME)
        cobj = O.Object.generate(csvg, 1)

    ms, _ = timed(lambda: VoxelObject.generate(svg).free())
    cpu = cpu_timed(lambda: O.Object.generate(csvg, 1))[0] if O else None
    record("compute_all_derived_state", "voxel_object.rs:101-109", ms,
           "ivx_object_generate of the sphere: generation AND derived state (the library has no separate derive call); "
           "CPU likewise", cpu)

    ms, mom = timed(lambda: obj.inertial_moments(dens))
    cpu = cpu_timed(lambda: cobj.inertial_moments(dens))[0] if O else None
    record("initialize_inertial_properties", "voxel_object.rs:111-118", ms, "ivx_object_inertial_moments", cpu)

    ms, mesh = timed(lambda: VoxelObjectMesh.create(obj))
    cpu = cpu_timed(lambda: cobj.mesh(1))[0] if O else None
    record("create_mesh", "voxel_object.rs:126-130", ms, "ivx_object_mesh (the mesh stays on the device)", cpu,
           vertices=int(mesh.n_vertices), indices=int(mesh.n_indices))

    mesh.collision_probes()
    ms, pr = timed(lambda: mesh.collision_probes())
    cpu = None
    if O:
        cmesh = cobj.mesh(1)
        cpu, _ = cpu_timed(lambda: O.CollisionProbes(cobj, cmesh))
    record("compute_collision_probes", "voxel_object.rs:132-138", ms,
           "ivx_object_collision_probes + ivx_collision_probes_download", cpu, points=int(len(pr["points"])))

    n = (np.ones(3) / np.sqrt(3.0)).astype(np.float32)
    ms, sv = timed(lambda: obj.surface_voxels_within_plane(n, 0.4 * R))
    record("obtain_surface_voxels_within_negative_halfspace_of_plane", "voxel_object.rs:155-172", ms,
           "ivx_object_surface_voxels_within_plane, records downloaded", None, voxels=int(len(sv)))
    ms, sv = timed(lambda: obj.surface_voxels_touching_sphere(absorber_center, absorber_radius))
    record("obtain_surface_voxels_within_sphere", "voxel_object.rs:174-192", ms,
           "ivx_object_surface_voxels_touching_sphere, records downloaded", None, voxels=int(len(sv)))

    # modify_voxels_within_sphere: the reference's closure only looks at the voxels (black_box); the library's modification
    # is the absorption closure (the one the engine runs), same sphere, same object modified over and over
    mobj = VoxelObject.generate(svg)
    ms, st = timed(lambda: mobj.absorb_sphere(absorber_center, absorber_radius, absorber_radius))
    cpu = None
    if O:
        cm = O.Object.generate(csvg, 1)
        cpu, _ = cpu_timed(lambda: cm.absorb_sphere(absorber_center, absorber_radius, absorber_radius))
    record("modify_voxels_within_sphere", "voxel_object.rs:209-224", ms,
           "ivx_object_absorb_sphere (visit + signed-distance update + internal state + boundary refresh)", cpu,
           touched_voxels=int(st["touched_voxels"]))

    # update_mesh: modify + mesh.sync_with_voxel_object, repeated on the same object and mesh
    uobj = VoxelObject.generate(svg)
    VoxelObjectMesh.create(uobj)

    def run_update():
        uobj.absorb_sphere(absorber_center, absorber_radius, absorber_radius)
        return VoxelObjectMesh.sync(uobj)

    ms, _ = timed(run_update)
    cpu = None
    if O:
        cu = O.Object.generate(csvg, 1)
        csm = O.SyncedMesh(cu, 1)
        cu.clear_dirty()

        def cpu_update():
            cu.absorb_sphere(absorber_center, absorber_radius, absorber_radius)
            d = np.sort(cu.dirty())
            csm.sync(cu, d)
            cu.clear_dirty()

        cpu, _ = cpu_timed(cpu_update)
    record("update_mesh", "voxel_object.rs:343-362", ms, "ivx_object_absorb_sphere + ivx_object_mesh_sync", cpu)

    # apply_mutual_voxel_absorption: two balls of radius 40, the second 70 along x, smoothness 2, both updaters
    bg = sphere_graph(40.0)
    bvg = SDFVoxelGenerator(1.0, ctx.build_generator(bg), SAME)
    q, t = np.float32([0, 0, 0, 1]), np.float32([-70.0, 0.0, 0.0])

    def mutual_setup():
        a, b = VoxelObject.generate(bvg), VoxelObject.generate(bvg)
        ranges = V.intersection_voxel_ranges(a.info()["occupied_voxel_ranges"], 1.0, b.info()["occupied_voxel_ranges"], 1.0, q, t)
        return a, b, ranges, a.inertial_moments(dens).copy(), b.inertial_moments(dens).copy()

    ms, _ = timed(lambda s: V.absorb_mutually(s[0], s[1], q, t, 2.0, s[2][0], s[2][1], dens, s[3], s[4]), setup=mutual_setup)
    cpu = None
    if O:
        cbg = cpu_generator(bg, SAME)
        ranges = mutual_setup()[2]

        def cpu_mutual_setup():
            a, b = O.Object.generate(cbg, 1), O.Object.generate(cbg, 1)
            return a, b, a.inertial_moments(dens).copy(), b.inertial_moments(dens).copy()

        cpu, _ = cpu_timed(lambda s: O.absorb_mutually(s[0], s[1], q, t, 2.0, ranges[0], ranges[1], dens, s[2], s[3]),
                           setup=cpu_mutual_setup)
    record("apply_mutual_voxel_absorption", "voxel_object.rs:226-268", ms,
           "ivx_objects_absorb_mutually with both inertial updaters (fresh objects per repetition, generated outside "
           "the timed region, where the reference clones inside it)", cpu)

    # split_off_disconnected_region: two balls of radius 50, 120 apart
    tg = SDFGraph()
    tg.union(tg.sphere(50.0), tg.translation(tg.sphere(50.0), [120.0, 0.0, 0.0]), 1.0)
    tvg = SDFVoxelGenerator(1.0, ctx.build_generator(tg), SAME)

    def split(o):
        xi, frag = o.extract_any_disconnected_region()
        assert xi["extracted"]
        return xi

    ms, _ = timed(split, setup=lambda: VoxelObject.generate(tvg))
    cpu = None
    if O:
        ctg = cpu_generator(tg, SAME)
        cpu, _ = cpu_timed(lambda o: o.extract_any_disconnected_region(), setup=lambda: O.Object.generate(ctg, 1))
    record("split_off_disconnected_region", "voxel_object.rs:270-292", ms,
           "ivx_object_extract_disconnected_region on a fresh object (first resolve labels every chunk; the reference "
           "clones inside the timed region instead)", cpu)

    # contacts
    qz = np.float32([0.0, 0.0, np.sin(0.5), np.cos(0.5)])  # from_axis_angle(unit_z, 1.0)
    ms, c = timed(lambda: obj.sphere_contacts(qz, absorber_center, np.float32([0, 0, 0]), absorber_radius))
    cpu = cpu_timed(lambda: cobj.sphere_contacts(qz, absorber_center, np.float32([0, 0, 0]), absorber_radius))[0] if O else None
    record("obtain_sphere_voxel_object_contacts", "voxel_object.rs:365-387", ms, "ivx_object_sphere_contacts, contacts downloaded",
           cpu, contacts=int(len(c)))
    q0 = np.float32([0, 0, 0, 1])
    ms, c = timed(lambda: obj.plane_contacts(q0, center.astype(np.float32), np.float32([0, 1, 0]), -0.92 * R))
    cpu = cpu_timed(lambda: cobj.plane_contacts(q0, center.astype(np.float32), np.float32([0, 1, 0]), -0.92 * R))[0] if O else None
    record("obtain_plane_voxel_object_contacts", "voxel_object.rs:389-408", ms, "ivx_object_plane_contacts, contacts downloaded",
           cpu, contacts=int(len(c)))

    # obtain_mutual_voxel_object_contacts: the big sphere against one of radius 15 sitting on its surface
    small_g = sphere_graph(0.15 * R)
    small = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(small_g), SAME))
    smesh = VoxelObjectMesh.create(small)
    smesh.collision_probes()
    bmesh = VoxelObjectMesh.create(obj)
    bmesh.collision_probes()
    sinfo = small.info()
    socc = np.asarray(sinfo["occupied_voxel_ranges"], np.float64)
    scenter = 0.5 * (socc[:, 0] + socc[:, 1])
    mom_a, mom_b = obj.inertial_moments(dens).copy(), small.inertial_moments(dens).copy()

    def rot(qq, v):
        u, w = np.float64(qq[:3]), float(qq[3])
        return v * (w * w - u @ u) + u * (2.0 * (v @ u)) + np.cross(u, v) * (2.0 * w)

    world_to_a = np.float32([*qz, *absorber_center])
    world_to_b = np.float32([0, 0, 0, 1, *scenter])
    # transform_from_b_to_a = world_to_a * world_to_b.inverted()
    t_b_inv = -np.float64(scenter)
    b_to_a = np.float32([*qz, *(rot(qz, t_b_inv) + np.float64(absorber_center))])
    ranges = V.intersection_voxel_ranges(info["occupied_voxel_ranges"], 1.0, sinfo["occupied_voxel_ranges"], 1.0, b_to_a[:4], b_to_a[4:])
    if ranges is not None:
        ms, (c_ab, c_ba) = timed(lambda: V.mutual_contacts(obj, small, world_to_a, world_to_b, ranges[0], ranges[1], mom_a, mom_b))
        cpu = None
        if O:
            cs = O.Object.generate(cpu_generator(small_g, SAME), 1)
            cpa, cpb = O.CollisionProbes(cobj, cobj.mesh(1)), O.CollisionProbes(cs, cs.mesh(1))
            cpu, _ = cpu_timed(lambda: O.mutual_contacts(cobj, cpa, mom_a, world_to_a, cs, cpb, mom_b, world_to_b, ranges[0], ranges[1]))
        record("obtain_mutual_voxel_object_contacts", "voxel_object.rs:410-448", ms,
               "ivx_objects_mutual_contacts (probes of both objects already computed, like MeshedVoxelObject::create), "
               "contacts downloaded", cpu, contacts=[int(len(c_ab)), int(len(c_ba))])
    else:
        out["benchmarks"]["obtain_mutual_voxel_object_contacts"] = {"note": "the boxes do not intersect"}

    out["not_mirrored"] = ("update_signed_distances_for_block (a CPU SIMD micro-benchmark), clone_object, get_each_voxel, "
                           "for_each_exposed_chunk_with_sdf (host iteration helpers), the Voronoi-fracture benchmarks "
                           "(voxel_object.rs:450-900, SURVEY 8: out of scope)")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
