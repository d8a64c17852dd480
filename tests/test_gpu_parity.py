"""GPU parity tests proper: the CUDA library (through its C ABI) against the CPU oracle on the
same inputs — bit-exact for voxels (type, signed-distance code, flags), chunk kinds / flags / face
distributions, mesh topology, index materials, and f32-bit-exact for SDF values, vertex positions
and normals. At full BASELINE sizes, size-independent properties replace the oracle."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import workloads as W
import invariants as INV
from impact_b200.graph import VoxelTypeGenerator
from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

pytestmark = pytest.mark.gpu


def _both(ctx, oracle, graph, types=H.SAME0, extent=1.0, threads=4):
    gen_gpu = ctx.build_generator(graph)
    gen_cpu = oracle.Generator(graph.nodes(), graph.root_node_id)
    vg_cpu = oracle.VoxelGenerator(gen_cpu, extent, types)
    obj_cpu = oracle.Object.generate(vg_cpu, threads)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(extent, gen_gpu, types))
    return gen_gpu, gen_cpu, vg_cpu, obj_gpu, obj_cpu


GRAPHS = {
    "sphere64": (lambda: H.sphere_graph(31.0), H.SAME0),                 # BASELINE config 1
    "sphere_big_interior": (lambda: H.sphere_graph(70.0), H.SAME0),      # chunks filled with -margin
    "box": (lambda: H.box_graph(45.0), H.SAME0),
    "box_types": (lambda: H.box_graph(45.0), H.GRADIENT4),               # generate_box_with_gradient_noise_voxel_types
    "sphere_union": (lambda: H.sphere_union_graph(0.5), H.SAME0),
    "complex": (lambda: H.complex_graph(0.6), H.SAME0),
    "noisy_sphere": (lambda: H.noisy_sphere_graph(30.0, 4), H.SAME0),
    "noisy_box": (lambda: H.noisy_box_graph(38.0, 8), H.SAME0),           # BASELINE config 2, reduced
    "zoo": (H.csg_zoo_graph, H.GRADIENT4),
    "mid_noise": (H.mid_noise_graph, H.SAME0),
    # an editor file (tests/golden/mini.graph.ron) through the .graph.ron loader and the meta-graph compiler
    "editor_mini": (lambda: __import__("impact_b200.meta", fromlist=["x"]).compile_graph_file(
        __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "mini.graph.ron"))[0], H.GRADIENT4),
    "asteroid_like": (lambda: H.asteroid_like_graph(24, 40.0), H.GRADIENT4),  # BASELINE config 3 stand-in
    # same node kinds / counts as asteroid.vgen.ron (446 leaves, depth 9) at a grid the oracle finishes in seconds
    "asteroid_stand_in": (lambda: W.asteroid_stand_in(0.6), H.GRADIENT4),
    "asteroid_stand_in_same": (lambda: W.asteroid_stand_in(0.45, seed=2), H.SAME0),
}


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_signed_distances_per_chunk_are_bit_exact(ctx, oracle, name):
    make, types = GRAPHS[name]
    g = make()
    gen_gpu = ctx.build_generator(g)
    gen_cpu = oracle.Generator(g.nodes(), g.root_node_id)
    vg = oracle.VoxelGenerator(gen_cpu, 1.0, types)
    cc = [(s + 15) // 16 for s in vg.grid_shape]
    origins = np.array([[i, j, k] for i in range(cc[0]) for j in range(cc[1]) for k in range(cc[2])], np.float32) * 16
    rng = np.random.default_rng(1)
    if len(origins) > 160:
        origins = origins[rng.choice(len(origins), 160, replace=False)]
    lo = origins - vg.shifted_center  # generation.rs:314-315
    # plus unaligned origins, as the meta compiler's probes use
    lo = np.concatenate([lo, lo[:8] + rng.uniform(-3, 3, (8, 3)).astype(np.float32)])
    got = gen_gpu.compute_signed_distances_for_chunks(lo)
    for n, o in enumerate(lo):
        want, _ = gen_cpu.eval_chunk(o)
        eq = H.f32_bits_equal(got[n], want)
        assert eq.all(), (name, n, o, np.flatnonzero(~eq)[:5], got[n][~eq][:5], want[~eq][:5])


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_generated_object_and_mesh_are_bit_exact(ctx, oracle, name):
    make, types = GRAPHS[name]
    _, _, vg_cpu, obj_gpu, obj_cpu = _both(ctx, oracle, make(), types)
    gi, ci = obj_gpu.info(), obj_cpu.info()
    assert gi["grid_shape"] == vg_cpu.grid_shape
    assert gi["chunk_counts"] == ci["chunk_counts"]
    assert np.array_equal(gi["occupied_voxel_ranges"], ci["occupied_voxel_ranges"])
    assert np.array_equal(gi["occupied_chunk_ranges"], ci["occupied_chunk_ranges"])
    gch, gvx = obj_gpu.download()
    H.assert_objects_equal(gch, gvx, obj_cpu.chunks(), obj_cpu.voxels())
    assert gi["n_non_uniform"] == int((gch["kind"] == 2).sum())
    # the reference's own invariants on the GPU output
    INV.validate_adjacencies(gch, gvx, gi["chunk_counts"])
    INV.validate_chunk_obscuredness(gch, gi["chunk_counts"])
    INV.validate_occupied_voxel_ranges(gch, gvx, gi["chunk_counts"], gi["occupied_voxel_ranges"])
    mesh = VoxelObjectMesh.create(obj_gpu)
    H.assert_meshes_equal(mesh.download(), obj_cpu.mesh(4))


@pytest.mark.parametrize("label,types", [
    ("one_type", VoxelTypeGenerator.gradient_noise([5], 0.02, 1.0, 0)),
    ("nine_types_seeded", VoxelTypeGenerator.gradient_noise(list(range(9)), 0.031, 0.37, 12345)),
    # 70 types: their lattice tables do not fit one batch; type coordinates beyond the first 8-wide vector
    ("seventy_types", VoxelTypeGenerator.gradient_noise(list(range(70)), 0.02, 1.0, 3)),
    # noise cells much smaller than a chunk: the gradient table does not fit, the direct evaluation runs
    ("high_frequency", VoxelTypeGenerator.gradient_noise([0, 1, 2, 3, 4], 0.9, 1.0, 1)),
    ("medium_frequency", VoxelTypeGenerator.gradient_noise([0, 1, 2], 0.21, 2.5, 7)),
])
def test_voxel_type_generators_are_bit_exact(ctx, oracle, label, types):
    # voxel_type.rs:54-168: every path of k_types (one batch, several batches, direct evaluation)
    g = H.sphere_graph(21.0)
    _, _, _, obj_gpu, obj_cpu = _both(ctx, oracle, g, types)
    gch, gvx = obj_gpu.download()
    H.assert_objects_equal(gch, gvx, obj_cpu.chunks(), obj_cpu.voxels())
    if label != "one_type":
        assert len(np.unique(gvx["type"][(gvx["flags"] & 1) == 0])) > 1
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh(4))


@pytest.mark.parametrize("name", sorted(H.GOLDEN_OBJECTS))
def test_cuda_objects_match_the_committed_digests(ctx, name):
    # the same committed fixtures the oracle is frozen by (tests/golden/object_digests.json), without the oracle in
    # the loop: chunk table, every voxel, and the mesh buffers in the reference's order
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "object_digests.json")) as f:
        want = json.load(f)[name]
    make, types_name = H.GOLDEN_OBJECTS[name]
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(make()), getattr(H, types_name)))
    assert list(obj.info()["chunk_counts"]) == want["chunk_counts"]
    assert H.object_digest(*obj.download()) == want["object"]
    m = VoxelObjectMesh.create(obj).download()
    assert (len(m["positions"]), len(m["indices"])) == (want["vertices"], want["indices"])
    assert H.mesh_digest(m["positions"], m["normals"], m["indices"], m["index_materials"], m["submeshes"],
                         m["vertex_ranges"]) == want["mesh"]


def test_voxel_extent_scales_positions_only(ctx, oracle):
    _, _, _, obj_gpu, obj_cpu = _both(ctx, oracle, H.complex_graph(0.4), H.SAME0, extent=0.25)
    gch, gvx = obj_gpu.download()
    H.assert_objects_equal(gch, gvx, obj_cpu.chunks(), obj_cpu.voxels())
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh(2))


def test_empty_graph_and_degenerate_objects(ctx, oracle):
    g = H.SDFGraph()
    gen = ctx.build_generator(g)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, H.SAME0))
    assert obj.info()["grid_shape"] == (0, 0, 0)
    m = VoxelObjectMesh.create(obj)
    assert m.n_vertices == 0 and m.n_indices == 0
    # a sphere smaller than a voxel: grid exists, nothing (or next to nothing) inside
    _, _, _, obj_gpu, obj_cpu = _both(ctx, oracle, H.sphere_graph(0.3))
    gch, gvx = obj_gpu.download()
    H.assert_objects_equal(gch, gvx, obj_cpu.chunks(), obj_cpu.voxels())
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh())


def test_graph_errors_surface_through_the_abi(ctx):
    from impact_b200._lib import IvxError

    g = H.SDFGraph()
    g.add_node((7, (0, 0), 0, 0, [0.0]))
    with pytest.raises(IvxError, match="cycle"):
        ctx.build_generator(g)


def test_uploading_a_host_compiled_program_equals_building_it(ctx, oracle):
    from impact_b200.voxel import SDFGenerator, compile_program_host

    g = H.csg_zoo_graph()
    nodes, depth, lo, hi = compile_program_host(g)
    gen = SDFGenerator.from_processed_nodes(ctx, nodes, depth, lo, hi)
    a = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, H.GRADIENT4)).download()
    b = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(g), H.GRADIENT4)).download()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("name,types", [("sphere", H.SAME0), ("asteroid_like", H.GRADIENT4)])
def test_absorption_and_dirty_remesh_are_bit_exact(ctx, oracle, name, types):
    g = H.sphere_graph(40.0) if name == "sphere" else H.asteroid_like_graph(16, 36.0)
    _, _, _, obj_gpu, obj_cpu = _both(ctx, oracle, g, types)
    cc = obj_cpu.info()["chunk_counts"]
    shape = np.array(cc) * 16
    rng = np.random.default_rng(5)
    # BASELINE config 5 geometry: sphere of radius 0.15 R marching inward along the diagonal
    R = 0.5 * float(shape.max())
    radius = np.float32(0.15 * R)
    center = (0.5 * shape - R / np.sqrt(3.0)).astype(np.float32)
    for step in range(6):
        c = (center + step * radius * np.float32(0.6) + rng.uniform(-0.3, 0.3, 3)).astype(np.float32)
        st_c = obj_cpu.absorb_sphere(c, float(radius), float(radius + np.float32(2.0)))
        st_g = obj_gpu.absorb_sphere(c, float(radius), float(radius + np.float32(2.0)))
        for f in ("touched_chunks", "touched_voxels", "emptied_voxels", "removed_chunks"):
            assert st_g[f] == st_c[f], (step, f, st_g, st_c)
        gch, gvx = obj_gpu.download()
        H.assert_objects_equal(gch, gvx, obj_cpu.chunks(), obj_cpu.voxels())
        assert np.array_equal(obj_gpu.info()["occupied_voxel_ranges"], obj_cpu.info()["occupied_voxel_ranges"])
        dirty_g = np.sort(obj_gpu.invalidated_mesh_chunk_indices())
        dirty_c = np.sort(obj_cpu.dirty())
        assert np.array_equal(dirty_g, dirty_c), step
        if step % 2 == 1:
            patch = VoxelObjectMesh.sync_with_voxel_object(obj_gpu)
            pm = patch.download()
            seen = set()
            for s, r in zip(pm["submeshes"], pm["vertex_ranges"]):
                idx3 = tuple(int(x) for x in s["chunk_indices"])
                seen.add(idx3)
                cm = obj_cpu.mesh_chunk(*idx3)
                assert cm is not None, idx3
                assert H.f32_bits_equal(pm["positions"][r[0]: r[1]], cm["positions"]).all()
                assert H.f32_bits_equal(pm["normals"][r[0]: r[1]], cm["normals"]).all()
                sl = slice(s["index_offset"], s["index_offset"] + s["index_count"])
                assert np.array_equal(pm["indices"][sl] - r[0], cm["indices"].astype(np.uint32))
                assert np.array_equal(pm["index_materials"][sl], cm["index_materials"])
            for lin in dirty_c:
                idx3 = (int(lin // (cc[1] * cc[2])), int((lin // cc[2]) % cc[1]), int(lin % cc[2]))
                if idx3 not in seen:
                    assert obj_cpu.mesh_chunk(*idx3) is None, idx3
            obj_cpu.clear_dirty()
            assert len(obj_gpu.invalidated_mesh_chunk_indices()) == 0
    # after all modifications a full re-mesh still agrees
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh(4))


@pytest.mark.parametrize("name,types", [("sphere", H.SAME0), ("asteroid_like", H.GRADIENT4)])
def test_capsule_absorption_is_bit_exact(ctx, oracle, name, types):
    # apply_capsule_absorption (absorption.rs:846-889) → modify_voxels_within_capsule (intersection.rs:417-537):
    # slanted, axis-aligned, zero-length, grazing and missing capsules; uniform interior chunks get converted
    g = H.sphere_graph(40.0) if name == "sphere" else H.asteroid_like_graph(16, 36.0)
    _, _, _, obj_gpu, obj_cpu = _both(ctx, oracle, g, types)
    cc = obj_cpu.info()["chunk_counts"]
    shape = (np.array(cc) * 16).astype(np.float32)
    f = np.float32
    mid = f(0.5) * shape
    capsules = [
        (mid - f([30.5, 8.25, 3.0]), f([61.0, 16.5, 6.0]), f(5.0)),      # slanted, through the centre
        (f([mid[0], 2.3, mid[2]]), f([0.0, shape[1] - 4.0, 0.0]), f(3.5)),  # axis aligned (zero components)
        (mid + f([9.1, -7.7, 11.3]), f([0.0, 0.0, 0.0]), f(6.0)),        # zero-length segment
        (f([-6.0, mid[1], mid[2]]), f([14.0, 3.0, -2.0]), f(4.0)),       # enters from outside the grid
        (f([-50.0, -50.0, -50.0]), f([10.0, 0.0, 0.0]), f(4.0)),         # misses the object
        (mid - f([3.0, 40.0, 1.0]), f([5.5, 80.0, 2.5]), f(9.0)),        # thick, crosses everything again
    ]
    for step, (a, v, radius) in enumerate(capsules):
        st_c = obj_cpu.absorb_capsule(a, v, float(radius), float(radius + f(2.0)))
        st_g = obj_gpu.absorb_capsule(a, v, float(radius), float(radius + f(2.0)))
        for k in ("touched_chunks", "touched_voxels", "emptied_voxels", "removed_chunks"):
            assert st_g[k] == st_c[k], (step, k, st_g, st_c)
        gch, gvx = obj_gpu.download()
        H.assert_objects_equal(gch, gvx, obj_cpu.chunks(), obj_cpu.voxels())
        assert np.array_equal(obj_gpu.info()["occupied_voxel_ranges"], obj_cpu.info()["occupied_voxel_ranges"])
        assert np.array_equal(np.sort(obj_gpu.invalidated_mesh_chunk_indices()), np.sort(obj_cpu.dirty())), step
    assert st_c["touched_voxels"] > 0
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh(4))
    with pytest.raises(Exception):
        obj_gpu.absorb_capsule(mid, f([1, 0, 0]), -1.0, 1.0)


@pytest.mark.parametrize("name", ["sphere_big_interior", "asteroid_like", "asteroid_stand_in", "box_types"])
def test_streamed_generation_fills_host_buffers_like_generate_plus_download(ctx, name):
    # ivx_object_generate_streamed: planes generated in parts, each part's cross-chunk state, packing and copy
    # overlapped with the next parts; the result must be the bytes of generate → download, and the object left on
    # the device must mesh identically
    import torch
    from impact_b200 import _lib as L
    make, types = GRAPHS[name]
    vg = SDFVoxelGenerator(1.0, ctx.build_generator(make()), types)
    ref = VoxelObject.generate(vg)
    rch, rvx = ref.download()
    n_chunks = len(rch)
    h_chunks = torch.empty(n_chunks * 16, dtype=torch.uint8, pin_memory=True).numpy()
    h_vox = torch.empty(max(1, len(rvx)) * 3 + 4096 * 3, dtype=torch.uint8, pin_memory=True).numpy()
    h_chunks[:] = 0xAB
    h_vox[:] = 0xCD
    obj, nnu = VoxelObject.generate_streamed(vg, h_chunks, h_vox)
    ctx.synchronize()
    assert nnu * 4096 == len(rvx)
    assert np.array_equal(h_chunks.view(L.CHUNK_DTYPE), rch)
    assert np.array_equal(h_vox[: len(rvx) * 3].view(L.VOXEL_DTYPE), rvx)
    assert (h_vox[len(rvx) * 3:] == 0xCD).all()  # nothing written past the end
    gch, gvx = obj.download()
    assert np.array_equal(gch, rch) and np.array_equal(gvx, rvx)
    ia, ib = obj.info(), ref.info()
    for k in ib:
        if k != "device_bytes":
            assert np.array_equal(ia[k], ib[k]), k
    ma, mb = VoxelObjectMesh.create(obj).download(), VoxelObjectMesh.create(ref).download()
    for k in ma:
        assert np.array_equal(np.asarray(ma[k]).view(np.uint8), np.asarray(mb[k]).view(np.uint8)), k
    # too small a voxel buffer is reported, not overrun
    if len(rvx):
        small = torch.empty(4096 * 3, dtype=torch.uint8, pin_memory=True).numpy()
        with pytest.raises(Exception):
            VoxelObject.generate_streamed(vg, h_chunks, small)
        ctx.synchronize()


def test_full_size_sphere_properties(ctx):
    # engine bench shape: Sphere(r = 100) → 202³ (benchmarks/voxel_object.rs:650-660); no oracle here,
    # only size-independent properties: reference invariants + closed manifold + radius
    gen = ctx.build_generator(H.sphere_graph(100.0))
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, H.SAME0))
    info = obj.info()
    assert info["grid_shape"] == (202, 202, 202)
    ch, vx = obj.download()
    INV.validate_adjacencies(ch, vx, info["chunk_counts"])
    INV.validate_chunk_obscuredness(ch, info["chunk_counts"])
    INV.validate_occupied_voxel_ranges(ch, vx, info["chunk_counts"], info["occupied_voxel_ranges"])
    m = VoxelObjectMesh.create(obj).download()
    chi, hist = H.euler_characteristic(m["positions"], m["indices"])
    assert chi == 2 and hist[1] == 0 and len(hist) == 3
    r = np.linalg.norm(m["positions"] - 101.0, axis=1)
    assert abs(r - 100.0).max() < 0.05


def test_generation_is_deterministic_and_pool_reuse_is_clean(ctx):
    gen = ctx.build_generator(H.csg_zoo_graph())
    vg = SDFVoxelGenerator(1.0, gen, H.GRADIENT4)
    a = VoxelObject.generate(vg)
    ra = a.download()
    ma = VoxelObjectMesh.create(a).download()
    a.free()
    b = VoxelObject.generate(vg)  # reuses the freed device blocks
    rb = b.download()
    mb = VoxelObjectMesh.create(b).download()
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    for k in ma:
        assert np.array_equal(ma[k].view(np.uint8), mb[k].view(np.uint8)), k


def test_absorption_over_a_box_of_many_chunks_is_bit_exact(ctx, oracle):
    # a sphere of influence reaching over 24^3 chunks: the boundary refresh of a box this large (more chunks than
    # BOUNDARY_BOX_ONE_CTA, csrc/kernels.h) is prepared by a grid and the library's prefix sum instead of one CTA;
    # uniform chunks deep inside are converted, many chunks are emptied and removed
    g = H.sphere_graph(190.0)
    _, _, _, obj_gpu, obj_cpu = _both(ctx, oracle, g, H.SAME0, threads=8)
    shape = np.array(obj_cpu.info()["chunk_counts"]) * 16
    for c, r in ((0.5 * shape + np.float32([20.0, -10.0, 5.0]), 170.0), (0.5 * shape, 185.0)):
        c = np.float32(c)
        st_c = obj_cpu.absorb_sphere(c, r, r + 2.0)
        st_g = obj_gpu.absorb_sphere(c, r, r + 2.0)
        assert st_c["touched_chunks"] > 3000  # (the refreshed box around them: 23^3 - 24^3 chunks)
        for f in ("touched_chunks", "touched_voxels", "emptied_voxels", "removed_chunks"):
            assert st_g[f] == st_c[f], (f, st_g, st_c)
        gch, gvx = obj_gpu.download()
        H.assert_objects_equal(gch, gvx, obj_cpu.chunks(), obj_cpu.voxels())
        assert np.array_equal(obj_gpu.info()["occupied_voxel_ranges"], obj_cpu.info()["occupied_voxel_ranges"])
        assert np.array_equal(np.sort(obj_gpu.invalidated_mesh_chunk_indices()), np.sort(obj_cpu.dirty()))
