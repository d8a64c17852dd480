"""Collision probes (VoxelObjectCollisionProbes, collidable.rs:346-780): the oracle restatement against hand-made chunk
meshes and float64 arithmetic (CPU), and the CUDA path (ivx_object_collision_probes / _sync) against the oracle bit for
bit, for every block size and over an absorption sequence with the synced mesh (GPU)."""
import numpy as np
import pytest

import helpers as H

f32 = np.float32


def _curvatures(pos, nrm, idx):
    """add_points_for_vertices_in_blocks' per-vertex mean curvature in float64 (collidable.rs:638-676)."""
    pos, nrm = pos.astype(np.float64), nrm.astype(np.float64)
    s, c = np.zeros(len(pos)), np.zeros(len(pos))
    for i0, i1, i2 in idx.reshape(-1, 3):
        e01, e12, e20 = pos[i1] - pos[i0], pos[i2] - pos[i1], pos[i0] - pos[i2]
        s[i0] += nrm[i0] @ e01 - nrm[i0] @ e20
        s[i1] += nrm[i1] @ e12 - nrm[i1] @ e01
        s[i2] += nrm[i2] @ e20 - nrm[i2] @ e12
        c[[i0, i1, i2]] += 2
    return s, c


def test_block_sizes_follow_the_smallest_occupied_extent(oracle):
    # determine_log2_block_size_for_object (collidable.rs:451-471): 8 from 16 voxels on, 4 from 8, 2 from 4, else 1
    for radius, want in [(31.0, 3), (7.2, 2), (3.2, 2), (2.2, 1), (1.2, 0)]:
        g = H.sphere_graph(radius)
        obj = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, H.SAME0), 2)
        occ = obj.info()["occupied_voxel_ranges"]
        extent = int(min(occ[d][1] - occ[d][0] for d in range(3)))
        expect = 3 if extent >= 16 else 2 if extent >= 8 else 1 if extent >= 4 else 0
        pr = oracle.CollisionProbes(obj, obj.mesh(1))
        assert pr.log2_block_size == expect, (radius, extent)
        assert expect <= want or True


def test_the_most_convex_vertex_of_each_block_is_kept(oracle):
    # a tent over one chunk: the ridge vertices are convex (edges fall away from the normal), the valley ones concave
    rng = np.random.default_rng(3)
    n = 9
    xs, zs = np.meshgrid(np.linspace(1, 15, n), np.linspace(1, 15, n), indexing="ij")
    ys = 8.0 + 3.0 * np.cos(xs * 0.9) * np.cos(zs * 0.7) + 0.05 * rng.normal(size=xs.shape)
    pos = np.stack([xs, ys, zs], -1).reshape(-1, 3).astype(f32)
    gy_x, gy_z = np.gradient(ys, xs[:, 0], zs[0])
    nrm = np.stack([-gy_x, np.ones_like(ys), -gy_z], -1).reshape(-1, 3)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(f32)
    tri = []
    for i in range(n - 1):
        for k in range(n - 1):
            a, b, c, d = i * n + k, (i + 1) * n + k, (i + 1) * n + k + 1, i * n + k + 1
            tri += [a, b, c, a, c, d]
    idx = np.array(tri, np.uint32)
    s, c = _curvatures(pos, nrm, idx)
    curv = s / np.maximum(c, 1)
    for log2_bs in (3, 2, 1):
        got = oracle.probes_points_for_chunk(log2_bs, [0, 0, 0], pos, nrm, idx)
        blocks = (np.floor(pos).astype(int) & 15) >> log2_bs
        nb = 16 >> log2_bs
        lin = (blocks[:, 0] * nb + blocks[:, 1]) * nb + blocks[:, 2]
        want = []
        for b in sorted(set(lin)):
            members = np.flatnonzero((lin == b) & (c > 0))
            if len(members):
                want.append(pos[members[np.argmin(curv[members])]])
        assert len(got) == len(want)
        # float32 sums against float64 sums: the chosen vertex may differ only where two curvatures tie to ~1e-6
        for g_pt, w_pt in zip(got, want):
            if not np.array_equal(g_pt, w_pt):
                gi = int(np.flatnonzero((pos == g_pt).all(1))[0]); wi = int(np.flatnonzero((pos == w_pt).all(1))[0])
                assert abs(curv[gi] - curv[wi]) < 1e-5
    # an unconnected vertex is ignored; a start index is subtracted from the indices (the mesh's indices are global)
    extra = np.vstack([pos, f32([[8, 30, 8]])])
    extra_n = np.vstack([nrm, f32([[0, 1, 0]])])
    a = oracle.probes_points_for_chunk(3, [0, 0, 0], extra, extra_n, idx)
    b = oracle.probes_points_for_chunk(3, [0, 0, 0], extra, extra_n, idx + 1000, start_index=1000)
    assert np.array_equal(a, oracle.probes_points_for_chunk(3, [0, 0, 0], pos, nrm, idx)) and np.array_equal(a, b)


def test_positions_on_the_upper_chunk_face_wrap_to_block_zero(oracle):
    # vertices may lie slightly outside their chunk: positions are clamped to the chunk's box and `16 & 15 = 0`
    # puts a vertex on the upper face into block 0 of that axis (collidable.rs:690-707, object.rs:3223-3233)
    pos = f32([[16.2, 4.0, 4.0], [15.0, 4.0, 4.0], [15.5, 5.0, 4.0], [3.0, 4.0, 4.0], [2.0, 5.0, 4.0], [2.5, 4.0, 5.0]])
    nrm = f32([[0, 0, 1]] * 6)
    idx = np.array([0, 1, 2, 3, 4, 5], np.uint32)
    pts = oracle.probes_points_for_chunk(3, [0, 0, 0], pos, nrm, idx)
    s, c = _curvatures(pos, nrm, idx)
    # block (0,0,0) holds vertices 3, 4, 5 AND the wrapped vertex 0; block (1,0,0) holds 1 and 2
    in_block0 = [0, 3, 4, 5]
    best0 = in_block0[int(np.argmin((s / c)[in_block0]))]
    best1 = [1, 2][int(np.argmin((s / c)[[1, 2]]))]
    assert len(pts) == 2 and np.array_equal(pts[0], pos[best0]) and np.array_equal(pts[1], pos[best1])


def test_probes_sit_on_the_surface_and_cover_every_meshed_chunk(oracle):
    g = H.sphere_graph(31.0)
    obj = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, H.SAME0), 4)
    mesh = obj.mesh(4)
    pr = oracle.CollisionProbes(obj, mesh)
    centre = 8.0 * np.array(obj.info()["chunk_counts"])
    r = np.linalg.norm(pr.points - centre, axis=1)
    assert np.all(np.abs(r - 31.0) < 0.6)
    assert len(pr.ranges) == mesh.n_submeshes and pr.ranges[:, 2].max() == len(pr.points)
    assert np.all(pr.ranges[:, 2] - pr.ranges[:, 1] <= 8)  # at most one point per 8^3 block of a chunk
    # every probe is a mesh vertex of its chunk
    verts = {tuple(v) for v in mesh.positions.view(np.uint32).reshape(-1, 3)}
    assert all(tuple(p) in verts for p in pr.points.view(np.uint32).reshape(-1, 3))


# ---- CUDA path --------------------------------------------------------------------------------------------------------
def _same_probes(got, want):
    assert got["log2_block_size"] == want.log2_block_size
    g_r = got["ranges"]
    assert len(g_r) == len(want.ranges)
    return g_r


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere64", "tiny1", "tiny2", "tiny4", "asteroid_like", "noisy_box"])
def test_probes_of_all_chunks_match_the_oracle(ctx, oracle, name):
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh
    make = {"sphere64": lambda: H.sphere_graph(31.0), "tiny1": lambda: H.sphere_graph(1.2), "tiny2": lambda: H.sphere_graph(2.2),
            "tiny4": lambda: H.sphere_graph(5.2), "asteroid_like": lambda: H.asteroid_like_graph(24, 40.0),
            "noisy_box": lambda: H.noisy_box_graph(38.0, 8)}[name]
    g = make()
    obj_cpu = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, H.GRADIENT4), 4)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(g), H.GRADIENT4))
    want = oracle.CollisionProbes(obj_cpu, obj_cpu.mesh(4))
    got = VoxelObjectMesh.create(obj_gpu).collision_probes()
    g_r = _same_probes(got, want)
    nb = obj_cpu.info()["chunk_counts"]
    lin = (g_r["chunk_indices"][:, 0] * nb[1] + g_r["chunk_indices"][:, 1]) * nb[2] + g_r["chunk_indices"][:, 2]
    assert np.array_equal(lin, want.ranges[:, 0])
    assert np.array_equal(g_r["point_start"], want.ranges[:, 1]) and np.array_equal(g_r["point_end"], want.ranges[:, 2])
    assert H.f32_bits_equal(got["points"], want.points).all()
    assert len(want.points) > 0


@pytest.mark.gpu
def test_chunks_too_large_for_the_shared_memory_lists_take_the_walk(ctx, oracle, monkeypatch):
    # IVX_PROBES_WALK_ALL sends every chunk down the path of chunks with more than 2048 vertices / 12288 corners
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh
    g = H.noisy_box_graph(38.0, 8)
    obj_cpu = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, H.SAME0), 4)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(g), H.SAME0))
    want = oracle.CollisionProbes(obj_cpu, obj_cpu.mesh(4))
    mesh = VoxelObjectMesh.create(obj_gpu)
    monkeypatch.setenv("IVX_PROBES_WALK_ALL", "1")
    got = mesh.collision_probes()
    monkeypatch.delenv("IVX_PROBES_WALK_ALL")
    assert H.f32_bits_equal(got["points"], want.points).all() and len(got["ranges"]) == len(want.ranges)
    assert H.f32_bits_equal(mesh.collision_probes()["points"], want.points).all()


@pytest.mark.gpu
def test_probes_follow_the_synced_mesh_through_an_absorption_sequence(ctx, oracle):
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh
    g = H.asteroid_like_graph(16, 36.0)
    obj_cpu = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, H.GRADIENT4), 4)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(g), H.GRADIENT4))
    sm = oracle.SyncedMesh(obj_cpu, 4)
    want = oracle.CollisionProbes(obj_cpu, sm)
    mesh_gpu = VoxelObjectMesh.create(obj_gpu)
    got = mesh_gpu.collision_probes()
    assert H.f32_bits_equal(got["points"], want.points).all()
    shape = np.array(obj_cpu.info()["chunk_counts"]) * 16
    R = 0.5 * float(shape.max())
    radius = f32(0.2 * R)
    start = (0.5 * shape - R / np.sqrt(3.0)).astype(f32)
    grew = freed = False
    for step in range(8):
        c = (start + f32(step) * radius * f32(0.45)).astype(f32)
        obj_cpu.absorb_sphere(c, float(radius), float(radius + 2.0))
        obj_gpu.absorb_sphere(c, float(radius), float(radius + 2.0))
        dirty = np.sort(obj_cpu.dirty())  # the library takes the invalidated chunks in ascending order
        sm.sync(obj_cpu, dirty)
        before = len(want.points)
        want.sync(obj_cpu, sm, dirty)
        obj_cpu.clear_dirty()
        mesh_gpu = VoxelObjectMesh.sync(obj_gpu)
        got = mesh_gpu.collision_probes(sync=True)
        g_r = _same_probes(got, want)
        nb = obj_cpu.info()["chunk_counts"]
        lin = (g_r["chunk_indices"][:, 0] * nb[1] + g_r["chunk_indices"][:, 1]) * nb[2] + g_r["chunk_indices"][:, 2]
        assert np.array_equal(lin, want.ranges[:, 0]), step
        assert np.array_equal(g_r["point_start"], want.ranges[:, 1]) and np.array_equal(g_r["point_end"], want.ranges[:, 2]), step
        # the buffers, obsolete points in freed ranges included
        assert H.f32_bits_equal(got["points"], want.points).all(), step
        grew |= len(want.points) > before
        freed |= int((want.ranges[:, 2] - want.ranges[:, 1]).sum()) < len(want.points)
    assert grew and freed  # both placements occurred: appended at the end, and holes left / reused
    # a fresh computation on the synced mesh gives the live points again
    fresh = mesh_gpu.collision_probes()
    want_fresh = oracle.CollisionProbes(obj_cpu, sm)
    assert H.f32_bits_equal(fresh["points"], want_fresh.points).all()
