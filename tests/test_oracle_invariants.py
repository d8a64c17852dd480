"""The reference's structural tests and fuzz invariants run against the oracle. CPU only."""
import numpy as np
import pytest

import helpers as H
import invariants as INV


def _gen(oracle, graph, types=H.SAME0, extent=1.0, threads=1):
    gen = oracle.Generator(graph.nodes(), graph.root_node_id)
    vg = oracle.VoxelGenerator(gen, extent, types)
    return vg, oracle.Object.generate(vg, threads)


def _check_all(oracle, obj):
    info = obj.info()
    ch, vx = obj.chunks(), obj.voxels()
    cc = info["chunk_counts"]
    INV.validate_occupied_voxel_ranges(ch, vx, cc, info["occupied_voxel_ranges"])
    INV.validate_adjacencies(ch, vx, cc)
    INV.validate_chunk_obscuredness(ch, cc)
    return ch, vx, cc


@pytest.mark.parametrize("name", ["sphere", "box", "capsule", "union", "complex", "zoo"])
def test_generated_objects_satisfy_reference_invariants(oracle, name):
    # fuzz targets of the reference: random single primitives + validate_* (object.rs:3371-3377)
    g = {"sphere": lambda: H.sphere_graph(23.3), "box": lambda: H.box_graph(37.0),
         "capsule": lambda: _capsule(), "union": lambda: H.sphere_union_graph(0.35),
         "complex": lambda: H.complex_graph(0.4), "zoo": H.csg_zoo_graph}[name]()
    _, obj = _gen(oracle, g)
    ch, vx, cc = _check_all(oracle, obj)
    assert (ch["kind"] == 2).any()


def _capsule():
    g = H.SDFGraph()
    g.capsule(30.0, 9.0)
    return g


def test_grid_shape_and_centre_follow_sdf_voxel_generator_new(oracle):
    # generation.rs:230-258: ceil(extent) + 2 per axis; config 1: Sphere(31) → 64³
    vg, _ = _gen(oracle, H.sphere_graph(31.0))
    assert vg.grid_shape == (64, 64, 64)
    assert np.array_equal(vg.shifted_center, np.float32([31.5, 31.5, 31.5]))
    vg, _ = _gen(oracle, H.noisy_box_graph(22.0, 2))
    assert vg.grid_shape == (32, 32, 32)  # 22 + 2*4 (noise amplitude) + 2


def test_empty_graph_generates_empty_object(oracle):
    g = H.SDFGraph()
    gen = oracle.Generator(g.nodes(), 0)
    vg = oracle.VoxelGenerator(gen, 1.0, H.SAME0)
    assert vg.grid_shape == (0, 0, 0)
    obj = oracle.Object.generate(vg)
    assert obj.info()["n_voxels"] == 0


def test_brick_sign_and_type_invariant(oracle):
    # validate_sdf (object/sdf.rs:510-571)
    _, obj = _gen(oracle, H.complex_graph(0.3), H.GRADIENT4)
    info = obj.info()
    ch, vx = obj.chunks(), obj.voxels()
    cc = info["chunk_counts"]
    dense = INV.dense_fields(ch, vx, cc)
    checked = 0
    for c in range(len(ch)):
        idx3 = (c // (cc[1] * cc[2]), (c // cc[2]) % cc[1], c % cc[2])
        b = obj.fill_brick(*idx3)
        if b is None:
            continue
        if checked < 6:
            INV.validate_brick(b[0], b[1], b[2], ch[c], idx3, ch, vx, cc, dense)
        checked += 1
    assert checked > 0


def test_sphere_mesh_is_a_closed_manifold(oracle):
    _, obj = _gen(oracle, H.sphere_graph(31.0))
    m = obj.mesh()
    chi, hist = H.euler_characteristic(m.positions, m.indices)
    assert chi == 2
    assert len(hist) == 3 and hist[1] == 0  # every edge used by exactly two triangles
    r = np.linalg.norm(m.positions - 32.0, axis=1)
    assert abs(r - 31.0).max() < 0.05
    # normals point outward
    n_dot = ((m.positions - 32.0) * m.normals).sum(1)
    assert (n_dot > 0).all()


def test_parallel_generation_and_meshing_equal_serial(oracle):
    # object.rs:406-558: contiguous chunk ranges per worker give the same object
    vg, a = _gen(oracle, H.csg_zoo_graph(), H.GRADIENT4, threads=1)
    b = oracle.Object.generate(vg, 5)
    assert np.array_equal(a.chunks(), b.chunks())
    assert np.array_equal(a.voxels(), b.voxels())
    ma, mb = a.mesh(1), b.mesh(4)
    assert np.array_equal(ma.indices, mb.indices)
    assert np.array_equal(ma.positions.view(np.uint32), mb.positions.view(np.uint32))


def test_sphere_modification_touches_exactly_the_brute_force_voxel_set(oracle):
    # intersection.rs:1093-1348: voxels visited == voxels whose centre is strictly inside the sphere,
    # restricted to the occupied voxel ranges
    _, obj = _gen(oracle, H.sphere_graph(23.3))
    info = obj.info()
    occ = info["occupied_voxel_ranges"].astype(np.int64)
    before = obj.voxels().copy()
    ch_before = obj.chunks().copy()
    center, radius = np.float32([30.2, 24.9, 11.3]), 9.5
    st = obj.absorb_sphere(center, radius - 2.0, radius)
    idx = np.stack(np.meshgrid(*[np.arange(occ[d, 0], occ[d, 1]) for d in range(3)], indexing="ij"), -1).reshape(-1, 3)
    p = idx.astype(np.float32) + np.float32(0.5)
    d = center - p
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    inside = d2 < np.float32(radius) * np.float32(radius)
    # voxels inside void chunks are skipped (intersection.rs:314-316)
    cc = info["chunk_counts"]
    cidx = ((idx[:, 0] // 16) * cc[1] + idx[:, 1] // 16) * cc[2] + idx[:, 2] // 16
    not_void = ch_before["kind"][cidx] != 0
    assert st["touched_voxels"] == int((inside & not_void).sum())
    assert st["touched_voxels"] > 0 and st["emptied_voxels"] > 0
    _check_all(oracle, obj)
    assert len(obj.dirty()) >= st["touched_chunks"]
    assert before.shape[0] <= obj.voxels().shape[0]


def test_capsule_modification_touches_exactly_the_brute_force_voxel_set(oracle):
    # intersection.rs:955-982 (fuzz_test_obtaining_voxels_within_capsule): voxels visited == voxels whose centre lies
    # within or on the capsule (CapsulePointContainmentTester, capsule.rs:225-250), restricted to the occupied ranges
    _, obj = _gen(oracle, H.sphere_graph(23.3))
    info = obj.info()
    occ = info["occupied_voxel_ranges"].astype(np.int64)
    ch_before = obj.chunks().copy()
    f = np.float32
    start, vec, radius = f([4.3, 30.1, 9.7]), f([37.0, -11.5, 21.25]), f(6.5)
    st = obj.absorb_capsule(start, vec, float(radius - f(2.0)), float(radius))
    # whole chunks of the touched chunk range are visited: the per-chunk voxel range comes from the capsule trimmed
    # to the chunk, not from the occupied voxel ranges (intersection.rs:445-461), so empty voxels beyond the occupied
    # ranges are visited too (the reference's fuzz test only counts non-empty ones)
    end = start + vec
    lo = np.maximum(np.floor(np.minimum(start, end) - radius), 0).astype(np.int64)
    hi = np.ceil(np.maximum(start, end) + radius).astype(np.int64)
    tlo, thi = np.maximum(occ[:, 0], lo), np.minimum(occ[:, 1], hi)
    assert np.all(tlo < thi)
    clo, chi = tlo // 16, (thi + 15) // 16
    idx = np.stack(np.meshgrid(*[np.arange(clo[d] * 16, chi[d] * 16) for d in range(3)], indexing="ij"), -1).reshape(-1, 3)
    p = idx.astype(np.float32) + f(0.5)
    len2 = (vec[0] * vec[0] + vec[1] * vec[1]) + vec[2] * vec[2]
    vol = vec / len2
    sp = p - start
    t = np.clip((sp[:, 0] * vol[0] + sp[:, 1] * vol[1]) + sp[:, 2] * vol[2], f(0.0), f(1.0)).astype(np.float32)
    d = p - (start + vec * t[:, None])
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    inside = d2 <= radius * radius
    cc = info["chunk_counts"]
    cidx = ((idx[:, 0] // 16) * cc[1] + idx[:, 1] // 16) * cc[2] + idx[:, 2] // 16
    not_void = ch_before["kind"][cidx] != 0
    assert st["touched_voxels"] == int((inside & not_void).sum())
    assert st["touched_voxels"] > 0 and st["emptied_voxels"] > 0
    _check_all(oracle, obj)
    # a capsule with a zero-length segment visits the voxels of the sphere, boundary included
    _, obj2 = _gen(oracle, H.sphere_graph(23.3))
    st2 = obj2.absorb_capsule(f([30.2, 24.9, 11.3]), f([0, 0, 0]), 7.5, 9.5)
    _, obj3 = _gen(oracle, H.sphere_graph(23.3))
    st3 = obj3.absorb_sphere(f([30.2, 24.9, 11.3]), 7.5, 9.5)
    assert st2["touched_voxels"] >= st3["touched_voxels"] > 0
    _check_all(oracle, obj2)
    # a capsule that misses the object leaves it untouched
    _, obj4 = _gen(oracle, H.sphere_graph(23.3))
    v0 = obj4.voxels().copy()
    st4 = obj4.absorb_capsule(f([-40, -40, -40]), f([5, 0, 0]), 3.0, 5.0)
    assert st4["touched_voxels"] == 0 and np.array_equal(v0, obj4.voxels()) and len(obj4.dirty()) == 0


def test_repeated_absorption_keeps_invariants_and_remesh_is_consistent(oracle):
    _, obj = _gen(oracle, H.sphere_graph(23.3))
    for step in range(4):
        c = np.float32([8.0 + 5.0 * step, 20.0, 25.0])
        obj.absorb_sphere(c, 6.0, 8.0)
        ch, vx, cc = _check_all(oracle, obj)
    full = obj.mesh()
    # every dirty chunk re-meshed on its own equals its slice of a full re-mesh
    by_chunk = {tuple(s["chunk_indices"]): (s, r) for s, r in zip(full.submeshes, full.vertex_ranges)}
    for lin in obj.dirty():
        idx3 = (lin // (cc[1] * cc[2]), (lin // cc[2]) % cc[1], lin % cc[2])
        cm = obj.mesh_chunk(*[int(x) for x in idx3])
        if cm is None:
            assert tuple(idx3) not in by_chunk
            continue
        s, r = by_chunk[tuple(idx3)]
        assert np.array_equal(full.positions[r[0]: r[1]].view(np.uint32), cm["positions"].view(np.uint32))
        sl = slice(s["index_offset"], s["index_offset"] + s["index_count"])
        assert np.array_equal(full.indices[sl] - r[0], cm["indices"].astype(np.uint32))
