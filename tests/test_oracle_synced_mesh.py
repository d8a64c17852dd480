"""The oracle's restatement of the mesh that is kept in sync with a modified object
(`VoxelObjectMesh::sync_with_voxel_object` + `ChunkSubmeshManager`, mesh.rs:360-456, 703-848) and of `RangeAllocator`
(impact_containers/src/range_allocator.rs), pinned by the reference's own unit tests of the allocator and by the
invariant the reference relies on: after any sequence of syncs the live ranges of the patched mesh hold exactly the
chunk meshes a fresh `recreate` of the same object produces."""
import numpy as np

import helpers as H


# ---- range_allocator.rs:150-249, test for test ----
def test_allocates_nothing_before_freed(oracle):
    assert oracle.RangeAllocator().allocate_range(1) is None


def test_frees_and_allocates_single_range(oracle):
    a = oracle.RangeAllocator()
    a.free_range(2, 6)
    assert a.allocate_range(4) == (2, 6)
    assert a.allocate_range(1) is None


def test_allocates_range_in_smallest_slot(oracle):
    a = oracle.RangeAllocator()
    a.free_range(2, 6)
    a.free_range(10, 12)
    assert a.allocate_range(2) == (10, 12)
    assert a.allocate_range(4) == (2, 6)


def test_uses_parts_of_larger_slots(oracle):
    a = oracle.RangeAllocator()
    a.free_range(2, 12)
    assert a.allocate_range(4) == (2, 6)
    assert a.allocate_range(4) == (6, 10)
    assert a.allocate_range(4) is None
    assert a.allocate_range(2) == (10, 12)
    assert a.allocate_range(1) is None


def test_does_not_merge_two_disconnected_free_ranges(oracle):
    a = oracle.RangeAllocator()
    a.free_range(2, 5)
    a.free_range(6, 9)
    a.merge_consecutive_ranges()
    assert a.allocate_range(6) is None


def test_merges_consecutive_free_ranges(oracle):
    for ranges, want in (([(2, 6), (6, 8)], (2, 8)), ([(2, 6), (6, 8), (8, 42)], (2, 42)),
                         ([(2, 6), (6, 8), (8, 42), (42, 50)], (2, 50))):
        a = oracle.RangeAllocator()
        for r in ranges:
            a.free_range(*r)
        a.merge_consecutive_ranges()
        assert a.allocate_range(want[1] - want[0]) == want
        assert a.allocate_range(1) is None


# ---- the synced mesh ----
def live_chunk_meshes(m):
    """{chunk indices: (positions, normals, chunk-local indices, index materials, obscuredness)} from the live ranges."""
    out = {}
    for s, (v0, v1) in zip(m.submeshes, m.vertex_ranges):
        i0, i1 = int(s["index_offset"]), int(s["index_offset"]) + int(s["index_count"])
        out[tuple(int(x) for x in s["chunk_indices"])] = (
            m.positions[v0:v1].view(np.uint32).tobytes(), m.normals[v0:v1].view(np.uint32).tobytes(),
            (m.indices[i0:i1] - np.uint32(v0)).tobytes(), m.index_materials[i0:i1].tobytes(), s["obscured"].tobytes())
    return out


def assert_ranges_disjoint(m):
    for ranges in (m.vertex_ranges.astype(np.int64),
                   np.stack([m.submeshes["index_offset"], m.submeshes["index_offset"] + m.submeshes["index_count"]], 1).astype(np.int64)):
        r = ranges[np.argsort(ranges[:, 0])]
        assert np.all(r[1:, 0] >= r[:-1, 1]), "live ranges overlap"


def test_synced_mesh_equals_a_fresh_mesh_after_every_absorption(oracle):
    g = H.asteroid_like_graph(24, 40.0)
    obj = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, H.GRADIENT4), 4)
    sm = oracle.SyncedMesh(obj, 4)
    fresh = obj.mesh(4)
    assert live_chunk_meshes(sm) == live_chunk_meshes(fresh) and sm.n_vertices == fresh.n_vertices
    obj.clear_dirty()
    rng = np.random.default_rng(5)
    shape = np.array(obj.info()["chunk_counts"]) * 16
    grew = reused = False
    for step in range(10):
        c = (shape * rng.uniform(0.2, 0.8, 3)).astype(np.float32)
        obj.absorb_sphere(c, float(rng.uniform(4, 9)), float(rng.uniform(9, 12)))
        dirty = np.sort(obj.dirty())
        before = (sm.n_vertices, sm.n_indices)
        sm.sync(obj, dirty)
        obj.clear_dirty()
        fresh = obj.mesh(4)
        assert live_chunk_meshes(sm) == live_chunk_meshes(fresh), f"step {step}"
        assert_ranges_disjoint(sm)
        upd, removed = sm.modifications()
        assert len(upd) <= len(dirty)
        grew |= (sm.n_vertices, sm.n_indices) != before
        reused |= len(upd) > 0 and bool(np.any(upd[:, 1] <= before[0]))
        sm.report_synchronized()
        assert len(sm.modifications()[0]) == 0
    assert grew and reused  # both placements happened: appended at the end, and into freed ranges
