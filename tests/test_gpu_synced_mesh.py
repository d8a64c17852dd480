"""`ivx_object_mesh_sync` = `VoxelObjectMesh::sync_with_voxel_object` with its ChunkSubmeshManager / RangeAllocator
(mesh.rs:360-456, 703-848): after every absorption of a sequence the device mesh must equal the oracle's synced mesh —
buffer lengths, the submesh table and the vertex ranges row for row (placement into freed ranges, appends, swap-removes),
the list of updated ranges, and every live vertex / index / index material bit for bit."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import workloads as W
from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

pytestmark = pytest.mark.gpu


def _compare(gm: dict, mesh, om, step):
    assert (mesh.n_vertices, mesh.n_indices, mesh.n_submeshes) == (om.n_vertices, om.n_indices, om.n_submeshes), f"step {step}"
    for f in ("chunk_indices", "index_offset", "index_count", "obscured"):
        assert np.array_equal(gm["submeshes"][f], om.submeshes[f]), f"step {step}: submesh {f}"
    assert np.array_equal(gm["vertex_ranges"], om.vertex_ranges), f"step {step}: vertex ranges"
    for s, (v0, v1) in zip(om.submeshes, om.vertex_ranges):
        i0, i1 = int(s["index_offset"]), int(s["index_offset"]) + int(s["index_count"])
        assert H.f32_bits_equal(gm["positions"][v0:v1], om.positions[v0:v1]).all(), f"step {step}: positions"
        assert H.f32_bits_equal(gm["normals"][v0:v1], om.normals[v0:v1]).all(), f"step {step}: normals"
        assert np.array_equal(gm["indices"][i0:i1], om.indices[i0:i1]), f"step {step}: indices"
        assert np.array_equal(gm["index_materials"][i0:i1], om.index_materials[i0:i1]), f"step {step}: index materials"


@pytest.mark.parametrize("name", ["asteroid_like", "stand_in"])
def test_synced_mesh_follows_the_oracle_through_an_absorption_sequence(ctx, oracle, name):
    graph = H.asteroid_like_graph(24, 40.0) if name == "asteroid_like" else W.asteroid_stand_in(0.5)
    gen = ctx.build_generator(graph)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, H.GRADIENT4))
    oobj = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(graph.nodes(), graph.root_node_id), 1.0, H.GRADIENT4), 4)
    VoxelObjectMesh.create(obj)
    osm = oracle.SyncedMesh(oobj, 4)
    oobj.clear_dirty()
    mesh = VoxelObjectMesh.sync(obj)  # nothing invalidated yet: the mesh as created
    _compare(mesh.download(), mesh, osm, -1)
    rng = np.random.default_rng(11)
    shape = np.array(obj.info()["grid_shape"], np.float64)
    removed_any = False
    for step in range(12):
        c = (shape * rng.uniform(0.15, 0.85, 3)).astype(np.float32)
        r = float(rng.uniform(4, 14))
        obj.absorb_sphere(c, r, r + 2.0)
        oobj.absorb_sphere(c, r, r + 2.0)
        dirty = np.sort(oobj.dirty())
        assert np.array_equal(np.sort(obj.invalidated_mesh_chunk_indices()), dirty)
        osm.sync(oobj, dirty)
        oobj.clear_dirty()
        mesh = VoxelObjectMesh.sync(obj)
        assert len(obj.invalidated_mesh_chunk_indices()) == 0
        _compare(mesh.download(), mesh, osm, step)
        upd, removed = mesh.modifications()
        oupd, oremoved = osm.modifications()
        assert np.array_equal(upd, oupd) and removed == oremoved, f"step {step}: modifications"
        removed_any |= removed
        if step % 3 == 2:  # the renderer catches up every third step
            mesh.report_synchronized()
            osm.report_synchronized()
    # the live chunk meshes are those of a fresh mesh of the final object
    fresh = oobj.mesh(4)
    import test_oracle_synced_mesh as T
    assert T.live_chunk_meshes(osm) == T.live_chunk_meshes(fresh)


def test_mesh_sync_needs_the_full_mesh(ctx):
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(H.sphere_graph(20.0)), H.SAME0))
    with pytest.raises(Exception, match="ivx_object_mesh first"):
        VoxelObjectMesh.sync(obj)
    VoxelObjectMesh.create(obj)
    obj.absorb_sphere(np.float32([20, 20, 38]), 5.0, 7.0)
    VoxelObjectMesh.sync_with_voxel_object(obj)  # the patch call replaces the object's mesh
    with pytest.raises(Exception, match="ivx_object_mesh first"):
        VoxelObjectMesh.sync(obj)
