"""GPU parity for the disconnected-region extraction (ivx_object_extract_disconnected_region vs the oracle's
restatement of object/extraction.rs): which region leaves, both resulting objects voxel for voxel, chunk tables,
occupied ranges, invalidated chunks and meshes; tiny fragments dropped, small ones re-packed into one chunk."""
import numpy as np
import pytest

import helpers as H
from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh
from test_oracle_extraction import debris_graph
from test_oracle_split_detection import two_spheres_graph

pytestmark = pytest.mark.gpu


def _both(ctx, oracle, graph, types=H.SAME0):
    gen_cpu = oracle.Generator(graph.nodes(), graph.root_node_id)
    obj_cpu = oracle.Object.generate(oracle.VoxelGenerator(gen_cpu, 1.0, types), 4)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    return obj_gpu, obj_cpu


def _same_object(g, c, mesh=True):
    gi, ci = g.info(), c.info()
    assert tuple(gi["chunk_counts"]) == tuple(ci["chunk_counts"])
    H.assert_objects_equal(*g.download(), c.chunks(), c.voxels())
    assert np.array_equal(gi["occupied_voxel_ranges"], ci["occupied_voxel_ranges"])
    if mesh:
        H.assert_meshes_equal(VoxelObjectMesh.create(g).download(), c.mesh(2))


def _extract_both(obj_gpu, obj_cpu):
    ic, ec = obj_cpu.extract_any_disconnected_region()
    ig, eg = obj_gpu.extract_any_disconnected_region()
    for k in ("found_two", "extracted", "discarded", "single_chunk"):
        assert bool(ig[k]) == bool(ic[k]), (k, ig, ic)
    if ic["found_two"]:
        assert ig["region_label"] == ic["region_label"]
    if ic["extracted"]:
        assert tuple(ig["origin_offset_in_parent"]) == tuple(ic["origin_offset_in_parent"])
        _same_object(eg, ec)
    else:
        assert eg is None
    _same_object(obj_gpu, obj_cpu, mesh=False)
    assert np.array_equal(np.sort(obj_gpu.invalidated_mesh_chunk_indices()), np.sort(obj_cpu.dirty()))
    return ic, ec, eg


def test_split_off_sphere_is_bit_exact(ctx, oracle):
    obj_gpu, obj_cpu = _both(ctx, oracle, two_spheres_graph(80.0, 40.0, 12.0))  # uniform interior stays behind
    ic, *_ = _extract_both(obj_gpu, obj_cpu)
    assert ic["extracted"]
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh(2))
    ic, *_ = _extract_both(obj_gpu, obj_cpu)
    assert not ic["found_two"]
    # two big spheres: the one that leaves takes its uniform interior chunks along (first and second found region)
    for args in ((100.0, 40.0, 36.0), (90.0, 36.0, 40.0)):
        obj_gpu, obj_cpu = _both(ctx, oracle, two_spheres_graph(*args), H.SAME0)
        ic, ec, eg = _extract_both(obj_gpu, obj_cpu)
        assert ic["extracted"] and (ec.chunks()["kind"] == 1).any() and (obj_cpu.chunks()["kind"] == 1).any()
    # the extracted object is a full citizen: absorb into it and re-mesh
    if ic["extracted"]:
        c = (np.array(ec.info()["chunk_counts"]) * 8).astype(np.float32)
        assert eg.absorb_sphere(c, 5.0, 7.0) == {**ec.absorb_sphere(c, 5.0, 7.0), "dirty_chunks": len(ec.dirty())}
        _same_object(eg, ec)


def test_small_fragment_is_repacked_into_one_chunk(ctx, oracle):
    from impact_b200.graph import SDFGraph

    g = SDFGraph()
    a = g.sphere(20.0)
    b = g.translation(g.sphere(5.0), [33.0, 9.0, 10.0])
    g.union(a, b, 0.5)
    obj_gpu, obj_cpu = _both(ctx, oracle, g, H.GRADIENT4)
    ic, ec, eg = _extract_both(obj_gpu, obj_cpu)
    assert ic["extracted"] and ic["single_chunk"] and eg.info()["chunk_counts"] == (1, 1, 1)


def test_debris_field_piece_by_piece(ctx, oracle):
    # 67 regions, mixed chunks, dropped crumbs, single-chunk and multi-chunk fragments
    obj_gpu, obj_cpu = _both(ctx, oracle, debris_graph())
    outcomes = set()
    for _ in range(80):
        ic, *_ = _extract_both(obj_gpu, obj_cpu)
        if not ic["found_two"]:
            break
        outcomes.add("discarded" if ic["discarded"] else ("single" if ic["single_chunk"] else "multi"))
    assert not ic["found_two"] and "discarded" in outcomes and len(outcomes) >= 2
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh(2))
    assert obj_gpu.resolve_connected_regions()["n_regions"] == 1


def test_absorb_until_it_splits_then_extract(ctx, oracle):
    from impact_b200.graph import SDFGraph

    g = SDFGraph()
    a = g.sphere(14.0)
    b = g.translation(g.sphere(14.0), [44.0, 0.0, 0.0])
    bridge = g.capsule(30.0, 3.0)
    bridge = g.rotation_from_axis_angle(bridge, [0.0, 0.0, 1.0], float(np.pi / 2))
    bridge = g.translation(bridge, [22.0, 0.0, 0.0])
    g.union(g.union(a, b, 1.0), bridge, 1.0)
    obj_gpu, obj_cpu = _both(ctx, oracle, g, H.GRADIENT4)
    shape = np.array(obj_cpu.info()["chunk_counts"]) * 16
    center = np.float32([0.5 * shape[0], 0.5 * shape[1], 0.5 * shape[2]])
    ic, *_ = _extract_both(obj_gpu, obj_cpu)
    assert not ic["found_two"]
    for _ in range(3):
        obj_cpu.absorb_sphere(center, 7.0, 9.0)
        obj_gpu.absorb_sphere(center, 7.0, 9.0)
    VoxelObjectMesh.sync_with_voxel_object(obj_gpu)
    obj_cpu.clear_dirty()
    ic, ec, eg = _extract_both(obj_gpu, obj_cpu)
    assert ic["extracted"]
    patch = VoxelObjectMesh.sync_with_voxel_object(obj_gpu)
    assert patch.n_submeshes >= 0 and len(obj_gpu.invalidated_mesh_chunk_indices()) == 0
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh(2))
