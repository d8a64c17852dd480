"""GPU parity for `ivx_objects_absorb_mutually` (apply_mutual_absorption, interaction/absorption.rs:891-1080) against the
oracle: both objects voxel for voxel, chunk tables, invalidated chunks, stats, the two inertial-moment sets bit for bit,
and the meshes afterwards — identity and rotated transforms, equal and different voxel extents (snapshot padding 1 and
2), hard and smooth subtraction, random voxel grids as the objects."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import voxel as V
from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

pytestmark = pytest.mark.gpu
DENS = [1.0, 2.7, 0.3, 5.5]


def _pair(ctx, oracle, graph, extent, types):
    gen_cpu = oracle.Generator(graph.nodes(), graph.root_node_id)
    c = oracle.Object.generate(oracle.VoxelGenerator(gen_cpu, extent, types), 4)
    g = VoxelObject.generate(SDFVoxelGenerator(extent, ctx.build_generator(graph), types))
    return g, c


def _info(o, extent):
    i = o.info()
    i["voxel_extent"] = extent
    return i


def _same(g, c):
    H.assert_objects_equal(*g.download(), c.chunks(), c.voxels())
    assert np.array_equal(g.info()["occupied_voxel_ranges"], c.info()["occupied_voxel_ranges"])
    assert np.array_equal(np.sort(g.invalidated_mesh_chunk_indices()), np.sort(c.dirty()))


def _run(oracle, ga, ca, gb, cb, q, t, k, ranges, dens=DENS):
    ma_c, mb_c = ca.inertial_moments(dens).copy(), cb.inertial_moments(dens).copy()
    ma_g, mb_g = ga.inertial_moments(dens).copy(), gb.inertial_moments(dens).copy()
    assert H.f32_bits_equal(ma_c, ma_g).all() and H.f32_bits_equal(mb_c, mb_g).all()
    sc = oracle.absorb_mutually(ca, cb, q, t, k, ranges[0], ranges[1], dens, ma_c, mb_c)
    sg = V.absorb_mutually(ga, gb, q, t, k, ranges[0], ranges[1], dens, ma_g, mb_g)
    for side in (0, 1):
        for f in ("touched_chunks", "touched_voxels", "emptied_voxels", "removed_chunks"):
            assert sg[side][f] == sc[side][f], (side, f, sg, sc)
    _same(ga, ca)
    _same(gb, cb)
    assert H.f32_bits_equal(ma_g, ma_c).all(), (ma_g, ma_c)
    assert H.f32_bits_equal(mb_g, mb_c).all(), (mb_g, mb_c)
    H.assert_meshes_equal(VoxelObjectMesh.create(ga).download(), ca.mesh(4))
    H.assert_meshes_equal(VoxelObjectMesh.create(gb).download(), cb.mesh(4))
    return sc


def test_two_spheres_identity_rotation(ctx, oracle):
    ga, ca = _pair(ctx, oracle, H.sphere_graph(20.0), 1.0, H.SAME0)
    gb, cb = _pair(ctx, oracle, H.sphere_graph(12.0), 1.0, H.SAME0)
    q, t = np.float32([0, 0, 0, 1]), np.float32([26.0, 8.0, 8.0])
    # the library's own determine_voxel_ranges_encompassing_intersection (host side) on the GPU objects' occupied ranges
    ranges = V.intersection_voxel_ranges(ga.info()["occupied_voxel_ranges"], 1.0, gb.info()["occupied_voxel_ranges"], 1.0, q, t)
    assert ranges is not None
    sc = _run(oracle, ga, ca, gb, cb, q, t, 0.0, ranges, [1.0])
    assert sc[0]["emptied_voxels"] > 2000 and sc[1]["emptied_voxels"] > 2000
    # whole occupied ranges instead of the intersection's: more voxels visited, same rule
    full = [ca.info()["occupied_voxel_ranges"], cb.info()["occupied_voxel_ranges"]]
    _run(oracle, ga, ca, gb, cb, q, np.float32([24.0, 9.5, 7.25]), 0.0, full, [1.0])
    # no overlap: empty ranges, nothing happens
    empty = np.zeros((3, 2), np.uint32)
    sg = V.absorb_mutually(ga, gb, q, np.float32([300.0, 0, 0]), 0.0, empty, empty)
    assert sg[0]["touched_voxels"] == 0 and sg[1]["touched_voxels"] == 0
    with pytest.raises(Exception):
        V.absorb_mutually(ga, ga, q, t, 0.0, empty, empty)


@pytest.mark.parametrize("ea,eb,k", [(0.5, 0.25, 1.5), (0.25, 0.5, 0.0), (1.0, 1.0, 0.7)])
def test_rotated_objects_extents_and_smoothness(ctx, oracle, ea, eb, k):
    ga, ca = _pair(ctx, oracle, H.asteroid_like_graph(12, 30.0), ea, H.GRADIENT4)
    gb, cb = _pair(ctx, oracle, H.sphere_union_graph(0.25), eb, H.GRADIENT4)
    ia, ib = _info(ca, ea), _info(cb, eb)
    q = H.quat_from_axis_angle([1.0, 0.4, -0.3], 0.9)
    centre_a = 0.5 * ea * np.float64(ia["chunk_counts"]) * 16
    centre_b = 0.5 * eb * np.float64(ib["chunk_counts"]) * 16
    t = (centre_a + ea * np.array([22.0, 3.0, -2.0]) - H._rotate(q, centre_b)).astype(np.float32)
    ranges = V.intersection_voxel_ranges(ia["occupied_voxel_ranges"], ea, ib["occupied_voxel_ranges"], eb, q, t)
    assert ranges is not None
    sc = _run(oracle, ga, ca, gb, cb, q, t, k, ranges)
    assert sc[0]["emptied_voxels"] > 50 and sc[1]["emptied_voxels"] > 50
    # again from another side: uniform chunks converted by the first pass, removed chunks, stale flags
    q2 = H.quat_from_axis_angle([-0.2, 1.0, 0.1], 2.1)
    t2 = (centre_a + ea * np.array([-18.0, -6.0, 9.0]) - H._rotate(q2, centre_b)).astype(np.float32)
    ranges = H.intersection_voxel_ranges(_info(ca, ea), _info(cb, eb), q2, t2)
    if ranges is not None:
        _run(oracle, ga, ca, gb, cb, q2, t2, k, ranges)


def test_random_voxel_grids_absorb_each_other(ctx, oracle):
    fa, fb = H.random_voxel_chunks((64, 48, 50), 11), H.random_voxel_chunks((40, 37, 50), 12)
    ga = VoxelObject.from_generated_chunks(ctx, 0.5, fa[2], fa[0], fa[1])
    ca = oracle.Object.from_generated_chunks(fa[0], fa[1], fa[2], 0.5)
    gb = VoxelObject.from_generated_chunks(ctx, 0.5, fb[2], fb[0], fb[1])
    cb = oracle.Object.from_generated_chunks(fb[0], fb[1], fb[2], 0.5)
    q = H.quat_from_axis_angle([0.3, -1.0, 0.5], 1.3)
    t = np.float32([14.0, 6.0, 3.0])
    ranges = H.intersection_voxel_ranges(_info(ca, 0.5), _info(cb, 0.5), q, t)
    assert ranges is not None
    _run(oracle, ga, ca, gb, cb, q, t, 0.5, ranges)
