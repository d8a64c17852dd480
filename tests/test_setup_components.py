"""The setup components of the reference (impact_voxel/src/setup.rs) as the host mirror builds them: the graphs they add
(CPU, against hand-built graphs) and `setup_voxel_object` end to end on the device against the oracle."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import setup as S
from impact_b200.graph import SDFGraph


def test_shape_components_add_the_reference_s_nodes():
    g = SDFGraph()
    S.VoxelBox(0.25, 10.0, 12.0, 14.0).add(g)
    assert list(g.nodes()["kind"]) == [2] and list(g.nodes()[0]["p"][:3]) == [10.0, 12.0, 14.0]
    g = SDFGraph()
    S.VoxelSphere(0.25, 9.0).add(g)
    assert list(g.nodes()["kind"]) == [0] and g.nodes()[0]["p"][0] == 9.0
    g = SDFGraph()
    S.VoxelCapsule(0.25, 20.0, 4.0).add(g)
    assert list(g.nodes()["kind"]) == [1] and list(g.nodes()[0]["p"][:2]) == [20.0, 4.0]
    # sphere 1, sphere 2, translation of sphere 2, union (setup.rs:515-526)
    g = SDFGraph()
    root = S.VoxelSphereUnion(0.25, 10.0, 6.0, [9.0, 1.0, -2.0], 1.5).add(g)
    n = g.nodes()
    assert list(n["kind"]) == [0, 0, 3, 7] and root == 3 and g.root_node_id == 3
    assert list(n[3]["child"]) == [0, 2] and n[3]["p"][0] == np.float32(1.5) and list(n[2]["p"][:3]) == [9.0, 1.0, -2.0]
    # the modification goes on top and becomes the root
    S.apply_modifications(g, root, S.MultifractalNoiseSDFModification(3, 0.05, 2.0, 0.5, 1.25, 7))
    n = g.nodes()
    assert n[4]["kind"] == 6 and n[4]["child"][0] == 3 and n[4]["octaves"] == 3 and n[4]["seed"] == 7 and g.root_node_id == 4
    S.apply_modifications(g, 4, None)
    assert len(g) == 5
    for bad in (lambda: S.VoxelBox(0.0, 1, 1, 1), lambda: S.VoxelSphere(0.25, -1.0), lambda: S.VoxelCapsule(0.25, 1.0, 0.0)):
        with pytest.raises(AssertionError):
            bad()
    t = S.GradientNoiseVoxelTypes([0, 1, 2], 0.02, 1.0, 5).create_generator()
    assert (t.kind, t.n_types, t.seed) == (1, 3, 5) and S.SameVoxelType(4).create_generator().same_type == 4


@pytest.mark.gpu
def test_setup_voxel_object_gives_object_mesh_and_probes_like_the_oracle(ctx, oracle):
    from impact_b200.voxel import SDFVoxelGenerator
    shape = S.VoxelSphereUnion(1.0, 18.0, 12.0, [16.0, 3.0, -4.0], 2.0)
    noise = S.MultifractalNoiseSDFModification(3, 0.05, 2.0, 0.5, 1.5, 3)
    types = S.GradientNoiseVoxelTypes([0, 1, 2, 3], 0.02, 1.0, 0).create_generator()
    gen = S.create_sdf_generator(ctx, shape, noise)
    obj, mesh, probes = S.setup_voxel_object(SDFVoxelGenerator(shape.voxel_extent, gen, types))
    g = SDFGraph()
    S.apply_modifications(g, shape.add(g), noise)
    ocpu = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, types), 4)
    H.assert_objects_equal(*obj.download(), ocpu.chunks(), ocpu.voxels())
    omesh = ocpu.mesh(4)
    H.assert_meshes_equal(mesh.download(), omesh)
    want = oracle.CollisionProbes(ocpu, omesh)
    assert H.f32_bits_equal(probes["points"], want.points).all() and len(probes["ranges"]) == len(want.ranges)


@pytest.mark.gpu
def test_generated_voxel_object_compiles_its_meta_graph_in_the_library(ctx):
    from impact_b200 import meta as M
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject
    comp = S.GeneratedVoxelObject(M.asteroid_meta_nodes(), 1.0, 0.28, 0)
    graph = comp.build_graph(ctx)
    assert np.array_equal(graph.nodes(), M.MetaCompiler(M.asteroid_meta_nodes(), 0.28, 0, ctx).build().nodes())
    obj = VoxelObject.generate(SDFVoxelGenerator(comp.voxel_extent, ctx.build_generator(graph), S.SameVoxelType(0).create_generator()))
    assert obj.info()["n_non_uniform"] > 20
