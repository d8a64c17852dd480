"""Arguments the library must refuse instead of trusting (GPU box: they go through a context)."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import _lib as L
from impact_b200.voxel import SDFGenerator, SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

pytestmark = pytest.mark.gpu


def test_uploaded_programs_are_walked_before_they_run(ctx):
    # ivx_program_upload takes an already compiled post-order program: an operator without its operands, a list that
    # leaves more than one value, or an understated stack depth must not reach the kernels' operand stacks
    g = H.csg_zoo_graph()
    gen = ctx.build_generator(g)
    nodes = gen.nodes()
    lo, hi = gen.domain
    ok = SDFGenerator.from_processed_nodes(ctx, nodes, gen.stack_depth, lo, hi)
    assert ok.node_count == len(nodes)
    # an understated depth is corrected, not believed
    shallow = SDFGenerator.from_processed_nodes(ctx, nodes, 1, lo, hi)
    assert shallow.stack_depth >= gen.stack_depth or shallow.stack_depth >= 2
    obj_a = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, H.SAME0))
    obj_b = VoxelObject.generate(SDFVoxelGenerator(1.0, shallow, H.SAME0))
    ca, va = obj_a.download()
    cb, vb = obj_b.download()
    assert np.array_equal(ca, cb) and np.array_equal(va, vb)
    combine = int(np.flatnonzero(nodes["kind"] >= 7)[0])
    for bad, what in ((nodes[combine:], "operands"), (nodes[:-1], "instead of one"), (np.concatenate([nodes, nodes[:1]]), "instead of one")):
        with pytest.raises(L.IvxError, match=what):
            SDFGenerator.from_processed_nodes(ctx, np.ascontiguousarray(bad), gen.stack_depth, lo, hi)
    wrong_kind = nodes.copy()
    wrong_kind["kind"][0] = 11
    with pytest.raises(L.IvxError, match="Invalid SDF node kind"):
        SDFGenerator.from_processed_nodes(ctx, wrong_kind, gen.stack_depth, lo, hi)


def test_a_mesh_handle_from_before_a_remesh_is_refused(ctx):
    g = H.sphere_graph(20.0)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(g), H.SAME0))
    first = VoxelObjectMesh.create(obj)
    first.download()
    obj.absorb_sphere(np.float32([8.0, 20.0, 20.0]), 6.0, 8.0)
    patch = VoxelObjectMesh.sync_with_voxel_object(obj)  # replaces the object's mesh by the patch
    assert patch.n_vertices != first.n_vertices
    patch.download()
    with pytest.raises(L.IvxError, match="re-created or patched"):
        first.download()
