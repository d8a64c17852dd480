"""GPU parity for `ivx_object_from_generated_chunks` (VoxelObject::generate for a host-side ChunkedVoxelGenerator,
object.rs:239-263, 361-404, 1890-1964): the reference's own fake generators (object.rs:3387-3561) and random voxel
grids go through the CUDA classification + derived-state kernels and must equal the oracle chunk for chunk, voxel for
voxel; the resulting objects are then meshed, absorbed into, split and weighed like generated ones. Random grids reach
states smooth SDF objects never produce (salt-and-pepper emptiness, many materials per cube, Uniform chunks beside
Void ones)."""
import numpy as np
import pytest

import helpers as H
import invariants as INV
from impact_b200.voxel import VoxelObject, VoxelObjectMesh

pytestmark = pytest.mark.gpu

CELLS = [[[1, 1, 0], [1, 0, 1], [0, 1, 0]], [[0, 1, 1], [1, 0, 0], [1, 0, 1]], [[1, 1, 0], [1, 1, 1], [0, 0, 0]]]
SPARSE = np.zeros((20, 20, 20), np.uint8)
SPARSE[2, 2, 5] = SPARSE[18, 17, 19] = SPARSE[3, 3, 6] = SPARSE[17, 16, 18] = 1
FIXTURES = {
    "single_voxel": lambda: H.offset_box_chunks([1, 1, 1]),
    "single_empty": lambda: H.offset_box_chunks([1, 1, 1], voxel=H.OUTSIDE),
    "empty_2x3x4": lambda: H.offset_box_chunks([2, 3, 4], voxel=H.OUTSIDE),
    "uniform_chunk": lambda: H.offset_box_chunks([16, 16, 16]),
    "offset_uniform_chunk": lambda: H.offset_box_chunks([16, 16, 16], [16, 16, 16]),
    "barely_two_x": lambda: H.offset_box_chunks([17, 16, 16]),
    "barely_two_y": lambda: H.offset_box_chunks([16, 17, 16]),
    "barely_two_z": lambda: H.offset_box_chunks([16, 16, 17]),
    "column_x": lambda: H.offset_box_chunks([17, 1, 1]),
    "column_z": lambda: H.offset_box_chunks([1, 1, 17]),
    "offset_voxel": lambda: H.offset_box_chunks([1, 1, 1], [5, 5, 5]),
    "box_2x3x4_chunks": lambda: H.offset_box_chunks([32, 48, 64]),
    "box_3x3x3_chunks": lambda: H.offset_box_chunks([48, 48, 48]),
    "manual": lambda: H.manual_chunks(CELLS),
    "manual_offset": lambda: H.manual_chunks(CELLS, (14, 14, 14)),
    "manual_sparse": lambda: H.manual_chunks(SPARSE),
}


def _both(ctx, oracle, fixture, extent=0.25):
    vox, sp, grid = fixture
    return (VoxelObject.from_generated_chunks(ctx, extent, grid, vox, sp),
            oracle.Object.from_generated_chunks(vox, sp, grid, extent))


def _same(g, c, mesh=True):
    gi, ci = g.info(), c.info()
    assert tuple(gi["chunk_counts"]) == tuple(ci["chunk_counts"])
    H.assert_objects_equal(*g.download(), c.chunks(), c.voxels())
    assert np.array_equal(gi["occupied_voxel_ranges"], ci["occupied_voxel_ranges"])
    if mesh:
        H.assert_meshes_equal(VoxelObjectMesh.create(g).download(), c.mesh(2))


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_reference_fixtures_through_the_cuda_path(ctx, oracle, name):
    g, c = _both(ctx, oracle, FIXTURES[name]())
    _same(g, c)
    ch, vx = g.download()
    cc = g.info()["chunk_counts"]
    INV.validate_adjacencies(ch, vx, cc)
    INV.validate_chunk_obscuredness(ch, cc)
    INV.validate_occupied_voxel_ranges(ch, vx, cc, g.info()["occupied_voxel_ranges"])


def test_empty_grid_and_invalid_voxels(ctx):
    vox, sp, grid = H.offset_box_chunks([0, 0, 0])
    g = VoxelObject.from_generated_chunks(ctx, 0.25, grid, vox, sp)
    assert g.info()["chunk_counts"] == (0, 0, 0) and VoxelObjectMesh.create(g).n_vertices == 0
    vox, sp, grid = H.offset_box_chunks([5, 5, 5])
    vox[0]["flags"][0] = 1  # EMPTY flag on a voxel with a negative distance: not a `Voxel` any constructor makes
    with pytest.raises(Exception, match="EMPTY"):
        VoxelObject.from_generated_chunks(ctx, 0.25, grid, vox, sp)
    with pytest.raises(Exception):
        VoxelObject.from_generated_chunks(ctx, 0.0, grid, vox, sp)


@pytest.mark.parametrize("seed,shape,blobs", [(1, (40, 37, 50), True), (2, (64, 48, 33), True), (3, (40, 37, 50), False),
                                              (4, (96, 80, 72), True)])
def test_random_voxel_grids_are_bit_exact_through_every_stage(ctx, oracle, seed, shape, blobs):
    vox, sp, grid = H.random_voxel_chunks(shape, seed, blobs=blobs)
    g, c = _both(ctx, oracle, (vox, sp, grid), 0.5)
    _same(g, c)
    dens = [1.0, 2.7, 0.3, 5.5]
    assert H.f32_bits_equal(g.inertial_moments(dens), c.inertial_moments(dens)).all()
    f = np.float32
    mid = f(0.5) * f(shape)
    mg, mc = g.inertial_moments(dens).copy(), c.inertial_moments(dens).copy()
    for step, (ctr, r) in enumerate([(mid, 9.0), (mid + f([7.5, -3.0, 2.0]), 6.0), (f([3.0, 3.0, 3.0]), 8.0)]):
        sg = g.absorb_sphere_inertial(ctr, r, r + 2.0, dens, mg)
        sc = c.absorb_sphere_inertial(ctr, r, r + 2.0, dens, mc)
        for k in ("touched_chunks", "touched_voxels", "emptied_voxels", "removed_chunks"):
            assert sg[k] == sc[k], (step, k, sg, sc)
        assert H.f32_bits_equal(mg, mc).all()
        assert np.array_equal(np.sort(g.invalidated_mesh_chunk_indices()), np.sort(c.dirty()))
    g.absorb_capsule(mid - f([20, 2, 1]), f([40, 4, 2]), 3.0, 5.0)
    c.absorb_capsule(mid - f([20, 2, 1]), f([40, 4, 2]), 3.0, 5.0)
    _same(g, c)
    sd = c.split_detection()
    if not sd["overflow"]:
        rg = g.resolve_connected_regions()
        assert rg["n_regions"] == sd["n_regions"] and rg["has_two"] == sd["has_two"]
        for _ in range(6):  # split pieces off while there are any
            ic, ec = c.extract_any_disconnected_region()
            ig, eg = g.extract_any_disconnected_region()
            for k in ("found_two", "extracted", "discarded", "single_chunk"):
                assert bool(ig[k]) == bool(ic[k]), (k, ig, ic)
            if not ic["found_two"]:
                break
            if ic["extracted"]:
                _same(eg, ec)
        _same(g, c)
