"""The reference's structural unit tests of `VoxelObject` (object.rs:3563-4071) restated on the oracle: its fake
generators `OffsetBoxVoxelGenerator` / `ManualVoxelGenerator` are given as data (helpers.offset_box_chunks /
manual_chunks) and run through `Object.from_generated_chunks` = generate_without_derived_state (+ derived state).
Known answers: chunk counts, stored voxel counts, occupied ranges, adjacency flags of hand-made grids. CPU only."""
import numpy as np
import pytest

import helpers as H
import invariants as INV

X_DN, Y_DN, Z_DN, X_UP, Y_UP, Z_UP = 1 << 2, 1 << 3, 1 << 4, 1 << 5, 1 << 6, 1 << 7  # VoxelFlags (lib.rs:75-101)
FULL = 0xFC


def _obj(oracle, fixture, derive):
    vox, sp, grid = fixture
    return oracle.Object.from_generated_chunks(vox, sp, grid, 0.25, derive=derive)


def _flags_at(obj, i, j, k):
    cc = obj.info()["chunk_counts"]
    _, fl, _, _, empty = INV.dense_fields(obj.chunks(), obj.voxels(), cc)
    assert not empty[i, j, k]
    return int(fl[i, j, k])


def _occupied(obj):
    cc = obj.info()["chunk_counts"]
    return ~INV.dense_fields(obj.chunks(), obj.voxels(), cc)[4]


def test_empty_grids_and_empty_voxels_give_empty_objects(oracle):
    # should_yield_empty_object_when_generating_object_with_empty_grid / _of_empty_voxels (object.rs:3564-3590)
    o = _obj(oracle, H.offset_box_chunks([0, 0, 0]), False)
    assert o.info()["chunk_counts"] == (0, 0, 0) and len(o.chunks()) == 0
    for fx in (H.offset_box_chunks([1, 1, 1], voxel=H.OUTSIDE), H.offset_box_chunks([2, 3, 4], voxel=H.OUTSIDE)):
        o = _obj(oracle, fx, False)
        assert (o.chunks()["kind"] == 0).all()  # contains_only_empty_voxels: every chunk Void


def test_single_voxel_and_uniform_chunks(oracle):
    # should_generate_object_with_single_voxel / _single_uniform_chunk / _single_offset_uniform_chunk (object.rs:3593-3637)
    o = _obj(oracle, H.offset_box_chunks([1, 1, 1]), False)
    assert o.info()["chunk_counts"] == (1, 1, 1)
    assert np.array_equal(o.info()["occupied_voxel_ranges"], [[0, 16]] * 3)  # chunk granularity before the shrink
    assert len(o.voxels()) == 4096 and o.chunks()["kind"][0] == 2
    o = _obj(oracle, H.offset_box_chunks([16, 16, 16]), False)
    assert o.info()["chunk_counts"] == (1, 1, 1) and np.array_equal(o.info()["occupied_voxel_ranges"], [[0, 16]] * 3)
    assert len(o.voxels()) == 0 and o.chunks()["kind"][0] == 1  # one uniform voxel stored
    assert o.chunks()["uniform_flags"][0] == FULL and o.chunks()["uniform_sd"][0] == -128
    o = _obj(oracle, H.offset_box_chunks([16, 16, 16], [16, 16, 16]), False)
    assert o.info()["chunk_counts"] == (2, 2, 2) and np.array_equal(o.info()["occupied_voxel_ranges"], [[16, 32]] * 3)
    assert len(o.voxels()) == 0 and (o.chunks()["kind"] == 1).sum() == 1 and o.chunks()["kind"][7] == 1


CELLS_A = [[[1, 1, 0], [1, 0, 1], [0, 1, 0]], [[0, 1, 1], [1, 0, 0], [1, 0, 1]], [[1, 1, 0], [1, 1, 1], [0, 0, 0]]]


@pytest.mark.parametrize("offset", [(0, 0, 0), (14, 14, 14)])
def test_voxels_of_small_grids_are_where_the_generator_put_them(oracle, offset):
    # should_get_correct_voxels_in_small_grid / _small_offset_grid (object.rs:3640-3685)
    o = _obj(oracle, H.manual_chunks(CELLS_A, offset), False)
    occ = _occupied(o)
    sl = tuple(slice(a, a + 3) for a in offset)
    assert np.array_equal(occ[sl], np.asarray(CELLS_A) != 0)
    assert occ.sum() == np.count_nonzero(CELLS_A)


def test_internal_adjacency_in_chunk(oracle):
    # should_compute_correct_internal_adjacency_in_chunk (object.rs:3688-3725)
    cells = [[[0, 0, 0], [0, 1, 0], [0, 0, 0]], [[0, 1, 0], [1, 1, 1], [0, 1, 0]], [[0, 0, 0], [0, 1, 0], [0, 0, 0]]]
    o = _obj(oracle, H.manual_chunks(cells), True)
    assert _flags_at(o, 1, 1, 1) == FULL
    assert _flags_at(o, 0, 1, 1) == X_UP and _flags_at(o, 2, 1, 1) == X_DN
    assert _flags_at(o, 1, 0, 1) == Y_UP and _flags_at(o, 1, 2, 1) == Y_DN
    assert _flags_at(o, 1, 1, 0) == Z_UP and _flags_at(o, 1, 1, 2) == Z_DN


def test_internal_adjacency_in_chunk_corners(oracle):
    # should_compute_correct_internal_adjacency_in_lower_chunk_corner / _upper_chunk_corner (object.rs:3728-3802)
    cells = [[[1, 1, 0], [1, 0, 0], [0, 0, 0]], [[1, 0, 0], [0, 0, 0], [0, 0, 0]], [[0, 0, 0], [0, 0, 0], [0, 0, 0]]]
    o = _obj(oracle, H.manual_chunks(cells), True)
    assert _flags_at(o, 0, 0, 0) == X_UP | Y_UP | Z_UP
    assert _flags_at(o, 0, 0, 1) == Z_DN and _flags_at(o, 0, 1, 0) == Y_DN and _flags_at(o, 1, 0, 0) == X_DN
    cells = [[[0, 0, 0], [0, 0, 0], [0, 0, 0]], [[0, 0, 0], [0, 0, 0], [0, 0, 1]], [[0, 0, 0], [0, 0, 1], [0, 1, 1]]]
    o = _obj(oracle, H.manual_chunks(cells, (13, 13, 13)), True)
    assert _flags_at(o, 15, 15, 15) == X_DN | Y_DN | Z_DN
    assert _flags_at(o, 15, 15, 14) == Z_UP and _flags_at(o, 15, 14, 15) == Y_UP and _flags_at(o, 14, 15, 15) == X_UP


BOXES = [[1, 1, 1], [16, 16, 16], [17, 16, 16], [16, 17, 16], [16, 16, 17], [17, 1, 1], [1, 17, 1], [1, 1, 17]]


@pytest.mark.parametrize("shape", BOXES)
def test_adjacencies_and_obscuredness_of_boxes(oracle, shape):
    # should_compute_correct_adjacencies_for_single_voxel / _single_chunk / _barely_two_chunks /
    # _with_column_taking_barely_two_chunks (object.rs:3805-3860)
    o = _obj(oracle, H.offset_box_chunks(shape), True)
    cc = o.info()["chunk_counts"]
    INV.validate_adjacencies(o.chunks(), o.voxels(), cc)
    INV.validate_chunk_obscuredness(o.chunks(), cc)
    INV.validate_occupied_voxel_ranges(o.chunks(), o.voxels(), cc, o.info()["occupied_voxel_ranges"])


def test_occupied_voxel_ranges_shrink_to_the_voxels(oracle):
    # should_shrink_occupied_voxel_ranges_correctly_* (object.rs:3996-4045); compute_aabb = these ranges x voxel extent
    # (should_compute_correct_aabb_*, object.rs:3941-3993, 4048-4070)
    cases = [(H.offset_box_chunks([1, 1, 1]), [[0, 1]] * 3), (H.offset_box_chunks([9, 9, 9]), [[0, 9]] * 3),
             (H.offset_box_chunks([1, 1, 1], [5, 5, 5]), [[5, 6]] * 3),
             (H.offset_box_chunks([32, 48, 64]), [[0, 32], [0, 48], [0, 64]]),
             (H.offset_box_chunks([16, 16, 16], [16, 16, 16]), [[16, 32]] * 3)]
    cells = np.zeros((20, 20, 20), np.uint8)
    cells[2, 2, 5] = cells[18, 17, 19] = cells[3, 3, 6] = cells[17, 16, 18] = 1
    cases.append((H.manual_chunks(cells), [[2, 19], [2, 18], [5, 20]]))
    for fx, want in cases:
        o = _obj(oracle, fx, True)
        assert np.array_equal(o.info()["occupied_voxel_ranges"], want), (o.info()["occupied_voxel_ranges"], want)


def test_chunk_flag_bits(oracle):
    # should_mark_correct_{lower,upper}_face_as_{obscured,unobscured}_for_chunk_flags (object.rs:3863-3938): bit d =
    # IS_OBSCURED_{X,Y,Z}_DN, bit 3 + d = .._UP (object.rs:158-182). A box of 3 x 3 x 3 full chunks: the centre chunk
    # stays Uniform (no flags are kept for those), a face-centre chunk was converted to NonUniform and is obscured from
    # all sides but its outward one.
    o = _obj(oracle, H.offset_box_chunks([48, 48, 48]), True)
    ch = o.chunks().reshape(3, 3, 3)
    assert ch["kind"][1, 1, 1] == 1 and (np.delete(ch["kind"].reshape(-1), 13) == 2).all()
    for d in range(3):
        lo, hi = [1, 1, 1], [1, 1, 1]
        lo[d], hi[d] = 0, 2
        assert ch["flags"][tuple(lo)] & 0x3F == 0x3F & ~(1 << d), (d, ch["flags"][tuple(lo)])
        assert ch["flags"][tuple(hi)] & 0x3F == 0x3F & ~(1 << (3 + d)), (d, ch["flags"][tuple(hi)])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_voxel_grids_satisfy_the_reference_invariants(oracle, seed):
    # the reference's fuzz targets assert these on arbitrary generated objects (object.rs:3371-3377)
    vox, sp, grid = H.random_voxel_chunks((40, 37, 50), seed, blobs=seed != 3)
    o = oracle.Object.from_generated_chunks(vox, sp, grid, 1.0)
    cc = o.info()["chunk_counts"]
    INV.validate_adjacencies(o.chunks(), o.voxels(), cc)
    INV.validate_chunk_obscuredness(o.chunks(), cc)
    INV.validate_occupied_voxel_ranges(o.chunks(), o.voxels(), cc, o.info()["occupied_voxel_ranges"])
    kinds = o.chunks()["kind"]
    assert (kinds == 0).any() and (kinds == 2).any()
