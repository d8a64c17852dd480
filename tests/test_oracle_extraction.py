"""Oracle (CPU restatement) of the disconnected-region extraction (object/extraction.rs:78-600, 1902-2187), pinned by
the reference's own checks: `should_split_off_disconnected_sphere` (:2603-2623) and the assertions of
`fuzz_test_voxel_object_split_off_disconnected_region` (:2253-2318) — region counts before / after, the validate_*
invariants on both objects — plus conservation of the non-empty voxels (what the inertial-property transfer checks)."""
import numpy as np
import pytest

import helpers as H
import invariants as INV
from test_oracle_split_detection import _object, two_spheres_graph


def debris_graph():
    from impact_b200.graph import SDFGraph

    g = SDFGraph()
    g.multifractal_noise(g.sphere(14.0), 3, 0.12, 2.0, 0.6, 7.0, 1)  # 67 connected regions
    return g


def dense_nonempty(obj, offset=(0, 0, 0)):
    """{(i, j, k): (type, sd)} of the non-empty voxels, in the coordinates of the parent object."""
    info = obj.info()
    cc = info["chunk_counts"]
    ch, vx = obj.chunks(), obj.voxels().reshape(-1, 4096)
    out = {}
    for c in range(len(ch)):
        if ch["kind"][c] == 0:
            continue
        ci, cj, ck = c // (cc[1] * cc[2]), (c // cc[2]) % cc[1], c % cc[2]
        if ch["kind"][c] == 1:
            t, s = int(ch["uniform_type"][c]), int(ch["uniform_sd"][c])
            for v in range(4096):
                out[(ci * 16 + (v >> 8) + offset[0], cj * 16 + ((v >> 4) & 15) + offset[1], ck * 16 + (v & 15) + offset[2])] = (t, s)
            continue
        v4 = vx[ch["data_offset"][c]]
        for v in np.flatnonzero((v4["flags"] & 1) == 0):
            out[(ci * 16 + (v >> 8) + offset[0], cj * 16 + ((v >> 4) & 15) + offset[1], ck * 16 + (v & 15) + offset[2])] = (
                int(v4["type"][v]), int(v4["sd"][v]))
    return out


def check_invariants(obj):
    info = obj.info()
    ch, vx, cc = obj.chunks(), obj.voxels(), info["chunk_counts"]
    INV.validate_adjacencies(ch, vx, cc)
    INV.validate_chunk_obscuredness(ch, cc)
    INV.validate_occupied_voxel_ranges(ch, vx, cc, info["occupied_voxel_ranges"])
    sd = obj.split_detection()
    assert sd["n_regions"] == obj.count_regions_brute_force()  # validate_region_count
    return sd["n_regions"]


def extract_and_check(obj, expect_extracted=True):
    before = dense_nonempty(obj)
    n_before = obj.split_detection()["n_regions"]
    info, ext = obj.extract_any_disconnected_region()
    if n_before < 2:
        assert not info["found_two"] and ext is None
        assert dense_nonempty(obj) == before
        return info, ext
    assert info["found_two"]
    after = dense_nonempty(obj)
    if info["discarded"]:
        assert ext is None and len(before) - len(after) < 8  # NON_EMPTY_VOXEL_THRESHOLD (object.rs:203)
        assert all(before[k] == v for k, v in after.items())
        return info, ext
    assert info["extracted"] == expect_extracted and ext is not None
    moved = dense_nonempty(ext, info["origin_offset_in_parent"])
    # every non-empty voxel is in exactly one of the two objects, unchanged
    assert set(after).isdisjoint(moved) and len(after) + len(moved) == len(before)
    assert all(before[k] == v for k, v in after.items()) and all(before[k] == v for k, v in moved.items())
    assert check_invariants(ext) == 1
    assert check_invariants(obj) == n_before - 1
    if info["single_chunk"]:
        assert ext.info()["chunk_counts"] == (1, 1, 1)
    return info, ext


def test_should_split_off_disconnected_sphere(oracle):
    obj = _object(oracle, two_spheres_graph())  # extraction.rs:2603-2623
    info, ext = extract_and_check(obj)
    assert info["extracted"] and not info["single_chunk"]
    # nothing left to extract afterwards
    info2, ext2 = obj.extract_any_disconnected_region()
    assert not info2["found_two"] and ext2 is None


def test_smaller_region_leaves_and_uniform_chunks_move_whole(oracle):
    obj = _object(oracle, two_spheres_graph(80.0, 40.0, 12.0))
    big_uniform = int((obj.chunks()["kind"] == 1).sum())
    assert big_uniform > 0
    info, ext = extract_and_check(obj)
    assert int((obj.chunks()["kind"] == 1).sum()) == big_uniform  # the r = 12 sphere left, the big one kept its interior
    # and the other way round: a big detached sphere next to a small remainder takes its uniform chunks along
    obj = _object(oracle, two_spheres_graph(80.0, 12.0, 40.0))
    info, ext = extract_and_check(obj)
    assert int((ext.chunks()["kind"] == 1).sum()) > 0 or int((obj.chunks()["kind"] == 1).sum()) > 0


def test_small_fragment_is_repacked_into_a_single_chunk(oracle):
    from impact_b200.graph import SDFGraph

    g = SDFGraph()
    a = g.sphere(20.0)
    # a r = 5 blob whose voxels straddle a chunk corner of the parent grid: 2 x 2 x 2 chunks before re-packing
    b = g.translation(g.sphere(5.0), [33.0, 9.0, 10.0])
    g.union(a, b, 0.5)
    obj = _object(oracle, g)
    info, ext = extract_and_check(obj)
    assert info["extracted"] and info["single_chunk"]
    occ = ext.info()["occupied_voxel_ranges"]
    assert np.all(occ[:, 0] >= 1) or np.any(occ[:, 0] == 0)  # one empty boundary layer where there was room


def test_debris_is_extracted_or_dropped_piece_by_piece(oracle):
    # strong noise sheds detached blobs of all sizes, mixed chunks included; tiny ones are dropped
    obj = _object(oracle, debris_graph())
    n = obj.split_detection()["n_regions"]
    assert n > 20
    outcomes = set()
    for _ in range(n + 2):
        info, ext = extract_and_check(obj)
        if not info["found_two"]:
            break
        outcomes.add("discarded" if info["discarded"] else ("single" if info["single_chunk"] else "multi"))
    assert obj.split_detection()["n_regions"] == 1
    assert "discarded" in outcomes and ("single" in outcomes or "multi" in outcomes), outcomes


def test_absorbed_bridge_then_extraction(oracle):
    from impact_b200.graph import SDFGraph

    g = SDFGraph()
    a = g.sphere(14.0)
    b = g.translation(g.sphere(14.0), [44.0, 0.0, 0.0])
    bridge = g.capsule(30.0, 3.0)
    bridge = g.rotation_from_axis_angle(bridge, [0.0, 0.0, 1.0], float(np.pi / 2))
    bridge = g.translation(bridge, [22.0, 0.0, 0.0])
    g.union(g.union(a, b, 1.0), bridge, 1.0)
    obj = _object(oracle, g, H.GRADIENT4)
    shape = np.array(obj.info()["chunk_counts"]) * 16
    center = np.float32([0.5 * shape[0], 0.5 * shape[1], 0.5 * shape[2]])
    for _ in range(3):
        obj.absorb_sphere(center, 7.0, 9.0)
    obj.clear_dirty()
    info, ext = extract_and_check(obj)
    assert info["extracted"]
    assert len(obj.dirty()) > 0  # the parent's chunks that lost voxels are re-meshed
    # both halves still mesh
    assert ext.mesh(1).indices.size > 0 and obj.mesh(1).indices.size > 0
