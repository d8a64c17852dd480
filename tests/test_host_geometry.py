"""Host-side geometry of the mutual absorption path (`ivx_box_intersection_bounds`, `ivx_intersection_voxel_ranges`; no
device needed): the reference's own known-answer tests of `compute_box_intersection_bounds`
(impact_geometry/src/oriented_box.rs:598-830), a randomised containment check in float64, and
`determine_voxel_ranges_encompassing_intersection` (object/intersection.rs:707-745) on voxel objects from the oracle."""
import numpy as np
import pytest

import helpers as H
import invariants as INV
from impact_b200 import voxel as V

ID = [0.0, 0.0, 0.0, 1.0]
A = ([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0])


def _bounds(center, q=ID, half=(1.0, 1.0, 1.0), a=A):
    return V.box_intersection_bounds(a[0], a[1], center, q, half)


def test_reference_known_answers_for_box_intersection_bounds():
    # ..._with_non_intersecting_boxes_returns_none
    assert _bounds([5.0, 0.0, 0.0]) is None
    # ..._with_identical_axis_aligned_boxes_works
    (alo, ahi), (blo, bhi) = _bounds([0.0, 0.0, 0.0])
    assert np.array_equal(alo, [-1, -1, -1]) and np.array_equal(ahi, [1, 1, 1])
    assert np.array_equal(blo, [-1, -1, -1]) and np.array_equal(bhi, [1, 1, 1])
    # ..._with_partial_overlap_works
    (alo, ahi), (blo, bhi) = _bounds([1.0, 0.0, 0.0])
    assert np.array_equal(alo, [0, -1, -1]) and np.array_equal(ahi, [1, 1, 1])
    assert np.array_equal(blo, [-1, -1, -1]) and np.array_equal(bhi, [0, 1, 1])
    # ..._with_one_box_inside_other_works
    (alo, ahi), (blo, bhi) = _bounds([0.0, 0.0, 0.0], half=(0.5, 0.5, 0.5), a=([-2.0] * 3, [2.0] * 3))
    assert np.array_equal(alo, [-0.5] * 3) and np.array_equal(ahi, [0.5] * 3)
    assert np.array_equal(blo, [-0.5] * 3) and np.array_equal(bhi, [0.5] * 3)
    # ..._with_rotated_box_works: 45 degrees about z, both bounds inside their boxes
    (alo, ahi), (blo, bhi) = _bounds([0.0, 0.0, 0.0], q=H.quat_from_axis_angle([0, 0, 1], np.pi / 4))
    assert (alo >= -1).all() and (ahi <= 1).all() and (blo >= -1).all() and (bhi <= 1).all()
    assert ahi[0] - alo[0] > 1.9  # the rotated square reaches both x faces of A
    # ..._with_touching_boxes_works: zero width in x
    (alo, ahi), (blo, bhi) = _bounds([2.0, 0.0, 0.0])
    assert np.array_equal(alo, [1, -1, -1]) and np.array_equal(ahi, [1, 1, 1])
    assert np.array_equal(blo, [-1, -1, -1]) and np.array_equal(bhi, [-1, 1, 1])
    # ..._with_corner_intersection_works
    (alo, ahi), (blo, bhi) = _bounds([1.5, 1.5, 1.5])
    assert np.array_equal(alo, [0.5] * 3) and np.array_equal(ahi, [1.0] * 3)
    assert np.array_equal(blo, [-1.0] * 3) and np.array_equal(bhi, [-0.5] * 3)


@pytest.mark.parametrize("seed", range(8))
def test_bounds_contain_the_overlap_of_random_boxes(seed):
    rng = np.random.default_rng(seed)
    a_lo = rng.uniform(-4, 0, 3)
    a_hi = a_lo + rng.uniform(1, 6, 3)
    q = H.quat_from_axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
    half = rng.uniform(0.5, 3.0, 3)
    centre = rng.uniform(-3, 4, 3)
    res = V.box_intersection_bounds(a_lo, a_hi, centre, q, half)
    # points of A that lie in B (float64), and the same points in B's frame
    pts = rng.uniform(a_lo, a_hi, (20000, 3))
    x, y, z, w = [float(c) for c in q]
    b = np.array([-x, -y, -z])
    v = pts - centre
    in_b = v * (w * w - b @ b) + np.outer(v @ b, b) * 2.0 + np.cross(b, v) * (2.0 * w)
    inside = (np.abs(in_b) <= half).all(axis=1)
    if res is None:
        assert not inside.any()
        return
    (alo, ahi), (blo, bhi) = res
    eps = 1e-4
    assert (alo >= a_lo - eps).all() and (ahi <= a_hi + eps).all()
    assert (blo >= -half - eps).all() and (bhi <= half + eps).all()
    if inside.any():
        assert (pts[inside] >= alo - eps).all() and (pts[inside] <= ahi + eps).all()
        assert (in_b[inside] >= blo - eps).all() and (in_b[inside] <= bhi + eps).all()


def test_voxel_ranges_encompass_the_overlap_of_two_objects(oracle):
    def sphere(r, extent):
        g = H.sphere_graph(r)
        return oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), extent, H.SAME0), 2)

    ea, eb = 0.5, 0.25
    a, b = sphere(20.0, ea), sphere(12.0, eb)
    oa, ob = a.info()["occupied_voxel_ranges"], b.info()["occupied_voxel_ranges"]
    q = H.quat_from_axis_angle([0.2, 1.0, -0.4], 0.8)
    centre_a, centre_b = ea * 21.0 * np.ones(3), eb * 13.0 * np.ones(3)
    t = (centre_a + np.array([ea * 18.0, 0.0, 0.0]) - H._rotate(q, centre_b)).astype(np.float32)
    ranges = V.intersection_voxel_ranges(oa, ea, ob, eb, q, t)
    assert ranges is not None
    ra, rb = ranges
    assert (ra[:, 0] >= oa[:, 0]).all() and (ra[:, 1] <= oa[:, 1]).all() and (ra[:, 0] < ra[:, 1]).all()
    assert (rb[:, 0] >= ob[:, 0]).all() and (rb[:, 1] <= ob[:, 1]).all() and (rb[:, 0] < rb[:, 1]).all()
    assert (ra[0, 1] - ra[0, 0]) < (oa[0, 1] - oa[0, 0])  # really a sub-range along the axis the objects are offset on
    # every voxel the absorption empties lies inside the ranges: run it with the whole occupied ranges and compare
    ea_before = INV.dense_fields(a.chunks(), a.voxels(), a.info()["chunk_counts"])[4]
    eb_before = INV.dense_fields(b.chunks(), b.voxels(), b.info()["chunk_counts"])[4]
    oracle.absorb_mutually(a, b, q, t, 0.0, oa, ob)
    ea_after = INV.dense_fields(a.chunks(), a.voxels(), a.info()["chunk_counts"])[4]
    eb_after = INV.dense_fields(b.chunks(), b.voxels(), b.info()["chunk_counts"])[4]
    for before, after, r in ((ea_before, ea_after, ra), (eb_before, eb_after, rb)):
        emptied = np.argwhere(after & ~before)
        assert len(emptied) > 100
        assert (emptied >= r[:, 0]).all() and (emptied < r[:, 1]).all(), (emptied.min(0), emptied.max(0), r)
    # far apart: None
    assert V.intersection_voxel_ranges(oa, ea, ob, eb, q, t + np.float32([100.0, 0, 0])) is None


def test_voxel_ranges_within_plane_keep_every_voxel_of_the_negative_halfspace():
    # voxel_ranges_within_plane (object/intersection.rs:751-761): axis-aligned planes give the obvious cut ...
    occ = np.array([[3, 40], [0, 25], [7, 33]], np.uint32)
    assert np.array_equal(V.voxel_ranges_within_plane(occ, [1.0, 0.0, 0.0], 10.3), [[3, 11], [0, 25], [7, 33]])
    assert np.array_equal(V.voxel_ranges_within_plane(occ, [0.0, -1.0, 0.0], -10.3), [[3, 40], [10, 25], [7, 33]])
    assert np.array_equal(V.voxel_ranges_within_plane(occ, [0.0, 0.0, 1.0], 100.0), occ)
    r = V.voxel_ranges_within_plane(occ, [0.0, 0.0, 1.0], 2.0)  # the whole box is on the positive side
    assert r[2, 0] >= r[2, 1]
    # ... and for any plane no voxel whose centre is on the negative side is cut away
    rng = np.random.default_rng(4)
    ii, jj, kk = np.meshgrid(*[np.arange(a, b) + 0.5 for a, b in occ], indexing="ij")
    centres = np.stack([ii, jj, kk], -1).reshape(-1, 3)
    for _ in range(20):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        d = float(rng.uniform(-5, 45))
        r = V.voxel_ranges_within_plane(occ, n.astype(np.float32), d).astype(np.float64)
        neg = centres[centres @ n <= d]
        if len(neg):
            assert (neg >= r[:, 0]).all() and (neg <= r[:, 1]).all(), (n, d, r)
        assert (r[:, 0] >= occ[:, 0]).all() and (r[:, 1] <= occ[:, 1]).all()


def test_voxel_ranges_within_plane_equal_the_oracle(oracle):
    # the library's host function against the oracle's restatement of projected_onto_negative_halfspace, same f32 steps
    rng = np.random.default_rng(9)
    for _ in range(300):
        lo = rng.integers(0, 40, 3)
        occ = np.stack([lo, lo + rng.integers(1, 90, 3)], 1).astype(np.uint32)
        n = rng.normal(size=3)
        if rng.random() < 0.3:
            n[rng.integers(0, 3)] = 0.0  # planes parallel to an axis: the tolerance branch
        n = (n / np.linalg.norm(n)).astype(np.float32)
        d = float(np.float32(rng.uniform(-20, 150)))
        assert np.array_equal(V.voxel_ranges_within_plane(occ, n, d), oracle.voxel_ranges_within_plane(occ, n, d)), (occ, n, d)
