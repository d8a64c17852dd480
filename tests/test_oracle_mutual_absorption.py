"""Oracle checks for mutual absorption (interaction/absorption.rs:891-1080): what the reference's benchmark and engine
rely on — both objects lose (about) the intersection volume, the derived state stays valid, the inertial updaters stay
within `validate_for_object`'s tolerance of a from-scratch integration. The reference has no unit test of its own for
this function. CPU only."""
import numpy as np

import helpers as H
import invariants as INV


def _sphere(oracle, r, extent=1.0, types=H.SAME0):
    g = H.sphere_graph(r)
    return oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), extent, types), 4)


def _dict_info(o, extent):
    i = o.info()
    i["voxel_extent"] = extent
    return i


def _valid(o):
    cc = o.info()["chunk_counts"]
    INV.validate_adjacencies(o.chunks(), o.voxels(), cc)
    INV.validate_chunk_obscuredness(o.chunks(), cc)


def _rel_close(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all((np.abs(a - b) <= tol) | (np.abs(a - b) <= tol * np.maximum(np.abs(a), np.abs(b))))


def test_two_overlapping_spheres_lose_the_lens(oracle):
    a, b = _sphere(oracle, 20.0), _sphere(oracle, 12.0)
    q = np.float32([0, 0, 0, 1])
    t = np.float32([21.0 + 18.0 - 13.0, 21.0 - 13.0, 21.0 - 13.0])  # B's centre 18 voxels from A's along x
    ma, mb = a.inertial_moments([1.0]).copy(), b.inertial_moments([1.0]).copy()
    va, vb = float(ma[0]), float(mb[0])
    ra, rb = H.intersection_voxel_ranges(_dict_info(a, 1.0), _dict_info(b, 1.0), q, t)
    sa, sb = oracle.absorb_mutually(a, b, q, t, 0.0, ra, rb, [1.0], ma, mb)
    # the lens of two balls (radii 20 and 12, centres 18 apart)
    R, r, d = 20.0, 12.0, 18.0
    lens = np.pi * (R + r - d) ** 2 * (d * d + 2 * d * r - 3 * r * r + 2 * d * R + 6 * r * R - 3 * R * R) / (12 * d)
    assert abs(sa["emptied_voxels"] - lens) < 0.12 * lens and abs(sb["emptied_voxels"] - lens) < 0.12 * lens
    assert ma[0] == va - sa["emptied_voxels"] and mb[0] == vb - sb["emptied_voxels"]  # unit densities: exact
    assert _rel_close(ma, a.inertial_moments([1.0]), 1e-3) and _rel_close(mb, b.inertial_moments([1.0]), 1e-3)
    _valid(a)
    _valid(b)
    # what is left of the two no longer overlaps: a second pass removes (next to) nothing
    sa2, sb2 = oracle.absorb_mutually(a, b, q, t, 0.0, ra, rb)
    assert sa2["emptied_voxels"] + sb2["emptied_voxels"] < 0.05 * lens
    # objects that miss each other are left alone
    c = _sphere(oracle, 12.0)
    before = c.voxels().copy()
    far = np.float32([200.0, 0.0, 0.0])
    assert H.intersection_voxel_ranges(_dict_info(a, 1.0), _dict_info(c, 1.0), q, far) is None
    empty = np.zeros((3, 2), np.uint32)
    oracle.absorb_mutually(a, c, q, far, 0.0, empty, empty)
    assert np.array_equal(c.voxels(), before)


def test_rotated_objects_with_different_voxel_extents_and_smoothness(oracle):
    g = H.asteroid_like_graph(12, 30.0)
    dens = [1.0, 2.7, 0.3, 5.5]
    a = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 0.5, H.GRADIENT4), 4)
    b = _sphere(oracle, 16.0, 0.25, H.GRADIENT4)
    ia, ib = _dict_info(a, 0.5), _dict_info(b, 0.25)
    q = H.quat_from_axis_angle([1.0, 0.4, -0.3], 0.9)
    ca = 0.5 * 0.5 * np.float64(ia["chunk_counts"]) * 16  # world-ish centre of A's grid
    t = (ca + np.array([12.0, 2.0, -1.0]) - H._rotate(q, np.full(3, 0.25 * 17.0))).astype(np.float32)
    ranges = H.intersection_voxel_ranges(ia, ib, q, t)
    assert ranges is not None
    ma, mb = a.inertial_moments(dens).copy(), b.inertial_moments(dens).copy()
    sa, sb = oracle.absorb_mutually(a, b, q, t, 1.5, ranges[0], ranges[1], dens, ma, mb)
    assert sa["emptied_voxels"] > 100 and sb["emptied_voxels"] > 100
    assert _rel_close(ma, a.inertial_moments(dens), 1e-3) and _rel_close(mb, b.inertial_moments(dens), 1e-3)
    _valid(a)
    _valid(b)
    assert len(a.dirty()) > 0 and len(b.dirty()) > 0
