"""Contacts between two voxel objects (for_each_mutual_voxel_object_contact, collidable.rs:859-1050): the collision probes
of one object sampled in the other's distance field. The oracle against analytic spheres (CPU); the CUDA path
(ivx_objects_mutual_contacts) against the oracle bit for bit — identity and rotated poses, different voxel extents, deep
penetration into uniform chunks, no intersection (GPU)."""
import numpy as np
import pytest

import helpers as H

f32 = np.float32
DENS = [1.0, 2.7, 0.3, 5.5]


def quat_mul(a, b):
    ax, ay, az, aw = [float(x) for x in a]
    bx, by, bz, bw = [float(x) for x in b]
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def iso_inverse(iso):
    q = np.array([-iso[0], -iso[1], -iso[2], iso[3]], np.float64)
    return np.concatenate([q, -H._rotate(q, np.float64(iso[4:]))])


def iso_mul(a, b):
    return np.concatenate([quat_mul(a[:4], b[:4]), H._rotate(a[:4], np.float64(b[4:])) + np.float64(a[4:])])


def poses(q_a, t_a, q_b, t_b):
    """world_to_a, world_to_b (7 floats each) and transform_from_b_to_a = world_to_a * world_to_b.inverted()"""
    wa = np.concatenate([q_a, t_a]).astype(f32)
    wb = np.concatenate([q_b, t_b]).astype(f32)
    b_to_a = iso_mul(np.float64(wa), iso_inverse(np.float64(wb))).astype(f32)
    return wa, wb, b_to_a


def _cpu_pair(oracle, graph, extent, types=H.SAME0):
    obj = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(graph.nodes(), graph.root_node_id), extent, types), 4)
    return obj, oracle.CollisionProbes(obj, obj.mesh(4)), obj.inertial_moments(DENS).copy()


def _centre(obj):
    """centre of a sphere object in its own frame: the middle of its occupied voxel ranges"""
    occ = np.float64(obj.info()["occupied_voxel_ranges"])
    return 0.5 * (occ[:, 0] + occ[:, 1])


def _info(o, extent):
    i = o.info()
    i["voxel_extent"] = extent
    return i


def test_contacts_of_two_spheres_against_the_analytic_lens(oracle):
    ra, rb = 20.0, 14.0
    a, pa, ma = _cpu_pair(oracle, H.sphere_graph(ra), 1.0)
    b, pb, mb = _cpu_pair(oracle, H.sphere_graph(rb), 1.0)
    ca, cb = _centre(a), _centre(b)
    # world = A's frame; B's centre 28 away from A's along x: the lens is 6 deep
    qi = f32([0, 0, 0, 1])
    wa, wb, b_to_a = poses(qi, f32([0, 0, 0]), qi, (cb - (ca + [28.0, 0, 0])).astype(f32))
    ranges = H.intersection_voxel_ranges(_info(a, 1.0), _info(b, 1.0), b_to_a[:4], b_to_a[4:])
    assert ranges is not None
    ab, ba = oracle.mutual_contacts(a, pa, ma, wa, b, pb, mb, wb, ranges[0], ranges[1])
    assert len(ab) >= 5 and len(ba) >= 3
    centre_b_world = ca + [28.0, 0, 0]
    for recs, own_c, own_r, other_c, other_r, sign in ((ab, ca, ra, centre_b_world, rb, 1.0), (ba, centre_b_world, rb, ca, ra, -1.0)):
        pos = recs["position"].astype(np.float64)
        # the probes are mesh vertices of their own sphere ...
        assert np.all(np.abs(np.linalg.norm(pos - own_c, axis=1) - own_r) < 0.6)
        # ... inside the other one, as deep as the analytic distance says (the field is trilinear in i8 samples)
        # (signed distances are stored down to -2.56 voxels: deeper than that the depth saturates)
        depth = np.minimum(other_r - np.linalg.norm(pos - other_c, axis=1), 2.56)
        assert np.all(depth > -0.35) and np.all(np.abs(recs["depth"] - depth) < 0.35), np.abs(recs["depth"] - depth).max()
        # the normal is the other sphere's outward normal at the point (flipped for B's probes: it always points from B to A ... of the pair)
        radial = (pos - other_c) / np.linalg.norm(pos - other_c, axis=1, keepdims=True)
        cosine = np.sum(recs["normal"].astype(np.float64) * radial, axis=1) * sign
        assert np.all(cosine > 0.9), cosine.min()  # (gradients flatten where the stored distance saturates)
        assert np.all(np.abs(np.linalg.norm(recs["normal"], axis=1) - 1.0) < 1e-5)
    # ids are the voxels of the probing object that hold the probes
    assert np.array_equal(ab["indices"], np.floor(ab["position"]).astype(np.uint32))
    # far apart: empty ranges, no contacts
    wa2, wb2, b2a = poses(qi, f32([0, 0, 0]), qi, f32([-300, 0, 0]))
    assert H.intersection_voxel_ranges(_info(a, 1.0), _info(b, 1.0), b2a[:4], b2a[4:]) is None


def test_deep_penetration_uses_the_centre_of_mass_direction(oracle):
    # a small sphere pushed far into a big one: its probes reach uniform chunks / the clamped minimum distance of the big
    # one, where the gradient vanishes; the normal is then the direction from the big one's centre of mass
    a, pa, ma = _cpu_pair(oracle, H.sphere_graph(6.0), 1.0)
    b, pb, mb = _cpu_pair(oracle, H.sphere_graph(40.0), 1.0)
    ca, cb = _centre(a), _centre(b)
    qi = f32([0, 0, 0, 1])
    wa, wb, b_to_a = poses(qi, f32([0, 0, 0]), qi, (cb - (ca + [9.0, 4.0, -3.0])).astype(f32))
    ranges = H.intersection_voxel_ranges(_info(a, 1.0), _info(b, 1.0), b_to_a[:4], b_to_a[4:])
    ab, ba = oracle.mutual_contacts(a, pa, ma, wa, b, pb, mb, wb, ranges[0], ranges[1])
    deep = ab[np.abs(ab["depth"] - 2.56) < 1e-6]  # -MIN_F32 * voxel extent
    assert len(deep) > 0 and len(ba) == 0  # none of the big sphere's surface lies inside the small one
    com_b_world = ca + [9.0, 4.0, -3.0]
    radial = deep["position"] - com_b_world
    radial /= np.linalg.norm(radial, axis=1, keepdims=True)
    assert np.all(np.sum(deep["normal"] * radial, axis=1) > 0.999)


# ---- CUDA path ----------------------------------------------------------------------------------------------------------
def _gpu_pair(ctx, oracle, graph, extent, types):
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh
    c = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(graph.nodes(), graph.root_node_id), extent, types), 4)
    g = VoxelObject.generate(SDFVoxelGenerator(extent, ctx.build_generator(graph), types))
    pc = oracle.CollisionProbes(c, c.mesh(4))
    mesh = VoxelObjectMesh.create(g)
    pg = mesh.collision_probes()
    assert H.f32_bits_equal(pg["points"], pc.points).all()
    mom = c.inertial_moments(DENS).copy()
    assert H.f32_bits_equal(g.inertial_moments(DENS), mom).all()
    return g, c, pc, mom, mesh


def _same_contacts(got, want):
    assert len(got) == len(want), (len(got), len(want))
    assert np.array_equal(got["indices"], want["indices"])
    for f in ("position", "normal", "depth"):
        assert H.f32_bits_equal(got[f], want[f]).all(), f


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["spheres", "rotated_extents", "rotated_equal", "deep", "apart"])
def test_mutual_contacts_match_the_oracle(ctx, oracle, case):
    from impact_b200 import voxel as V
    if case in ("spheres", "deep", "apart"):
        ea = eb = 1.0
        ga, ca, pa, ma, _ = _gpu_pair(ctx, oracle, H.sphere_graph(6.0 if case == "deep" else 20.0), ea, H.SAME0)
        gb, cb, pb, mb, _ = _gpu_pair(ctx, oracle, H.sphere_graph(40.0 if case == "deep" else 14.0), eb, H.SAME0)
        q_a = q_b = f32([0, 0, 0, 1])
        off = {"spheres": [28.0, 0.0, 0.0], "deep": [9.0, 4.0, -3.0], "apart": [300.0, 0.0, 0.0]}[case]
    else:
        ea, eb = (0.5, 0.25) if case == "rotated_extents" else (1.0, 1.0)
        ga, ca, pa, ma, _ = _gpu_pair(ctx, oracle, H.asteroid_like_graph(12, 30.0), ea, H.GRADIENT4)
        gb, cb, pb, mb, _ = _gpu_pair(ctx, oracle, H.sphere_union_graph(0.25), eb, H.GRADIENT4)
        q_a = H.quat_from_axis_angle([0.2, -1.0, 0.4], 0.7)
        q_b = H.quat_from_axis_angle([1.0, 0.4, -0.3], 0.9)
        off = [ea * 20.0, ea * 3.0, -ea * 2.0]
    centre_a = 0.5 * ea * np.float64(ca.info()["chunk_counts"]) * 16
    centre_b = 0.5 * eb * np.float64(cb.info()["chunk_counts"]) * 16
    # world frame: A's centre at the origin, B's centre at `off`; world_to_x maps world points into the object's own frame
    t_a = centre_a - H._rotate(q_a, np.zeros(3))
    t_b = centre_b - H._rotate(q_b, np.float64(off))
    wa, wb, b_to_a = poses(q_a, t_a.astype(f32), q_b, t_b.astype(f32))
    ranges = V.intersection_voxel_ranges(ga.info()["occupied_voxel_ranges"], ea, gb.info()["occupied_voxel_ranges"], eb, b_to_a[:4], b_to_a[4:])
    if case == "apart":
        assert ranges is None
        empty = np.zeros((3, 2), np.uint32)
        g_ab, g_ba = V.mutual_contacts(ga, gb, wa, wb, empty, empty, ma, mb)
        assert len(g_ab) == 0 and len(g_ba) == 0
        return
    assert ranges is not None
    w_ab, w_ba = oracle.mutual_contacts(ca, pa, ma, wa, cb, pb, mb, wb, ranges[0], ranges[1])
    g_ab, g_ba = V.mutual_contacts(ga, gb, wa, wb, ranges[0], ranges[1], ma, mb)
    _same_contacts(g_ab, w_ab)
    _same_contacts(g_ba, w_ba)
    assert len(w_ab) + len(w_ba) > 0
    if case == "deep":
        assert np.any(np.abs(w_ab["depth"] - 2.56) < 1e-6)
    # the whole occupied ranges instead of the intersection's: more probes looked at, the same contacts or more
    full = [ca.info()["occupied_voxel_ranges"], cb.info()["occupied_voxel_ranges"]]
    w2 = oracle.mutual_contacts(ca, pa, ma, wa, cb, pb, mb, wb, full[0], full[1])
    g2 = V.mutual_contacts(ga, gb, wa, wb, full[0], full[1], ma, mb)
    _same_contacts(g2[0], w2[0])
    _same_contacts(g2[1], w2[1])
    assert len(w2[0]) >= len(w_ab)


@pytest.mark.gpu
def test_mutual_contacts_need_probes_on_both_objects(ctx):
    from impact_b200 import voxel as V
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject
    g = H.sphere_graph(10.0)
    a = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(g), H.SAME0))
    b = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(g), H.SAME0))
    iso = f32([0, 0, 0, 1, 0, 0, 0])
    full = a.info()["occupied_voxel_ranges"]
    with pytest.raises(Exception, match="collision probes"):
        V.mutual_contacts(a, b, iso, iso, full, full, np.ones(10, f32), np.ones(10, f32))
