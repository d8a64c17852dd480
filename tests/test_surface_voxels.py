"""Surface voxel queries (object/intersection.rs:51-151): the oracle against a brute-force scan of the dense voxel
fields (CPU), and `ivx_object_surface_voxels_*` against the oracle record for record, order included (GPU)."""
import numpy as np
import pytest

import helpers as H
import invariants as INV


def _brute_force(obj, ranges):
    cc = obj.info()["chunk_counts"]
    _, fl, _, _, empty = INV.dense_fields(obj.chunks(), obj.voxels(), cc)
    blocked = np.zeros(fl.shape, np.int32)
    for b in range(2, 8):
        blocked += (fl >> b) & 1
    kinds = obj.chunks()["kind"].reshape(cc)
    non_uniform = np.kron(kinds == 2, np.ones((16, 16, 16), bool))
    surface = ~empty & (blocked < 6) & non_uniform
    sel = np.zeros_like(surface)
    sel[tuple(slice(int(a), int(b)) for a, b in ranges)] = True
    idx = np.argwhere(surface & sel)
    place = np.where(blocked == 5, 0, np.where(blocked == 4, 1, 2))
    return {tuple(i): int(place[tuple(i)]) for i in idx}


@pytest.mark.parametrize("name", ["sphere", "asteroid_like", "random"])
def test_oracle_surface_voxels_equal_a_brute_force_scan(oracle, name):
    if name == "random":
        vox, sp, grid = H.random_voxel_chunks((40, 37, 50), 5)
        o = oracle.Object.from_generated_chunks(vox, sp, grid, 1.0)
    else:
        g = H.sphere_graph(24.0) if name == "sphere" else H.asteroid_like_graph(12, 24.0)
        o = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, H.GRADIENT4), 2)
    occ = o.info()["occupied_voxel_ranges"]
    for ranges in (occ, np.array([[5, 30], [0, 21], [17, 40]], np.uint32), np.array([[3, 3], [0, 9], [0, 9]], np.uint32)):
        got = o.surface_voxels_in_ranges(ranges)
        want = _brute_force(o, ranges)
        assert len(got) == len(want)
        assert {tuple(int(x) for x in r["indices"]): int(r["placement"]) for r in got} == want
        # the closure's call order: chunk by chunk, then voxel by voxel
        key = [(tuple(int(x) >> 4 for x in r["indices"]), tuple(int(x) for x in r["indices"])) for r in got]
        assert key == sorted(key)
    assert len(o.surface_voxels_in_ranges()) > 100
    placements = np.bincount(o.surface_voxels_in_ranges()["placement"], minlength=3)
    assert placements[0] > 0 and placements.sum() == len(o.surface_voxels_in_ranges())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere", "asteroid_like", "random"])
def test_gpu_surface_voxels_equal_the_oracle_in_order(ctx, oracle, name):
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject

    if name == "random":
        vox, sp, grid = H.random_voxel_chunks((64, 48, 50), 6)
        c = oracle.Object.from_generated_chunks(vox, sp, grid, 0.5)
        g = VoxelObject.from_generated_chunks(ctx, 0.5, grid, vox, sp)
    else:
        graph = H.sphere_graph(40.0) if name == "sphere" else H.asteroid_like_graph(16, 36.0)
        c = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(graph.nodes(), graph.root_node_id), 0.5, H.GRADIENT4), 4)
        g = VoxelObject.generate(SDFVoxelGenerator(0.5, ctx.build_generator(graph), H.GRADIENT4))

    def same(got, want):
        assert len(got) == len(want), (len(got), len(want))
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))

    same(g.surface_voxels_in_ranges(), c.surface_voxels_in_ranges())
    shape = np.array(c.info()["chunk_counts"]) * 16
    sub = np.array([[5, shape[0] - 9], [0, 21], [17, shape[2] - 3]], np.uint32)
    same(g.surface_voxels_in_ranges(sub), c.surface_voxels_in_ranges(sub))
    assert len(g.surface_voxels_in_ranges(np.array([[3, 3], [0, 9], [0, 9]], np.uint32))) == 0
    # for_each_surface_voxel_maybe_intersecting_sphere / _capsule: the shape's box clipped to the occupied ranges
    occ = c.info()["occupied_voxel_ranges"].astype(np.float64)
    f = np.float32
    centre, radius = (0.5 * shape + [0.4 * shape[0], -4.0, 3.0]).astype(f), f(11.5)  # at the object's +x side
    lo, hi = np.floor(np.maximum(centre - radius, 0.0)), np.ceil(centre + radius)
    r = np.stack([np.maximum(lo, occ[:, 0]), np.minimum(hi, occ[:, 1])], 1).astype(np.uint32)
    got = g.surface_voxels_touching_sphere(centre, float(radius))
    same(got, c.surface_voxels_in_ranges(r))
    assert len(got) > 0
    a, v = (0.5 * shape - [20.0, 3.0, 1.0]).astype(f), f([40.0, 6.0, 2.0])
    lo = np.floor(np.maximum(np.minimum(a - f(4.0), a + v - f(4.0)), 0.0))
    hi = np.ceil(np.maximum(a + f(4.0), a + v + f(4.0)))
    r = np.stack([np.maximum(lo, occ[:, 0]), np.minimum(hi, occ[:, 1])], 1).astype(np.uint32)
    same(g.surface_voxels_touching_capsule(a, v, 4.0), c.surface_voxels_in_ranges(r))
    # ..._negative_halfspace_of_plane: the occupied box fitted to the halfspace (host geometry of the library)
    from impact_b200 import voxel as V
    n = f([1.0, 0.05, -0.03])
    n = (n / f(np.linalg.norm(n))).astype(f)  # tilted a little off the x axis: the fitted box ends before the object does
    d = float(n @ (0.5 * shape).astype(f)) - 10.0
    same(g.surface_voxels_within_plane(n, d), c.surface_voxels_in_ranges(V.voxel_ranges_within_plane(occ, n, d)))
    assert 0 < len(g.surface_voxels_within_plane(n, d)) < len(c.surface_voxels_in_ranges())
    # after an absorption the exposed voxels change; still the same list
    g.absorb_sphere(centre, 12.0, 14.0)
    c.absorb_sphere(centre, 12.0, 14.0)
    same(g.surface_voxels_in_ranges(), c.surface_voxels_in_ranges())


def _contacts_float64(obj, extent, q, t, centre, radius):
    """for_each_sphere_voxel_object_contact in float64: the surface voxels inside the sphere's box (in normalized voxel
    space, clipped to the occupied ranges — voxels outside it are never visited, even if their own sphere would reach)."""
    c_norm = (H._rotate(q, np.asarray(centre, np.float64)) + np.asarray(t, np.float64)) / extent
    occ = obj.info()["occupied_voxel_ranges"].astype(np.float64)
    lo = np.maximum(np.floor(np.maximum(c_norm - radius / extent, 0.0)), occ[:, 0])
    hi = np.minimum(np.ceil(c_norm + radius / extent), occ[:, 1])
    sv = obj.surface_voxels_in_ranges(np.stack([lo, hi], 1).astype(np.uint32))
    idx = sv["indices"].astype(np.float64)
    c_obj = (idx + 0.5) * extent
    x, y, z, w = [float(c) for c in q]
    b = np.array([-x, -y, -z])
    v = c_obj - np.asarray(t, np.float64)
    vc = v * (w * w - b @ b) + np.outer(v @ b, b) * 2.0 + np.cross(b, v) * (2.0 * w)
    vr = -(sv["sd"].astype(np.float64) * 0.02) * extent
    disp = np.asarray(centre, np.float64) - vc
    dist = np.linalg.norm(disp, axis=1)
    hit = dist <= radius + vr
    margin = np.abs(dist - (radius + vr))
    n = disp / dist[:, None]
    return sv["indices"], hit, margin, vc + vr[:, None] * n, n, np.maximum(0.0, radius + vr - dist)


def _check_contacts(got, ref, tol=2e-4):
    indices, hit, margin, pos, nrm, depth = ref
    sure = margin > 1e-3  # voxels whose spheres graze the query sphere may go either way in f32
    got_set = {tuple(int(x) for x in r) for r in got["indices"]}
    for i in np.flatnonzero(sure):
        assert (tuple(int(x) for x in indices[i]) in got_set) == bool(hit[i]), indices[i]
    lookup = {tuple(int(x) for x in ix): n for n, ix in enumerate(indices)}
    for r in got:
        n = lookup[tuple(int(x) for x in r["indices"])]
        assert np.allclose(r["position"], pos[n], atol=tol) and np.allclose(r["normal"], nrm[n], atol=tol)
        assert abs(float(r["depth"]) - depth[n]) < tol
    order = [(tuple(int(x) >> 4 for x in r["indices"]), tuple(int(x) for x in r["indices"])) for r in got]
    assert order == sorted(order)


def test_oracle_sphere_contacts_equal_a_float64_evaluation(oracle):
    g = H.asteroid_like_graph(12, 24.0)
    extent = 0.5
    o = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), extent, H.GRADIENT4), 2)
    shape = np.array(o.info()["chunk_counts"]) * 16
    q = H.quat_from_axis_angle([0.3, 1.0, -0.2], 0.7)
    t = np.float32([1.5, -2.0, 0.75])
    # a sphere whose centre, carried into object space, sits near the +x side of the object
    target_obj = extent * (0.5 * shape + [0.36 * shape[0] + 0.3, 2.37, -3.41])  # off the voxel lattice: no borderline floor / ceil
    qc = np.array([-q[0], -q[1], -q[2], q[3]])
    centre = H._rotate(qc, target_obj - t.astype(np.float64)).astype(np.float32)
    got = o.sphere_contacts(q, t, centre, 3.0)
    assert len(got) > 20
    _check_contacts(got, _contacts_float64(o, extent, q, t, centre, 3.0))
    assert len(o.sphere_contacts(q, t, centre + np.float32([500.0, 0, 0]), 3.0)) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere", "asteroid_like", "random"])
def test_gpu_sphere_contacts_equal_the_oracle_bit_for_bit(ctx, oracle, name):
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject

    extent = 0.5
    if name == "random":
        vox, sp, grid = H.random_voxel_chunks((64, 48, 50), 6)
        c = oracle.Object.from_generated_chunks(vox, sp, grid, extent)
        g = VoxelObject.from_generated_chunks(ctx, extent, grid, vox, sp)
    else:
        graph = H.sphere_graph(40.0) if name == "sphere" else H.asteroid_like_graph(16, 36.0)
        c = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(graph.nodes(), graph.root_node_id), extent, H.GRADIENT4), 4)
        g = VoxelObject.generate(SDFVoxelGenerator(extent, ctx.build_generator(graph), H.GRADIENT4))
    shape = np.array(c.info()["chunk_counts"]) * 16
    total = 0
    for case, (axis, angle, t, off, radius) in enumerate([
            ([0.3, 1.0, -0.2], 0.7, [1.5, -2.0, 0.75], [0.4, 0.03, -0.04], 3.0),
            ([0.0, 0.0, 1.0], 0.0, [0.0, 0.0, 0.0], [-0.38, 0.1, 0.05], 6.5),
            ([1.0, 0.2, 0.4], 2.4, [-7.0, 3.0, 11.0], [0.02, -0.41, 0.013], 1.25),
            ([1.0, 0.0, 0.0], 1.0, [0.0, 0.0, 0.0], [3.0, 3.0, 3.0], 2.0)]):  # far outside: nothing
        q = H.quat_from_axis_angle(axis, angle)
        t = np.float32(t)
        target_obj = extent * (0.5 * shape + np.array(off) * shape)
        qc = np.array([-q[0], -q[1], -q[2], q[3]])
        centre = H._rotate(qc, target_obj - t.astype(np.float64)).astype(np.float32)
        got, want = g.sphere_contacts(q, t, centre, radius), c.sphere_contacts(q, t, centre, radius)
        assert len(got) == len(want), (case, len(got), len(want))
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), case
        total += len(got)
    assert total > 50


def test_oracle_plane_contacts_equal_a_float64_evaluation(oracle):
    # for_each_voxel_object_plane_contact (collidable.rs:1176-1209): corner voxels whose sphere reaches below the plane
    g = H.asteroid_like_graph(12, 24.0)
    extent = 0.5
    o = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), extent, H.GRADIENT4), 2)
    shape = np.array(o.info()["chunk_counts"]) * 16
    q = H.quat_from_axis_angle([0.3, 1.0, -0.2], 0.7)
    t = np.float32([1.5, -2.0, 0.75])
    qc = np.array([-q[0], -q[1], -q[2], q[3]])
    # a plane (in the outer space) that cuts the object a few voxels above its lowest point along `normal`
    normal = np.float64([0.2, -0.1, 1.0])
    normal /= np.linalg.norm(normal)
    sv = o.surface_voxels_in_ranges()
    centres_outer = np.array([H._rotate(qc, (ix + 0.5) * extent - t.astype(np.float64)) for ix in sv["indices"].astype(np.float64)])
    heights = centres_outer @ normal
    displacement = float(np.float32(heights.min() + 2.2))
    n32 = normal.astype(np.float32)
    got = o.plane_contacts(q, t, n32, displacement)
    assert len(got) > 5
    # float64: corner surface voxels (<= 3 blocked faces) whose sphere reaches the plane
    blocked = np.zeros(len(sv), np.int32)
    for b in range(2, 8):
        blocked += (sv["flags"] >> b) & 1
    vr = -(sv["sd"].astype(np.float64) * 0.02) * extent
    sd = centres_outer @ n32.astype(np.float64) - displacement
    depth = vr - sd
    want = {tuple(int(x) for x in ix) for ix, dep, bl in zip(sv["indices"], depth, blocked) if bl <= 3 and dep > 1e-4}
    maybe = {tuple(int(x) for x in ix) for ix, dep, bl in zip(sv["indices"], depth, blocked) if bl <= 3 and dep > -1e-4}
    got_set = {tuple(int(x) for x in r) for r in got["indices"]}
    assert want <= got_set <= maybe, (len(want), len(got_set), len(maybe))
    lookup = {tuple(int(x) for x in ix): n for n, ix in enumerate(sv["indices"])}
    for r in got:
        n = lookup[tuple(int(x) for x in r["indices"])]
        assert abs(float(r["depth"]) - depth[n]) < 2e-4
        assert np.allclose(r["position"], centres_outer[n] - sd[n] * n32, atol=2e-4)
        assert np.array_equal(r["normal"], n32)


def test_oracle_capsule_contacts_equal_a_float64_evaluation(oracle):
    # for_each_capsule_voxel_object_contact (collidable.rs:1257-1288): every surface voxel's sphere against the capsule
    g = H.asteroid_like_graph(12, 24.0)
    extent = 0.5
    o = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), extent, H.GRADIENT4), 2)
    shape = np.array(o.info()["chunk_counts"]) * 16
    q = H.quat_from_axis_angle([0.3, 1.0, -0.2], 0.7)
    t = np.float32([1.5, -2.0, 0.75])
    qc = np.array([-q[0], -q[1], -q[2], q[3]])
    # a capsule lying across the +x side of the object (given in the outer space)
    a_obj = extent * (0.5 * shape + [0.36 * shape[0] + 0.3, -6.37, -3.41])
    b_obj = extent * (0.5 * shape + [0.30 * shape[0] + 0.3, 7.21, 4.13])
    a = H._rotate(qc, a_obj - t.astype(np.float64)).astype(np.float32)
    b = H._rotate(qc, b_obj - t.astype(np.float64))
    v = (b - a.astype(np.float64)).astype(np.float32)
    radius = 2.5
    got = o.capsule_contacts(q, t, a, v, radius)
    assert len(got) > 30
    # float64 over the voxels the reference's box admits
    a64, v64 = a.astype(np.float64), v.astype(np.float64)
    s_n = (H._rotate(q, a64) + t.astype(np.float64)) / extent
    e_n = s_n + H._rotate(q, v64) / extent
    occ = o.info()["occupied_voxel_ranges"].astype(np.float64)
    lo = np.maximum(np.floor(np.maximum(np.minimum(s_n, e_n) - radius / extent, 0.0)), occ[:, 0])
    hi = np.minimum(np.ceil(np.maximum(s_n, e_n) + radius / extent), occ[:, 1])
    sv = o.surface_voxels_in_ranges(np.stack([lo, hi], 1).astype(np.uint32))
    vc = np.array([H._rotate(qc, (ix + 0.5) * extent - t.astype(np.float64)) for ix in sv["indices"].astype(np.float64)])
    vr = -(sv["sd"].astype(np.float64) * 0.02) * extent
    par = np.clip(((vc - a64) @ v64) / (v64 @ v64), 0.0, 1.0)
    closest = a64 + par[:, None] * v64
    disp = vc - closest
    dist = np.linalg.norm(disp, axis=1)
    margin = np.abs(dist - (radius + vr))
    hit = dist <= radius + vr
    got_set = {tuple(int(x) for x in r) for r in got["indices"]}
    for i in np.flatnonzero(margin > 1e-3):
        assert (tuple(int(x) for x in sv["indices"][i]) in got_set) == bool(hit[i])
    lookup = {tuple(int(x) for x in ix): n for n, ix in enumerate(sv["indices"])}
    for r in got:
        n = lookup[tuple(int(x) for x in r["indices"])]
        nrm = -disp[n] / dist[n]
        assert np.allclose(r["normal"], nrm, atol=2e-4) and np.allclose(r["position"], vc[n] + vr[n] * nrm, atol=2e-4)
        assert abs(float(r["depth"]) - max(0.0, radius + vr[n] - dist[n])) < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere", "asteroid_like", "random"])
def test_gpu_plane_and_capsule_contacts_equal_the_oracle_bit_for_bit(ctx, oracle, name):
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject

    extent = 0.5
    if name == "random":
        vox, sp, grid = H.random_voxel_chunks((64, 48, 50), 6)
        c = oracle.Object.from_generated_chunks(vox, sp, grid, extent)
        g = VoxelObject.from_generated_chunks(ctx, extent, grid, vox, sp)
    else:
        graph = H.sphere_graph(40.0) if name == "sphere" else H.asteroid_like_graph(16, 36.0)
        c = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(graph.nodes(), graph.root_node_id), extent, H.GRADIENT4), 4)
        g = VoxelObject.generate(SDFVoxelGenerator(extent, ctx.build_generator(graph), H.GRADIENT4))
    shape = np.array(c.info()["chunk_counts"]) * 16
    rng = np.random.default_rng(3)
    n_plane = n_capsule = 0
    for case in range(6):
        q = H.quat_from_axis_angle(rng.normal(size=3), float(rng.uniform(0, 3.0))) if case else np.float32([0, 0, 0, 1])
        t = np.float32(rng.uniform(-8, 8, 3)) if case else np.float32([0, 0, 0])
        qc = np.array([-q[0], -q[1], -q[2], q[3]])
        # plane through a point inside the object, random orientation (outer space)
        normal = rng.normal(size=3)
        normal = (normal / np.linalg.norm(normal)).astype(np.float32)
        p_obj = extent * (0.5 * shape + rng.uniform(-0.3, 0.3, 3) * shape)
        p_outer = H._rotate(qc, p_obj - t.astype(np.float64))
        displacement = float(np.float32(p_outer @ normal.astype(np.float64)))
        got, want = g.plane_contacts(q, t, normal, displacement), c.plane_contacts(q, t, normal, displacement)
        assert len(got) == len(want), ("plane", case, len(got), len(want))
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), ("plane", case)
        n_plane += len(got)
        # capsule between two points around the object's surface
        a_obj = extent * (0.5 * shape + rng.uniform(-0.45, 0.45, 3) * shape)
        b_obj = extent * (0.5 * shape + rng.uniform(-0.45, 0.45, 3) * shape)
        a = H._rotate(qc, a_obj - t.astype(np.float64)).astype(np.float32)
        v = (H._rotate(qc, b_obj - t.astype(np.float64)) - a.astype(np.float64)).astype(np.float32)
        radius = float(rng.uniform(1.0, 5.0))
        got, want = g.capsule_contacts(q, t, a, v, radius), c.capsule_contacts(q, t, a, v, radius)
        assert len(got) == len(want), ("capsule", case, len(got), len(want))
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), ("capsule", case)
        n_capsule += len(got)
    # a degenerate capsule (zero segment) is a sphere's contact set
    centre, r = H._rotate(np.float64([0, 0, 0, 1]), extent * 0.5 * shape + [extent * 0.4 * shape[0], 0.3, 0.2]).astype(np.float32), 3.0
    gs = g.capsule_contacts([0, 0, 0, 1], [0, 0, 0], centre, [0, 0, 0], r)
    cs = c.capsule_contacts([0, 0, 0, 1], [0, 0, 0], centre, [0, 0, 0], r)
    assert np.array_equal(gs.view(np.uint8), cs.view(np.uint8))
    assert {tuple(x) for x in gs["indices"]} == {tuple(x) for x in g.sphere_contacts([0, 0, 0, 1], [0, 0, 0], centre, r)["indices"]}
    assert n_plane > 20 and n_capsule > 50
