"""Oracle pinning for the inertial moments (object/inertia.rs): the reference's own three tests
(inertia.rs:805-963) restated on the oracle, an independent float64 integration, and the incremental
updater against a from-scratch integration at the reference's `validate_for_object` tolerance
(intersection.rs:1022, 1068: 1e-3). CPU only."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import inertia as I


def _full_chunk(oracle, voxel_type=0):
    v = np.zeros(4096, oracle.VOXEL_DTYPE)
    v["type"] = voxel_type
    v["sd"] = -128       # Voxel::maximally_inside
    v["flags"] = 0xFC
    return v


def _rel_close(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    # approx::relative_eq with epsilon = max_relative = tol
    return np.all((np.abs(a - b) <= tol) | (np.abs(a - b) <= tol * np.maximum(np.abs(a), np.abs(b))))


def test_full_non_uniform_chunk_has_same_inertial_properties_as_uniform_chunk(oracle):
    # inertia.rs:805-853
    nu = oracle.moments_for_non_uniform_chunk(0.1, _full_chunk(oracle), [0.5], [1, 2, 3])
    un = oracle.moments_for_uniform_chunk(0.1, [0.5], 0, [1, 2, 3])
    assert un[0] > 0
    assert _rel_close(nu, un, 1e-3), (nu, un)


def test_box_voxel_object_has_box_inertial_properties(oracle):
    # inertia.rs:855-911
    extent = 0.1
    g = H.SDFGraph()
    g.box([22.0, 27.0, 19.0])
    gen = oracle.Generator(g.nodes(), g.root_node_id)
    obj = oracle.Object.generate(oracle.VoxelGenerator(gen, extent, H.SAME0), 2)
    r = obj.info()["occupied_voxel_ranges"].astype(np.float64)
    extents = extent * (r[:, 1] - r[:, 0])
    centers = 0.5 * extent * (r[:, 0] + r[:, 1])
    props = I.VoxelObjectInertialPropertyManager(obj.inertial_moments([0.5])).derive_inertial_properties()
    box = I.InertialProperties.of_uniform_box(*extents, 0.5).translated(centers)
    assert _rel_close(props.mass, box.mass, 1e-3)
    assert _rel_close(props.center_of_mass, box.center_of_mass, 1e-3)
    assert _rel_close(props.inertia_tensor, box.inertia_tensor, 1e-3), (props.inertia_tensor, box.inertia_tensor)
    prod = props.inertia_tensor.astype(np.float64) @ props.inverse_inertia_tensor.astype(np.float64)
    assert np.allclose(prod, np.eye(3), atol=1e-4)


def test_chunk_has_zero_moments_after_removing_each_voxel(oracle):
    # inertia.rs:913-962, with the host mirror's updater (whose voxel term must equal the oracle's bit for bit)
    extent, dens, cc = np.float32(0.1), [0.5], (1, 2, 3)
    m = I.VoxelObjectInertialPropertyManager(oracle.moments_for_non_uniform_chunk(extent, _full_chunk(oracle), dens, cc))
    upd = m.begin_update(extent, dens)
    for i in range(16):
        for j in range(16):
            for k in range(16):
                upd.remove_voxel((cc[0] * 16 + i, cc[1] * 16 + j, cc[2] * 16 + k), 0)
    assert np.all(np.abs(m.m) <= 1e-3), m.m
    rng = np.random.default_rng(0)
    for _ in range(50):
        ijk = rng.integers(0, 1000, 3)
        e = np.float32(rng.choice([1.0, 0.25, 0.1, 0.37]))
        t = int(rng.integers(0, 3))
        a = oracle.moments_for_voxel(e, [0.5, 2.25, 7.0], ijk, t)
        b = I.compute_moments_for_voxel(e, e * e, (e * e) * e, np.float32([0.5, 2.25, 7.0]), ijk, t)
        assert H.f32_bits_equal(a, b).all(), (ijk, e, a, b)


def _dense_view(oracle, obj):
    """(nx, ny, nz) arrays of emptiness and type from the oracle object."""
    cc = obj.info()["chunk_counts"]
    chunks, voxels = obj.chunks(), obj.voxels().reshape(-1, 4096)
    shape = tuple(int(c) * 16 for c in cc)
    solid = np.zeros(shape, bool)
    types = np.zeros(shape, np.uint8)
    for lin, c in enumerate(chunks):
        i, j, k = lin // (cc[1] * cc[2]), (lin // cc[2]) % cc[1], lin % cc[2]
        sl = (slice(16 * i, 16 * i + 16), slice(16 * j, 16 * j + 16), slice(16 * k, 16 * k + 16))
        if c["kind"] == 1:
            solid[sl] = True
            types[sl] = c["uniform_type"]
        elif c["kind"] == 2:
            v = voxels[c["data_offset"]].reshape(16, 16, 16)
            solid[sl] = (v["flags"] & 1) == 0
            types[sl] = v["type"]
    return solid, types


def _float64_moments(solid, types, densities, e):
    """The integrals the ten sums approximate, in float64 from the closed forms of a cube (no shared code)."""
    d = np.asarray(densities, np.float64)[np.where(solid, types, 0)] * solid  # empty voxels may carry type 255
    n = solid.shape
    ax = [np.arange(s, dtype=np.float64) * e for s in n]
    lo = np.meshgrid(*ax, indexing="ij", sparse=True)
    c1 = [x + 0.5 * e for x in lo]                                # ∫ x dx / e
    c2 = [((x + e) ** 3 - x ** 3) / (3.0 * e) for x in lo]         # ∫ x² dx / e
    vol = e ** 3
    out = np.zeros(10)
    out[0] = (d * vol).sum()
    for a in range(3):
        out[1 + a] = (d * vol * c1[a]).sum()
    out[4] = (d * vol * (c2[1] + c2[2])).sum()
    out[5] = (d * vol * (c2[0] + c2[2])).sum()
    out[6] = (d * vol * (c2[0] + c2[1])).sum()
    out[7] = (d * vol * c1[0] * c1[1]).sum()
    out[8] = (d * vol * c1[1] * c1[2]).sum()
    out[9] = (d * vol * c1[2] * c1[0]).sum()
    return out


@pytest.mark.parametrize("name,extent", [("sphere", 1.0), ("asteroid_like", 0.25), ("box", 0.1)])
def test_object_moments_match_a_float64_integration(oracle, name, extent):
    g = {"sphere": lambda: H.sphere_graph(31.0), "asteroid_like": lambda: H.asteroid_like_graph(12, 24.0),
         "box": lambda: H.box_graph(45.0)}[name]()
    types = H.SAME0 if name == "sphere" else H.GRADIENT4
    dens = [0.5] if name == "sphere" else [1.0, 2.7, 0.3, 5.5]
    gen = oracle.Generator(g.nodes(), g.root_node_id)
    obj = oracle.Object.generate(oracle.VoxelGenerator(gen, extent, types), 4)
    total, per_chunk = obj.inertial_moments(dens, per_chunk=True)
    ref = _float64_moments(*_dense_view(oracle, obj), dens, extent)
    assert ref[0] > 0
    # sequential f32 sums over up to 10^5 terms: a few 1e-5 relative
    assert np.all(np.abs(total - ref) <= 2e-4 * np.abs(ref)), (total, ref)
    # the total is the chunk terms added in linear chunk order
    acc = np.zeros(10, np.float32)
    for row in per_chunk:
        acc = acc + row
    assert H.f32_bits_equal(acc, total).all()
    kinds = obj.chunks()["kind"]
    assert not per_chunk[kinds == 0].any()


def test_incremental_update_during_absorption_stays_within_the_reference_tolerance(oracle):
    # intersection.rs:1000-1070: absorb, then validate_for_object(…, 1e-3)
    g = H.asteroid_like_graph(12, 30.0)
    dens = [1.0, 2.7, 0.3, 5.5]
    gen = oracle.Generator(g.nodes(), g.root_node_id)
    obj = oracle.Object.generate(oracle.VoxelGenerator(gen, 0.5, H.GRADIENT4), 4)
    m = obj.inertial_moments(dens).copy()
    shape = np.array(obj.info()["chunk_counts"]) * 16
    removed = 0
    for step in range(4):
        c = (0.5 * shape + np.float32([9.0 * step - 20.0, 3.5, -6.25])).astype(np.float32)
        st = obj.absorb_sphere_inertial(c, 7.0, 9.0, dens, m)
        removed += st["emptied_voxels"]
        scratch = obj.inertial_moments(dens)
        assert _rel_close(m, scratch, 1e-3), (step, m, scratch)
    st = obj.absorb_capsule_inertial(0.5 * shape - np.float32([25, 0, 0]), np.float32([50, 4, 2]), 4.0, 6.0, dens, m)
    removed += st["emptied_voxels"]
    assert removed > 1000
    assert _rel_close(m, obj.inertial_moments(dens), 1e-3)


def test_manager_reference_point_and_addition():
    # offset_reference_point_by (inertia.rs:255-268): moving the reference point there and back is the identity up
    # to rounding, and the centre-of-mass tensor does not depend on the reference point
    m = I.VoxelObjectInertialPropertyManager(np.float32([12.0, 30.0, -18.0, 6.0, 400.0, 380.0, 290.0, -40.0, 9.0, 14.0]))
    before = m.derive_inertial_properties()
    moved = I.VoxelObjectInertialPropertyManager(m.m)
    moved.offset_reference_point_by([1.5, -2.0, 0.25])
    after = moved.derive_inertial_properties()
    assert np.allclose(after.inertia_tensor, before.inertia_tensor, rtol=1e-4, atol=1e-3)
    assert np.allclose(after.center_of_mass, before.center_of_mass - np.float32([1.5, -2.0, 0.25]), atol=1e-5)
    moved.offset_reference_point_by([-1.5, 2.0, -0.25])
    assert np.allclose(moved.m, m.m, rtol=1e-5, atol=1e-3)
    s = m.add(m)
    assert np.array_equal(s.m, m.m + m.m)
    assert np.allclose(s.derive_center_of_mass(), m.derive_center_of_mass())
