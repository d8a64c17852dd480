"""GPU parity for the inertial moments (`ivx_object_inertial_moments`, object/inertia.rs:125-137, 629-789): the ten
f32 sums and every chunk's ten terms equal the oracle's bit for bit — fresh objects, after sphere / capsule
absorption (converted and removed chunks), on a split-off fragment, and chained over x-slabs. At full size, where
the oracle is too slow, the sums are checked against the closed forms of a ball."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import inertia as I
from impact_b200 import workloads as W
from impact_b200.voxel import SDFVoxelGenerator, VoxelObject

pytestmark = pytest.mark.gpu

DENS4 = [1.0, 2.7, 0.3, 5.5]
CASES = {
    "sphere64": (lambda: H.sphere_graph(31.0), H.SAME0, 1.0, [0.5]),
    "sphere_big_interior": (lambda: H.sphere_graph(70.0), H.SAME0, 0.1, [3.25]),   # uniform chunks' closed form
    "box_types": (lambda: H.box_graph(45.0), H.GRADIENT4, 0.37, DENS4),
    "zoo": (H.csg_zoo_graph, H.GRADIENT4, 1.0, DENS4),
    "asteroid_like": (lambda: H.asteroid_like_graph(24, 40.0), H.GRADIENT4, 0.25, DENS4),
    "asteroid_stand_in": (lambda: W.asteroid_stand_in(0.6), H.GRADIENT4, 1.0, DENS4),
}


def _both(ctx, oracle, graph, types, extent):
    gen_gpu = ctx.build_generator(graph)
    gen_cpu = oracle.Generator(graph.nodes(), graph.root_node_id)
    obj_cpu = oracle.Object.generate(oracle.VoxelGenerator(gen_cpu, extent, types), 4)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(extent, gen_gpu, types))
    return obj_gpu, obj_cpu


def _assert_same_moments(obj_gpu, obj_cpu, dens, what=""):
    tg, pg = obj_gpu.inertial_moments(dens, per_chunk=True)
    tc, pc = obj_cpu.inertial_moments(dens, per_chunk=True)
    bad = ~H.f32_bits_equal(pg, pc).all(axis=1)
    assert not bad.any(), f"{what}: chunk terms differ in {bad.sum()} chunks, first {np.flatnonzero(bad)[0]}: " \
                          f"{pg[bad][0]} vs {pc[bad][0]}"
    assert H.f32_bits_equal(tg, tc).all(), (what, tg, tc)
    return tg


@pytest.mark.parametrize("name", sorted(CASES))
def test_inertial_moments_are_bit_exact(ctx, oracle, name):
    make, types, extent, dens = CASES[name]
    obj_gpu, obj_cpu = _both(ctx, oracle, make(), types, extent)
    total = _assert_same_moments(obj_gpu, obj_cpu, dens, name)
    assert total[0] > 0
    # the call is a pure function of the object: a second call and the manager mirror agree
    m = I.VoxelObjectInertialPropertyManager.initialized_from(obj_gpu, dens)
    assert H.f32_bits_equal(m.m, total).all()
    props = m.derive_inertial_properties()
    assert np.all(np.linalg.eigvalsh(props.inertia_tensor.astype(np.float64)) > 0)


def test_inertial_moments_follow_absorption(ctx, oracle):
    g = H.asteroid_like_graph(16, 36.0)
    obj_gpu, obj_cpu = _both(ctx, oracle, g, H.GRADIENT4, 0.5)
    shape = (np.array(obj_cpu.info()["chunk_counts"]) * 16).astype(np.float32)
    f = np.float32
    m_inc = obj_cpu.inertial_moments(DENS4).copy()
    for step in range(4):
        c = (f(0.5) * shape + f([11.0 * step - 30.0, 2.5, -4.75])).astype(np.float32)
        obj_cpu.absorb_sphere_inertial(c, 9.0, 11.0, DENS4, m_inc)
        obj_gpu.absorb_sphere(c, 9.0, 11.0)
        total = _assert_same_moments(obj_gpu, obj_cpu, DENS4, f"sphere step {step}")
        # the reference's incremental updater (absorption.rs:836-840) and the from-scratch sums agree to its own
        # validate_for_object tolerance (intersection.rs:1022)
        assert np.all(np.abs(total - m_inc) <= 1e-3 * np.maximum(np.abs(total), 1.0))
    a, v = f(0.5) * shape - f([3.0, 40.0, 1.0]), f([5.5, 80.0, 2.5])
    obj_cpu.absorb_capsule(a, v, 9.0, 11.0)
    obj_gpu.absorb_capsule(a, v, 9.0, 11.0)
    _assert_same_moments(obj_gpu, obj_cpu, DENS4, "capsule")


@pytest.mark.parametrize("name,types,dens,extent", [("sphere", H.SAME0, [0.5], 1.0), ("asteroid_like", H.GRADIENT4, DENS4, 0.25)])
def test_absorption_updates_the_moments_bit_for_bit(ctx, oracle, name, types, dens, extent):
    # apply_sphere_absorption / apply_capsule_absorption with the VoxelObjectInertialPropertyUpdater attached
    # (absorption.rs:801-889): the manager's ten sums after every call, against the oracle's voxel-by-voxel chain
    g = H.sphere_graph(40.0) if name == "sphere" else H.asteroid_like_graph(16, 36.0)
    obj_gpu, obj_cpu = _both(ctx, oracle, g, types, extent)
    f = np.float32
    shape = (np.array(obj_cpu.info()["chunk_counts"]) * 16).astype(np.float32)
    mid = f(0.5) * shape
    m_cpu = obj_cpu.inertial_moments(dens).copy()
    m_gpu = obj_gpu.inertial_moments(dens).copy()
    assert H.f32_bits_equal(m_cpu, m_gpu).all()
    R = f(0.5) * shape.max()
    radius = f(0.15) * R
    start = (mid - R / f(np.sqrt(3.0))).astype(np.float32)
    removed = 0
    for step in range(5):  # BASELINE config 5 geometry: the absorber marches inward along the diagonal
        c = (start + f(step) * radius * f(0.6)).astype(np.float32)
        st_c = obj_cpu.absorb_sphere_inertial(c, float(radius), float(radius + f(2.0)), dens, m_cpu)
        st_g = obj_gpu.absorb_sphere_inertial(c, float(radius), float(radius + f(2.0)), dens, m_gpu)
        for k in ("touched_chunks", "touched_voxels", "emptied_voxels", "removed_chunks"):
            assert st_g[k] == st_c[k], (step, k, st_g, st_c)
        removed += st_c["emptied_voxels"]
        assert H.f32_bits_equal(m_gpu, m_cpu).all(), (step, m_gpu, m_cpu)
    capsules = [
        (mid - f([30.5, 8.25, 3.0]), f([61.0, 16.5, 6.0]), f(5.0)),      # slanted, through the centre
        (mid + f([9.1, -7.7, 11.3]), f([0.0, 0.0, 0.0]), f(6.0)),        # zero-length segment
        (f([-50.0, -50.0, -50.0]), f([10.0, 0.0, 0.0]), f(4.0)),         # misses the object: nothing changes
        (mid - f([3.0, 40.0, 1.0]), f([5.5, 80.0, 2.5]), f(9.0)),        # thick, crosses everything again
    ]
    for step, (a, v, rad) in enumerate(capsules):
        st_c = obj_cpu.absorb_capsule_inertial(a, v, float(rad), float(rad + f(2.0)), dens, m_cpu)
        st_g = obj_gpu.absorb_capsule_inertial(a, v, float(rad), float(rad + f(2.0)), dens, m_gpu)
        assert st_g["emptied_voxels"] == st_c["emptied_voxels"]
        removed += st_c["emptied_voxels"]
        assert H.f32_bits_equal(m_gpu, m_cpu).all(), ("capsule", step, m_gpu, m_cpu)
    assert removed > 20000
    # the voxels themselves went through the same path as without the updater
    H.assert_objects_equal(*obj_gpu.download(), obj_cpu.chunks(), obj_cpu.voxels())
    # and the incremental sums stay within the reference's validate_for_object tolerance of the from-scratch ones
    scratch = _assert_same_moments(obj_gpu, obj_cpu, dens, "after absorption")
    assert np.all(np.abs(scratch - m_gpu) <= 1e-3 * np.maximum(np.abs(scratch), 1.0))


def test_inertial_moments_of_a_split_off_fragment(ctx, oracle):
    # extract_any_disconnected_region (extraction.rs:78-113): both parts' sums from scratch, and the bookkeeping the
    # reference's PropertyTransferrer does incrementally (inertia.rs:397-470) holds between them
    from test_oracle_split_detection import two_spheres_graph

    f = np.float32
    extent = 0.5
    obj_gpu, obj_cpu = _both(ctx, oracle, two_spheres_graph(100.0, 40.0, 36.0), H.GRADIENT4, extent)
    before = _assert_same_moments(obj_gpu, obj_cpu, DENS4, "two spheres")
    info_c, frag_c = obj_cpu.extract_any_disconnected_region()
    info_g, frag_g = obj_gpu.extract_any_disconnected_region()
    assert info_c["extracted"] and info_g["extracted"]
    rest = _assert_same_moments(obj_gpu, obj_cpu, DENS4, "parent after extraction")
    part = _assert_same_moments(frag_g, frag_c, DENS4, "fragment")
    # the fragment's sums are about its own grid origin: move the reference point back to the parent's origin
    mgr = I.VoxelObjectInertialPropertyManager(part)
    mgr.offset_reference_point_by(-f(extent) * f(info_c["origin_offset_in_parent"]))
    assert np.allclose(mgr.m + rest, before, rtol=1e-3, atol=1e-3 * float(np.abs(before).max()))


def test_slabs_chain_to_the_whole_objects_sums(ctx, oracle):
    import torch
    from test_gpu_slabs import _local_exchange

    g = H.asteroid_like_graph(24, 40.0)
    gen = ctx.build_generator(g)
    vg = SDFVoxelGenerator(0.25, gen, H.GRADIENT4)
    whole = VoxelObject.generate(vg)
    want = whole.inertial_moments(DENS4)
    cx = whole.info()["chunk_counts"][0]
    cuts = [0, cx // 3, cx // 3, (2 * cx) // 3 + 1, cx]  # includes an empty slab
    ranges = list(zip(cuts[:-1], cuts[1:]))
    objs = [VoxelObject.generate(vg, r) for r in ranges]
    with pytest.raises(Exception):
        objs[0].inertial_moments(DENS4)  # derived state pending: uniform chunks may still convert
    _local_exchange(objs, ranges)
    torch.cuda.synchronize()
    acc = None
    for o in objs:
        acc = o.inertial_moments(DENS4, initial=acc)
    assert H.f32_bits_equal(acc, want).all(), (acc, want)


def test_missing_density_is_an_error(ctx):
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(H.box_graph(45.0)), H.GRADIENT4))
    with pytest.raises(Exception, match="density"):
        obj.inertial_moments([1.0, 2.0])
    assert obj.inertial_moments([1.0, 2.0, 3.0, 4.0])[0] > 0
    m = obj.inertial_moments([1.0, 2.0, 3.0, 4.0]).copy()
    keep = m.copy()
    with pytest.raises(Exception, match="density"):
        obj.absorb_sphere_inertial([24.0, 24.0, 24.0], 6.0, 8.0, np.zeros(0, np.float32), m)
    assert np.array_equal(m, keep)
    empty = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(H.SDFGraph()), H.SAME0))
    assert not empty.inertial_moments([1.0]).any()


def test_full_size_ball_has_a_balls_inertial_properties(ctx):
    # BASELINE-size object without the oracle: a ball of radius 500 voxels (1002^3 grid, 250k chunks)
    r, e, rho = 500.0, 0.1, 2.5
    obj = VoxelObject.generate(SDFVoxelGenerator(e, ctx.build_generator(H.sphere_graph(r)), H.SAME0))
    inf = obj.info()
    # chunks inside the sphere's inscribed cube are filled with -margin = -2.54 → code -127, not -128: they are stored
    # NonUniform like in the reference (atomic.rs:663-668, object.rs:1913-1918); only the shell between the cube and
    # the surface is Uniform
    assert inf["n_uniform"] > 60000 and inf["n_non_uniform"] > 50000
    m = I.VoxelObjectInertialPropertyManager.initialized_from(obj, [rho])
    props = m.derive_inertial_properties()
    R = r * e
    mass = rho * 4.0 / 3.0 * np.pi * R ** 3
    centre = e * 0.5 * np.float64(inf["grid_shape"])
    # the f32 chain over 2.5e5 chunk terms is the reference's own arithmetic: its rounding can reach N eps / 2 ~ 1e-2 in the
    # worst case (typically 1e-4); voxelisation of the ball adds ~1e-3
    assert abs(props.mass - mass) <= 1e-2 * mass
    assert np.all(np.abs(props.center_of_mass - centre) <= 1e-2 * R)
    want = 0.4 * mass * R * R
    assert np.all(np.abs(np.diag(props.inertia_tensor) - want) <= 3e-2 * want)
    off = props.inertia_tensor - np.diag(np.diag(props.inertia_tensor))
    assert np.all(np.abs(off) <= 3e-2 * want)
