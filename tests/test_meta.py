"""Meta-graph compiler (host mirror of meta.rs) and its GPU surface probes."""
import os

import numpy as np
import pytest

import helpers as H
from impact_b200 import meta as M

REF_RON = "/root/reference/engine/benches/data/asteroid.vgen.ron"


def _norm(x):
    if isinstance(x, M.Tagged):
        return (x.tag, _norm(x.fields))
    if isinstance(x, dict):
        return {k: _norm(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_norm(v) for v in x]
    if isinstance(x, (int, float)):
        return float(x)
    return x


@pytest.mark.skipif(not os.path.exists(REF_RON), reason="reference tree not present on this machine")
def test_transcribed_asteroid_graph_equals_the_reference_ron_file():
    assert _norm(M.load_vgen_ron(REF_RON)) == _norm(M.asteroid_meta_nodes())


def test_ron_reader_handles_the_generator_format():
    doc = M.parse_ron("""( sdf_graph: ( nodes: [ Spheres(( radius: Uniform( min: Fixed(1.5), max: Fixed(2), ),
        count: 3, seed: 7, sampling: PerInstance, )), // comment
        SDFInstantiation(( child_id: 0, )), ], ), )""")
    nodes = doc["sdf_graph"]["nodes"]
    assert [n.tag for n in nodes] == ["Spheres", "SDFInstantiation"]
    assert nodes[0]["radius"].tag == "Uniform" and nodes[0]["radius"]["min"].fields == 1.5
    assert nodes[0]["sampling"].tag == "PerInstance" and nodes[1]["child_id"] == 0


REF_GRAPH_RON = "/root/reference/apps/voxel_generator/examples/asteroid.graph.ron"


@pytest.mark.skipif(not os.path.exists(REF_GRAPH_RON), reason="reference tree not present on this machine")
def test_editor_graph_file_gives_the_same_meta_nodes_as_the_generator_file():
    # apps/voxel_generator/examples/asteroid.graph.ron is the editor's copy of engine/benches/data/asteroid.vgen.ron:
    # loading it through the editor's build rules must give the identical meta graph, node for node, in the same order
    nodes, voxel_extent, scale_factor, seed = M.load_graph_ron(REF_GRAPH_RON)
    assert (voxel_extent, scale_factor, seed) == (0.25, 1.0, 346)
    assert _norm(nodes) == _norm(M.load_vgen_ron(REF_RON))


@pytest.mark.skipif(not os.path.exists(REF_GRAPH_RON), reason="reference tree not present on this machine")
def test_editor_subgraph_fragments():
    # *.subgraph.ron (io.rs:31-39): asteroid_base is the first seven nodes of the asteroid graph; the crater fragments
    # have an open input (the surface they are cast onto) and cannot be built on their own
    ex = os.path.dirname(REF_GRAPH_RON)
    base, *_ = M.load_graph_ron(os.path.join(ex, "asteroid_base.subgraph.ron"))
    full, *_ = M.load_graph_ron(REF_GRAPH_RON)
    assert _norm(base) == _norm(full[:7])
    for name in ("big_craters", "medium_craters", "small_craters"):
        with pytest.raises(ValueError, match="unattached"):
            M.load_graph_ron(os.path.join(ex, f"{name}.subgraph.ron"))


def test_editor_graph_reader_follows_the_editors_build_rules():
    # tests/golden/mini.graph.ron: post-order ids, a child shared by two parents added once, enum variants by index,
    # discrete `fixed as u32`, FromParam sources, every distribution variant
    with open(os.path.join(os.path.dirname(__file__), "golden", "mini.graph.ron")) as f:
        nodes, voxel_extent, scale_factor, seed = M.graph_ron_nodes(f.read())
    assert (voxel_extent, scale_factor, seed) == (0.5, 2.0, 7)
    assert [n.tag for n in nodes] == ["Spheres", "SDFInstantiation", "SDFGroupUnion", "Scaling", "SDFInstantiation",
                                      "SDFGroupUnion", "SDFUnion", "MultifractalNoiseSDFModifier"]
    sph, inst1, gu1, scal, inst2, gu2, uni, noise = nodes
    assert (gu1["child_id"], gu1["smoothness"], gu2["child_id"], gu2["smoothness"]) == (1, 0.0, 4, 0.25)
    assert sph["count"] == 1 and sph["seed"] == 3 and sph["sampling"].tag == "OnlyOnce"
    assert sph["radius"].tag == "Constant" and sph["radius"].fields.fields == 12.0
    assert sph["center_z"].tag == "UniformCosAngle" and sph["center_z"]["max_angle"].fields == 20.0
    assert inst1["child_id"] == 0 and scal["child_id"] == 0 and inst2["child_id"] == 3
    assert scal["composition"].tag == "Pre" and scal["sampling"].tag == "OnlyOnce" and scal["seed"] == 11
    assert scal["scaling"].tag == "PowerLaw" and scal["scaling"]["exponent"].fields == 2.5
    assert (uni["child_1_id"], uni["child_2_id"], uni["smoothness"]) == (2, 5, 1.5)
    assert noise["child_id"] == 6 and noise["sampling"].tag == "PerInstance"
    assert noise["octaves"].fields.fields == 3 and isinstance(noise["octaves"].fields.fields, int)
    amp = noise["amplitude"].fields
    assert amp.tag == "FromParam" and amp["idx"] == 1 and amp["mapping"]["offset"] == 1.0 and amp["mapping"]["scale"] == 100.0
    assert noise["frequency"].tag == "Uniform" and noise["frequency"]["min"].fields == 0.01
    with pytest.raises(ValueError):
        M.graph_ron_nodes("(kind: Subgraph(root_node_id: 1), nodes: [], collapsed_nodes: [])")
    # and the file compiles to an atomic graph with the editor's own scale factor and seed (build.rs:102-128)
    g, ve = M.compile_graph_file(os.path.join(os.path.dirname(__file__), "golden", "mini.graph.ron"))
    kinds = [int(k) for k in g.nodes()["kind"]]
    assert ve == 0.5 and len(g) >= 4 and g.root_node_id == len(g) - 1
    from impact_b200 import workloads as W
    assert all(20 < d < 160 for d in W.grid_shape_of(g)), W.grid_shape_of(g)


def test_parameter_evaluation_order_is_topological_fifo():
    # params.rs:266-330: parameters without dependencies first (index order), dependents as they become ready
    node = M.asteroid_meta_nodes()[7]  # Capsules: segment_length and center_y depend on radius (idx 1)
    rng = M.Rng(1)
    p = M.sample_params(node, ["segment_length", "radius", "center_x", "center_y", "center_z"], rng)
    assert p["segment_length"] == p["radius"]
    assert np.float32(0.2) * p["radius"] <= p["center_y"] <= np.float32(0.3) * p["radius"]
    assert 50.0 <= p["radius"] <= 80.0
    cyc = M.Tagged("X", {"a": M.Tagged("Constant", M._from_param(1, 1.0)), "b": M.Tagged("Constant", M._from_param(0, 1.0))})
    with pytest.raises(ValueError, match="Cycle"):
        M.sample_params(cyc, ["a", "b"], M.Rng(0))


def test_rng_and_stable_seeds_are_deterministic():
    a, b = M.Rng(42), M.Rng(42)
    assert [a.gen_u64() for _ in range(4)] == [b.gen_u64() for _ in range(4)]
    r = M.Rng(3)
    xs = [r.f32() for _ in range(1000)]
    assert all(0.0 <= x < 1.0 for x in xs) and 0.4 < float(np.mean(xs)) < 0.6
    assert all(3 <= M.Rng(s).u32_inclusive(3, 6) <= 6 for s in range(50))
    assert M.splitmix(0) == 0xE220A8397B1DCDAF  # SplitMix64 reference value for state 0


def test_body_of_the_asteroid_compiles_without_a_device():
    g = M.MetaCompiler(M.asteroid_meta_nodes()[:7], 1.0, 0).build()
    kinds = list(g.nodes()["kind"])
    n_spheres = kinds.count(0)
    assert 3 <= n_spheres <= 6                   # StochasticSelection picks 3..6 of the 8 spheres
    assert kinds.count(7) == n_spheres - 1        # balanced union tree
    assert kinds[-1] == 6 and g.root_node_id == len(kinds) - 1
    g2 = M.MetaCompiler(M.asteroid_meta_nodes()[:7], 1.0, 0).build()
    assert np.array_equal(g.nodes(), g2.nodes())
    g3 = M.MetaCompiler(M.asteroid_meta_nodes()[:7], 1.0, 1).build()
    assert not np.array_equal(g.nodes()["p"], g3.nodes()["p"]) or len(g3) != len(g)


def test_errors_mirror_the_reference():
    with pytest.raises(ValueError, match="Root meta node must have single SDF output"):
        M.MetaCompiler([M.Tagged("Points", {"count": 2})], 1.0, 0).build()
    cyc = [M.Tagged("SDFGroupUnion", {"child_id": 0, "smoothness": 1.0})]
    with pytest.raises(ValueError, match="cycle"):
        M.MetaCompiler(cyc, 1.0, 0).build()


@pytest.mark.gpu
def test_block_probes_match_the_oracle_bit_for_bit(ctx, oracle):
    g = H.csg_zoo_graph()
    gen = ctx.build_generator(g)
    ogen = oracle.Generator(g.nodes(), g.root_node_id)
    rng = np.random.default_rng(0)
    org = rng.uniform(-30, 30, (200, 3)).astype(np.float32)
    for size in (1, 2):
        got = gen.compute_signed_distances_for_blocks_preserving_gradients(org, size)
        for i, o in enumerate(org):
            want = ogen.eval_block_preserving_gradients(o, size)
            assert H.f32_bits_equal(got[i], want).all(), (size, i)


@pytest.mark.gpu
def test_asteroid_compiles_and_matches_the_oracle_downstream(ctx, oracle):
    from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

    nodes = M.asteroid_meta_nodes()
    g = M.MetaCompiler(nodes, 0.28, 0, ctx).build()
    kinds = np.bincount(g.nodes()["kind"], minlength=10)
    n_caps = int(kinds[1])
    # 40 + 150 + 250 capsules are cast along their own (tilted up to 80 degrees) axis from a shell 2.5-3.5x the body
    # size; most rays miss the body and those instances are dropped (meta.rs:1779-1784)
    assert 5 <= n_caps <= 440
    n_passes = int(kinds[8])
    assert 3 <= kinds[0] <= 6 and kinds[6] == 2 and 1 <= n_passes <= 3
    assert kinds[7] == (kinds[0] - 1) + (n_caps - n_passes)  # balanced union trees
    # the placed capsules touch the surface they were cast onto: re-probing gives |sd| <= tolerance-ish
    gen = ctx.build_generator(g)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, H.GRADIENT4))
    ogen = oracle.Generator(g.nodes(), g.root_node_id)
    obj_cpu = oracle.Object.generate(oracle.VoxelGenerator(ogen, 1.0, H.GRADIENT4), 8)
    gch, gvx = obj_gpu.download()
    H.assert_objects_equal(gch, gvx, obj_cpu.chunks(), obj_cpu.voxels())
    H.assert_meshes_equal(VoxelObjectMesh.create(obj_gpu).download(), obj_cpu.mesh(8))
    # same seed → same graph; different seed → different asteroid
    g2 = M.MetaCompiler(nodes, 0.28, 0, ctx).build()
    assert np.array_equal(g.nodes(), g2.nodes())


@pytest.mark.gpu
def test_committed_bench_graphs_equal_a_fresh_compile(ctx):
    path = os.path.join(M.DATA_DIR, "asteroid_128_seed0.npz")
    if not os.path.exists(path):
        pytest.skip("bench graphs not generated yet")
    cached = M.asteroid_graph_scaled(112, 128, 0)
    fresh = M.asteroid_graph_scaled(112, 128, 0, ctx=ctx, use_cache=False) if False else None
    M._CACHE.clear()
    fresh = M.asteroid_graph_scaled(112, 128, 0, ctx=ctx, use_cache=False)
    M._CACHE.clear()
    assert np.array_equal(cached.nodes(), fresh.nodes()) and cached.root_node_id == fresh.root_node_id


# ---- Similarity, TransformApplication (host only); ClosestTranslationToSurface, RotationToGradient (device probes) ------
def _spheres(count, radius, seed=0):
    T = M.Tagged
    return T("Spheres", {"radius": M._const(radius), "center_x": M._const(0.0), "center_y": M._const(0.0),
                         "center_z": M._const(0.0), "count": count, "seed": seed, "sampling": T("OnlyOnce", None)})


def test_similarity_composes_scale_rotation_translation_like_similarity3():
    # meta.rs:1402-1446: Similarity3::from_parts(translation * scale_factor, tilt/turn/roll, scale), Post = T * instance
    T = M.Tagged
    sim = lambda comp: T("Similarity", {
        "child_id": 1, "composition": T(comp, None), "scale": M._const(2.0), "tilt_angle": M._const(90.0),
        "turn_angle": M._const(0.0), "roll_angle": M._const(0.0), "translation_x": M._const(5.0),
        "translation_y": M._const(0.0), "translation_z": M._const(0.0), "seed": 3, "sampling": T("OnlyOnce", None)})
    tr = T("Translation", {"child_id": 0, "composition": T("Post", None), "translation_x": M._const(0.0),
                           "translation_y": M._const(10.0), "translation_z": M._const(0.0), "seed": 1,
                           "sampling": T("OnlyOnce", None)})
    out = {}
    for comp in ("Post", "Pre"):
        nodes = [_spheres(1, 4.0), tr, sim(comp), T("SDFInstantiation", {"child_id": 2}),
                 T("SDFGroupUnion", {"child_id": 3, "smoothness": 0.0})]
        g = M.MetaCompiler(nodes, 3.0, 0).build()
        n = g.nodes()
        # sphere(r * scale_factor) -> scaling(2) -> rotation -> translation
        assert list(n["kind"]) == [0, 5, 4, 3] and n[0]["p"][0] == np.float32(12.0) and n[1]["p"][0] == np.float32(2.0)
        out[comp] = n[3]["p"][:3].copy()
    # Post: the earlier translation (0, 30, 0) is scaled by 2 and tilted 90° from +y to +x, then (15, 0, 0) is added
    assert np.allclose(out["Post"], [75.0, 0.0, 0.0], atol=1e-4)
    # Pre: the similarity acts inside the instance's frame, the instance's own translation stays (0, 30, 0) + (15, 0, 0)
    assert np.allclose(out["Pre"], [15.0, 30.0, 0.0], atol=1e-4)


def test_transform_application_stamps_every_sdf_onto_every_instance():
    # meta.rs:2012-2075: (sdf, instance) pairs in that order; scaling, rotation, translation nodes only where needed
    T = M.Tagged
    pts = T("Points", {"count": 3})
    grid = T("StratifiedGridTransforms", {"child_id": 0, "shape_x": M._const(3), "shape_y": M._const(1), "shape_z": M._const(1),
                                          "cell_extent_x": M._const(10.0), "cell_extent_y": M._const(10.0),
                                          "cell_extent_z": M._const(10.0), "jitter_fraction": M._const(0.0), "seed": 0})
    nodes = [pts, grid, _spheres(1, 2.0), T("SDFInstantiation", {"child_id": 2}),
             T("TransformApplication", {"sdf_id": 3, "instance_id": 1}), T("SDFGroupUnion", {"child_id": 4, "smoothness": 0.0})]
    g = M.MetaCompiler(nodes, 1.0, 0).build()
    n = g.nodes()
    tr = n[n["kind"] == 3]
    # the middle grid point sits at the origin: no translation node for it (abs_diff_ne), two for the outer ones
    assert len(tr) == 2 and sorted(float(t["p"][0]) for t in tr) == [-10.0, 10.0]
    assert (n["kind"] == 0).sum() == 1 and (n["kind"] == 7).sum() == 2  # one shared sphere, a union tree over 3 stamps
    with pytest.raises(ValueError, match="got Instances"):
        M.MetaCompiler([pts, pts, T("TransformApplication", {"sdf_id": 0, "instance_id": 1})], 1.0, 0).build()


def _surface_scene(subject_kind_node):
    """surface = sphere r = 20 (ids 0-2), subjects = 12 points on a sphere of radius 26 around it (ids 3-4)."""
    T = M.Tagged
    return [
        _spheres(1, 20.0), T("SDFInstantiation", {"child_id": 0}), T("SDFGroupUnion", {"child_id": 1, "smoothness": 0.0}),
        _spheres(12, 1.5, seed=5),
        T("SphereSurfaceTransforms", {"child_id": 3, "radius": M._const(26.0), "jitter_fraction": M._const(0.5),
                                      "rotation": T("Identity", None), "seed": 2}),
        subject_kind_node,
        T("SDFInstantiation", {"child_id": 5}), T("SDFGroupUnion", {"child_id": 6, "smoothness": 0.0}),
        T("SDFUnion", {"child_1_id": 2, "child_2_id": 7, "smoothness": 0.0}),
    ]


@pytest.mark.gpu
def test_closest_translation_lands_the_subjects_on_the_surface(ctx, oracle):
    # meta.rs:1620-1688, 2411-2479: Newton-Raphson on 2x2x2 samples, at most 5 steps, stop within 0.1 of the surface.
    # The batched device probes must give exactly what the reference's per-instance loop gives with the oracle's block
    # evaluation, and the subjects must end up on the sphere.
    T = M.Tagged
    nodes = _surface_scene(T("ClosestTranslationToSurface", {"surface_sdf_id": 2, "subject_id": 4}))
    g = M.MetaCompiler(nodes, 1.0, 0, ctx).build()
    n = g.nodes()
    centres = np.array([t["p"][:3] for t in n[n["kind"] == 3]], np.float32)
    assert len(centres) == 12
    r = np.linalg.norm(centres.astype(np.float64), axis=1)
    assert np.all(np.abs(r - 20.0) < 0.15), r
    # the reference's loop, one instance at a time, on the oracle
    before = M.MetaCompiler(nodes[:5] + [T("SDFInstantiation", {"child_id": 4}), T("SDFGroupUnion", {"child_id": 5, "smoothness": 0.0})],
                            1.0, 0).build().nodes()
    starts = np.array([t["p"][:3] for t in before[before["kind"] == 3]], np.float32)
    surface = M.MetaCompiler(nodes[:3], 1.0, 0).build()
    ogen = oracle.Generator(surface.nodes(), surface.root_node_id)
    f32 = np.float32
    for start, got in zip(starts, centres):
        pos = start.copy()
        for _ in range(5):
            d = ogen.eval_block_preserving_gradients((pos - f32(0.5)).astype(f32), 2)
            total = f32(0.0)
            for q in range(8):
                total = f32(total + d[q])
            sd = f32(total * f32(0.125))
            d000, d001, d010, d011, d100, d101, d110, d111 = [f32(x) for x in d]
            grad = (f32(0.25) * np.array([(d100 + d110 + d101 + d111) - (d000 + d010 + d001 + d011),
                                          (d010 + d110 + d011 + d111) - (d000 + d100 + d001 + d101),
                                          (d001 + d101 + d011 + d111) - (d000 + d100 + d010 + d110)], f32)).astype(f32)
            n2 = f32(f32(grad[0] * grad[0] + grad[1] * grad[1]) + grad[2] * grad[2])
            pos = (pos + (f32(-sd / n2) * grad).astype(f32)).astype(f32)
            if abs(sd) <= f32(0.1):
                break
        want = (start + (pos - start).astype(f32)).astype(f32)
        assert H.f32_bits_equal(got, want).all(), (start, got, want)


@pytest.mark.gpu
def test_rotation_to_gradient_turns_the_y_axis_along_the_gradient(ctx):
    # meta.rs:1798-1862, 2481-2540: the subject's +y ends up along the surface SDF's gradient at the subject's centre —
    # for a sphere around the origin, radially outwards
    T = M.Tagged
    nodes = _surface_scene(T("RotationToGradient", {"gradient_sdf_id": 2, "subject_id": 4}))
    # capsules instead of spheres so that the rotation is visible in the atomic graph
    nodes[3] = T("Capsules", {"segment_length": M._const(4.0), "radius": M._const(1.0), "center_x": M._const(0.0),
                              "center_y": M._const(0.0), "center_z": M._const(0.0), "count": 12, "seed": 5,
                              "sampling": T("OnlyOnce", None)})
    g = M.MetaCompiler(nodes, 1.0, 0, ctx).build()
    n = g.nodes()
    rot = n[n["kind"] == 4]
    tra = n[n["kind"] == 3]
    # where the subjects stood before the node
    before = M.MetaCompiler(nodes[:5] + [T("SDFInstantiation", {"child_id": 4}), T("SDFGroupUnion", {"child_id": 5, "smoothness": 0.0})],
                            1.0, 0).build().nodes()
    t0 = np.array([t["p"][:3] for t in before[before["kind"] == 3]], np.float64)
    assert len(rot) == 12 and len(tra) == 12 and len(t0) == 12
    for q, t, start in zip(rot, tra, t0):
        qq = q["p"][:4].astype(np.float32)
        y = M.quat_rotate(qq, M.v3(0, 1, 0)).astype(np.float64)
        radial = start / np.linalg.norm(start)  # gradient of a sphere's SDF at the subject's centre
        assert np.dot(y, radial) > 0.999, (y, radial)
        # Similarity3::rotated applies the rotation after the transform: the translation turns with it (similarity.rs:138-144)
        moved = M.quat_rotate(qq, start.astype(np.float32)).astype(np.float64)
        assert np.allclose(t["p"][:3], moved, atol=1e-3), (t["p"][:3], moved)
