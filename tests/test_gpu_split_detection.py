"""GPU connected-region detection (ivx_object_resolve_connected_regions) against the CPU oracle: local region
labels per voxel, region counts per chunk, the resolved root of every local region, count_regions,
find_two_disconnected_regions and the smallest-region choice — all bit-exact — plus the reference's own
invariant (region count == brute-force flood fill, split_detection.rs:490-560) at sizes the oracle skips."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import workloads as W
from impact_b200.voxel import SDFVoxelGenerator, VoxelObject
from test_oracle_split_detection import two_spheres_graph

pytestmark = pytest.mark.gpu


def _both(ctx, oracle, graph, types=H.SAME0):
    obj_cpu = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(graph.nodes(), graph.root_node_id), 1.0, types), 4)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    return obj_gpu, obj_cpu


def assert_split_equal(obj_gpu, obj_cpu):
    g = obj_gpu.resolve_connected_regions(download=True)
    c = obj_cpu.split_detection()
    assert not c["overflow"]
    assert g["n_regions"] == c["n_regions"] and g["has_two"] == c["has_two"]
    assert np.array_equal(g["per_chunk"]["region_count"], c["per_chunk"]["region_count"])
    assert np.array_equal(g["per_chunk"]["boundary_region_count"], c["per_chunk"]["boundary_region_count"])
    assert np.array_equal(g["per_chunk"]["first_region"], c["per_chunk"]["first_region"])
    # the library hands the labels out in linear chunk order, the reference keeps them by data offset (the same order
    # until a Uniform chunk is converted and takes a slot at the end)
    ch = obj_cpu.chunks()
    offsets = ch["data_offset"][ch["kind"] == 2]
    want = c["voxel_labels"].reshape(-1, 4096)[offsets]
    assert np.array_equal(g["voxel_labels"].reshape(-1, 4096), want), "local region labels differ"
    assert np.array_equal(g["region_roots"], c["region_roots"]), "resolved roots differ"
    if c["has_two"]:
        assert g["two"] == c["two"] and g["smallest"] == c["smallest"]
        for q in range(2):
            for f in ("chunk_count", "non_uniform_chunk_count"):
                assert g["candidates"][q][f] == c["candidates"][q][f]
            assert np.array_equal(g["candidates"][q]["chunk_min"], c["candidates"][q]["chunk_min"])
            assert np.array_equal(g["candidates"][q]["chunk_max"], c["candidates"][q]["chunk_max"])
    return g


CASES = {
    "single_voxel": (lambda: H.box_graph(1.0), H.SAME0, 1),          # connected_region_count_is_correct_for_single_voxel
    "sphere64": (lambda: H.sphere_graph(31.0), H.SAME0, 1),
    "two_spheres": (lambda: two_spheres_graph(), H.SAME0, 2),        # should_split_off_disconnected_sphere
    "two_spheres_unequal": (lambda: two_spheres_graph(60.0, 25.0, 9.0), H.GRADIENT4, 2),
    "complex": (lambda: H.complex_graph(0.5), H.SAME0, 1),
    "noisy_debris": (lambda: H.noisy_sphere_graph(20.0, 4), H.SAME0, None),
    "noisy_box": (lambda: H.noisy_box_graph(38.0, 8), H.SAME0, None),
    "asteroid_like": (lambda: H.asteroid_like_graph(24, 40.0), H.GRADIENT4, None),
    "asteroid_stand_in": (lambda: W.asteroid_stand_in(0.45, seed=2), H.SAME0, None),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_connected_regions_match_the_oracle(ctx, oracle, name):
    make, types, expected = CASES[name]
    obj_gpu, obj_cpu = _both(ctx, oracle, make(), types)
    g = assert_split_equal(obj_gpu, obj_cpu)
    assert g["n_regions"] == obj_cpu.count_regions_brute_force()
    if expected is not None:
        assert g["n_regions"] == expected


def test_capacities_that_start_too_small_are_found_by_retrying(ctx, oracle, monkeypatch):
    # the per-region arrays, the connection records and the tree-pair table are sized from the previous resolve; a
    # kernel that needs more reports it and the pass runs again. IVX_REGIONS_TINY_CAPACITIES starts all of them at
    # almost nothing: same result, region for region.
    monkeypatch.setenv("IVX_REGIONS_TINY_CAPACITIES", "1")
    make, types, expected = CASES["noisy_debris"]
    obj_gpu, obj_cpu = _both(ctx, oracle, make(), types)
    assert_split_equal(obj_gpu, obj_cpu)
    assert_split_equal(obj_gpu, obj_cpu)  # and again, with the object's own capacities in place


def test_the_replay_forest_in_global_memory_gives_the_same_roots(ctx, oracle, monkeypatch):
    # the disjoint-set forest of the replay lives in shared memory while the trees fit (~5 x 10^4); beyond that it is an
    # array in global memory. IVX_REGIONS_GLOBAL_FOREST takes that path for any object.
    monkeypatch.setenv("IVX_REGIONS_GLOBAL_FOREST", "1")
    for name in ("noisy_debris", "two_spheres_unequal"):
        make, types, expected = CASES[name]
        obj_gpu, obj_cpu = _both(ctx, oracle, make(), types)
        assert_split_equal(obj_gpu, obj_cpu)


def test_absorption_splits_a_dumbbell(ctx, oracle):
    from impact_b200.graph import SDFGraph

    g = SDFGraph()
    a = g.sphere(14.0)
    b = g.translation(g.sphere(14.0), [44.0, 0.0, 0.0])
    bridge = g.capsule(30.0, 3.0)
    bridge = g.rotation_from_axis_angle(bridge, [0.0, 0.0, 1.0], float(np.pi / 2))
    bridge = g.translation(bridge, [22.0, 0.0, 0.0])
    g.union(g.union(a, b, 1.0), bridge, 1.0)
    obj_gpu, obj_cpu = _both(ctx, oracle, g)
    assert assert_split_equal(obj_gpu, obj_cpu)["n_regions"] == 1
    shape = np.array(obj_cpu.info()["chunk_counts"]) * 16
    center = (0.5 * shape).astype(np.float32)
    for step in range(3):
        obj_cpu.absorb_sphere(center, 7.0, 9.0)
        obj_gpu.absorb_sphere(center, 7.0, 9.0)
        assert_split_equal(obj_gpu, obj_cpu)
    assert obj_gpu.count_regions() == obj_cpu.count_regions_brute_force() == 2


def test_fracturing_sequence_keeps_matching(ctx, oracle):
    # config 5 geometry on a small asteroid-like object: absorb along the diagonal, re-resolve every step
    obj_gpu, obj_cpu = _both(ctx, oracle, H.asteroid_like_graph(16, 36.0), H.GRADIENT4)
    shape = np.array(obj_cpu.info()["chunk_counts"]) * 16
    R = 0.5 * float(shape.max())
    radius = np.float32(0.15 * R)
    center = (0.5 * shape - R / np.sqrt(3.0)).astype(np.float32)
    for step in range(5):
        c = (center + step * radius * np.float32(0.6)).astype(np.float32)
        obj_cpu.absorb_sphere(c, float(radius), float(radius + 2.0))
        obj_gpu.absorb_sphere(c, float(radius), float(radius + 2.0))
        assert_split_equal(obj_gpu, obj_cpu)


@pytest.mark.parametrize("seed,fill", [(11, 0.8), (12, 1.0), (13, 1.3), (16, 0.9), (17, 0.75)])
def test_ragged_random_grids_match_the_oracle(ctx, oracle, seed, fill):
    # white-noise emptiness: 60-90 % of the voxels present, so a chunk has dozens to hundreds of voxels without a lower
    # neighbour (trees of the union pass) and up to thousands of links between them — the short event list of
    # k_local_regions overflows for the sparser grids (the run-by-run fallback) and is deduplicated for the denser ones
    vox, sp, grid = H.random_voxel_chunks((48, 40, 56), seed, fill=fill, blobs=False)
    obj_gpu = VoxelObject.from_generated_chunks(ctx, 0.5, grid, vox, sp)
    obj_cpu = oracle.Object.from_generated_chunks(vox, sp, grid, 0.5)
    if obj_cpu.split_detection()["overflow"]:
        pytest.skip("more than 254 regions in a chunk: the reference asserts")
    g = assert_split_equal(obj_gpu, obj_cpu)
    assert g["n_regions"] == obj_cpu.count_regions_brute_force()
    mid = np.float32([24.0, 20.0, 28.0])
    for r in (7.0, 11.0):
        obj_cpu.absorb_sphere(mid, r, r + 2.0)
        obj_gpu.absorb_sphere(mid, r, r + 2.0)
        assert_split_equal(obj_gpu, obj_cpu)


def test_region_count_of_a_large_sphere(ctx):
    # engine bench shape Sphere(r = 100) → 202³: one region; no oracle needed
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(H.sphere_graph(100.0)), H.SAME0))
    r = obj.resolve_connected_regions()
    assert r["n_regions"] == 1 and not r["has_two"]
