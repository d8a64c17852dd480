"""Parity at the BASELINE.json configurations' own sizes: the CUDA path through the C ABI against the
committed digests of the CPU oracle's output (tests/golden/baseline_digests.json, written by
tests/golden/make_baseline_digests.py) — chunk table, every voxel, and the mesh buffers in the reference's order, for
config 1 (64³ sphere, + the 202³ engine-bench sphere), config 2 (256³ noisy box), config 3 (asteroid ≤ 512³), config 4
(asteroid ≤ 1024³: whole object on one GPU, and as x-slabs with the halo protocol; plus the same recipe one size beyond
BASELINE, ≤ 2048³) and config 5 (32 absorption steps on the config-4 object: dirty set of every step, final object,
final mesh, dirty remesh patches against a full re-mesh).
The oracle is not in the loop here: the digests are the fixture."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from impact_b200 import digests as DG  # noqa: E402
from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh  # noqa: E402

pytestmark = pytest.mark.gpu

with open(os.path.join(os.path.dirname(__file__), "golden", "baseline_digests.json")) as f:
    WANT = json.load(f)


def _generate(ctx, name, chunk_i_range=None):
    graph, types, _ = bench.make_workload(name)
    vg = SDFVoxelGenerator(1.0, ctx.build_generator(graph), types)
    return VoxelObject.generate(vg, chunk_i_range), vg


def _mesh_digest(m):
    return DG.mesh_digest(m["positions"], m["normals"], m["indices"], m["index_materials"], m["submeshes"],
                          m["vertex_ranges"])


@pytest.mark.parametrize("name", ["sphere64", "sphere202", "noisybox256", "asteroid512", "asteroid1024", "asteroid2048"])
def test_object_and_mesh_match_the_oracle_digests_at_full_size(ctx, name):
    # asteroid2048 (1930 x 1685 x 2039 voxels, 8x the non-uniform chunks of config 4) is beyond BASELINE's largest
    # configuration: the same recipe one size up, to pin the 32-bit offsets and the capacities of a 4 GB object
    want = WANT[name]
    obj, _ = _generate(ctx, name)
    info = obj.info()
    assert list(info["grid_shape"]) == want["grid_shape"] and list(info["chunk_counts"]) == want["chunk_counts"]
    assert {"void": info["n_void"], "uniform": info["n_uniform"], "non_uniform": info["n_non_uniform"]} == want["chunks"]
    chunks, voxels = obj.download()
    planes = DG.object_plane_digests(chunks, voxels, info["chunk_counts"])
    bad = [p for p, (a, b) in enumerate(zip(planes, want["object_planes"])) if a != b]
    assert not bad, f"chunk planes {bad} differ from the oracle"
    assert DG.combine(planes) == want["object"]
    del chunks, voxels
    mesh = VoxelObjectMesh.create(obj)
    assert (mesh.n_vertices, mesh.n_indices, mesh.n_submeshes) == (want["vertices"], want["indices"], want["submeshes"])
    assert _mesh_digest(mesh.download()) == want["mesh"]
    obj.free()


def test_streamed_generation_matches_at_full_size(ctx):
    # ivx_object_generate_streamed (what bench.py's e2e leg calls): same object, straight into host buffers
    from impact_b200 import _lib as L

    want = WANT["asteroid1024"]
    graph, types, _ = bench.make_workload("asteroid1024")
    vg = SDFVoxelGenerator(1.0, ctx.build_generator(graph), types)
    n_chunks = int(np.prod(want["chunk_counts"]))
    chunks = np.zeros(n_chunks, L.CHUNK_DTYPE)
    voxels = np.zeros((want["chunks"]["non_uniform"] + 16) * 4096, L.VOXEL_DTYPE)
    obj, nnu = VoxelObject.generate_streamed(vg, chunks, voxels)
    ctx.synchronize()
    assert nnu == want["chunks"]["non_uniform"]
    planes = DG.object_plane_digests(chunks, voxels, want["chunk_counts"])
    assert planes == want["object_planes"]
    obj.free()


@pytest.mark.parametrize("n_slabs", [2, 8])
def test_config4_x_slabs_match_plane_by_plane(ctx, n_slabs):
    # config 4 in one process: every slab generated on its own, halos exchanged through the slab protocol
    # (device buffers), each slab's planes and the concatenated slab meshes against the whole-object digests
    from impact_b200 import distributed as D
    from impact_b200.voxel import plane_work

    want = WANT["asteroid1024"]
    graph, types, _ = bench.make_workload("asteroid1024")
    vg = SDFVoxelGenerator(1.0, ctx.build_generator(graph), types)
    ranges = D.slab_ranges_weighted(plane_work(vg), n_slabs)
    assert ranges[0][0] == 0 and ranges[-1][1] == want["chunk_counts"][0]
    slabs = [VoxelObject.generate(vg, r) for r in ranges]
    D.exchange_halos_single_process(slabs)
    got = []
    for s in slabs:
        info = s.info()
        chunks, voxels = s.download()
        got += DG.object_plane_digests(chunks, voxels, info["chunk_counts"])
    bad = [p for p, (a, b) in enumerate(zip(got, want["object_planes"])) if a != b]
    assert not bad, f"chunk planes {bad} differ from the oracle"
    parts = [VoxelObjectMesh.create(s).download() for s in slabs]
    merged = D.merge_mesh_parts(parts)
    assert (len(merged["positions"]), len(merged["indices"])) == (want["vertices"], want["indices"])
    assert _mesh_digest(merged) == want["mesh"]
    for s in slabs:
        s.free()


def test_config5_fracture_sequence_matches_step_by_step(ctx):
    want = WANT["asteroid1024"]["fracture"]
    obj, _ = _generate(ctx, "asteroid1024")
    VoxelObjectMesh.create(obj)
    for step, (c, w) in enumerate(zip(want["centers"], want["per_step"])):
        st = obj.absorb_sphere(np.float32(c), want["absorber_radius"], want["absorber_radius"] + 2.0)
        dirty = np.sort(obj.invalidated_mesh_chunk_indices().astype(np.uint32))
        assert len(dirty) == w["dirty_chunks"], f"step {step}"
        assert hashlib.sha256(dirty.tobytes()).hexdigest() == w["dirty"], f"step {step}: dirty set differs"
        for k, v in w["stats"].items():
            assert st[k] == v, f"step {step}: {k}"
        VoxelObjectMesh.sync_with_voxel_object(obj)  # dirty-chunk remesh, clears the dirty set
    info = obj.info()
    chunks, voxels = obj.download()
    planes = DG.object_plane_digests(chunks, voxels, info["chunk_counts"])
    bad = [p for p, (a, b) in enumerate(zip(planes, want["object_planes"])) if a != b]
    assert not bad, f"chunk planes {bad} differ from the oracle after the sequence"
    del chunks, voxels
    mesh = VoxelObjectMesh.create(obj)
    assert (mesh.n_vertices, mesh.n_indices) == (want["vertices"], want["indices"])
    assert _mesh_digest(mesh.download()) == want["mesh"]
    obj.free()


def test_config5_split_handling_matches_the_oracle_at_512(ctx, oracle):
    """Config 5's split handler at the size of config 3 (472 x 412 x 498, ~10^4 local regions), live against the oracle: after
    every absorption the roots of all local regions, the two reported regions and the extraction that follows."""
    import bench
    from test_gpu_split_detection import assert_split_equal
    graph, types, _ = bench.make_workload("asteroid512")
    obj_cpu = oracle.Object.generate(oracle.VoxelGenerator(oracle.Generator(graph.nodes(), graph.root_node_id), 1.0, types), 16)
    obj_gpu = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    shape = np.array(obj_cpu.info()["chunk_counts"]) * 16
    R = 0.5 * float(shape.max())
    radius = np.float32(0.15 * R)
    start = (0.5 * shape - R / np.sqrt(3.0)).astype(np.float32)
    d = np.float32(1.0 / np.sqrt(3.0))
    pieces = 0
    for step in range(6):
        c = (start + np.float32(step) * radius * d).astype(np.float32)
        obj_cpu.absorb_sphere(c, float(radius), float(radius + 2.0))
        obj_gpu.absorb_sphere(c, float(radius), float(radius + 2.0))
        g = assert_split_equal(obj_gpu, obj_cpu)
        for _ in range(4):  # split pieces off while there are any, piece by piece like the engine
            if g["n_regions"] < 2:
                break
            ic, ec = obj_cpu.extract_any_disconnected_region()
            ig, eg = obj_gpu.extract_any_disconnected_region()
            for k in ("found_two", "extracted", "discarded", "single_chunk"):
                assert bool(ig[k]) == bool(ic[k]), (step, k)
            pieces += 1
            g = assert_split_equal(obj_gpu, obj_cpu)
    assert g["n_local_regions"] > 5000
