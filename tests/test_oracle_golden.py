"""The oracle against every golden vector the reference holds for this path, plus a
quantisation table built independently of the C++ oracle. CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    with open(os.path.join(HERE, "golden", "surface_nets_materials.json")) as f:
        return json.load(f)


def test_vertex_materials_match_reference_known_answers(oracle):
    # surface_nets.rs:680-757
    for case in _golden()["vertex_materials"]:
        ind, w = oracle.vertex_materials(case["has_voxel"], case["materials"])
        n = ind[7]
        assert n == len(case["indices"]), case
        assert list(ind[:n]) == case["indices"], case
        assert list(w[:n]) == case["weights"], case
        assert not w[n:].any()


def test_triangle_index_materials_match_reference_known_answers(oracle):
    # surface_nets.rs:759-877; inputs built like `with_valid_indices_and_weights` (surface_nets.rs:497-510)
    for case in _golden()["triangle_index_materials"]:
        vms = []
        for v in case["vertices"]:
            ind = np.zeros(8, np.uint8)
            w = np.zeros(8, np.uint8)
            ind[: len(v["indices"])] = v["indices"]
            w[: len(v["weights"])] = v["weights"]
            ind[7] = len(v["indices"])
            vms.append((ind, w))
        out = oracle.index_materials(vms)
        for (ind, w), exp in zip(out, case["expected"]):
            assert list(ind) == exp["indices"], case
            assert list(w) == exp["weights"], case


def test_quantisation_matches_independent_float32_restatement(oracle):
    # lib.rs:154-201: INVERSE_QUANTIZATION_STEP_SIZE = 1.0 / 0.02 (== 50.0 in f32), `as i8` saturates
    # and truncates toward zero, NaN → 0; decode = code as f32 * 0.02.
    inv = np.float32(1.0) / np.float32(0.02)
    assert inv == np.float32(50.0)
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.uniform(-3.5, 3.5, 20000).astype(np.float32),
        (np.arange(-140, 141, dtype=np.float32) * np.float32(0.02)),
        np.nextafter(np.arange(-140, 141, dtype=np.float32) * np.float32(0.02), np.float32(10)),
        np.nextafter(np.arange(-140, 141, dtype=np.float32) * np.float32(0.02), np.float32(-10)),
        np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 2.54, -2.56, 1e30, -1e30], np.float32),
    ])
    with np.errstate(invalid="ignore"):
        scaled = vals * inv
        expected = np.where(np.isnan(scaled), 0, np.clip(np.trunc(scaled), -128, 127)).astype(np.int8)
    got = np.array([oracle.sd_encode(float(v)) for v in vals], np.int8)
    assert np.array_equal(got, expected)
    # MAX_F32 / MIN_F32 / VOID_LIMIT (lib.rs:158-162)
    assert np.float32(oracle.lib().orc_sd_decode(127)) == np.float32(0.02) * np.float32(127)
    assert np.float32(oracle.lib().orc_sd_decode(-128)) == np.float32(0.02) * np.float32(-128)


def test_decode_reencode_is_not_idempotent_like_the_reference(oracle):
    # SURVEY §7 hard part 3: trunc((e * 0.02f) * 50f) != e for a subset of codes; absorption relies on it
    codes = np.arange(-128, 128, dtype=np.int32)
    re = np.array([oracle.sd_encode(float(np.float32(c) * np.float32(0.02))) for c in codes])
    with np.errstate(invalid="ignore"):
        expected = np.trunc((codes.astype(np.float32) * np.float32(0.02)) * np.float32(50.0)).astype(np.int32)
    assert np.array_equal(re, expected)
    assert (re != codes).sum() > 0


@pytest.mark.parametrize("name", ["sphere64", "zoo", "asteroid_like", "mid_noise", "noisy_box"])
def test_oracle_objects_match_the_committed_digests(oracle, name):
    # tests/golden/object_digests.json (written by tests/golden/make_object_digests.py): the restatement's output for
    # these graphs is frozen; the GPU tests compare the CUDA path with the same digests
    import json
    import os

    import helpers as H

    with open(os.path.join(os.path.dirname(__file__), "golden", "object_digests.json")) as f:
        want = json.load(f)[name]
    make, types_name = H.GOLDEN_OBJECTS[name]
    g = make()
    gen = oracle.Generator(g.nodes(), g.root_node_id)
    obj = oracle.Object.generate(oracle.VoxelGenerator(gen, 1.0, getattr(H, types_name)), 4)
    m = obj.mesh(2)
    assert [int(x) for x in obj.info()["chunk_counts"]] == want["chunk_counts"]
    assert H.object_digest(obj.chunks(), obj.voxels()) == want["object"]
    assert (m.n_vertices, m.n_indices) == (want["vertices"], want["indices"])
    assert H.mesh_digest(m.positions, m.normals, m.indices, m.index_materials, m.submeshes, m.vertex_ranges) == want["mesh"]

