"""Writes tests/golden/baseline_digests.json: sha256 digests of what the CPU oracle produces for the BASELINE.json
configurations AT THEIR STATED SIZES: config 1 sphere 64³ (+ the engine-bench 202³ sphere), config 2
noisy box 256³, config 3 asteroid ≤ 512³, config 4 asteroid ≤ 1024³ (per chunk plane, so x-slab objects of any
partition can be checked), config 5 the 32-step absorption sequence on the config-4 object (dirty sets per step, final
object, final mesh).

    python tests/golden/make_baseline_digests.py [workload ...]        # ≈ 6 min on 8 cores for everything
    python tests/golden/make_baseline_digests.py asteroid2048          # one size beyond BASELINE; ≈ 11 min on 8 cores

The `-m gpu` tests (tests/test_gpu_baseline_sizes.py) and bench.py's `parity` block compare the CUDA path against
these committed digests without the oracle in the loop."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from impact_b200 import digests as DG  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402

WORKLOADS = ["sphere64", "sphere202", "noisybox256", "asteroid512", "asteroid1024"]
FRACTURE_STEPS = 32
OUT = os.path.join(HERE, "baseline_digests.json")


def absorber_path(shape, steps):
    """BASELINE config 5 (benchmarks/voxel_object.rs:343-362): radius 0.15 R, start on the bounding sphere along the
    (1,1,1) diagonal, one absorber radius inward per step."""
    R = 0.5 * float(max(shape))
    radius = np.float32(0.15 * R)
    start = (0.5 * np.asarray(shape, np.float64) - R / np.sqrt(3.0)).astype(np.float32)
    d = np.float32(1.0 / np.sqrt(3.0))
    return [(start + np.float32(s) * radius * d).astype(np.float32) for s in range(steps)], float(radius)


def digest_object(obj):
    info = obj.info()
    planes = DG.object_plane_digests(obj.chunks(), obj.voxels(), info["chunk_counts"])
    return planes


def digest_mesh(m):
    return DG.mesh_digest(m.positions, m.normals, m.indices, m.index_materials, m.submeshes, m.vertex_ranges)


def main():
    names = sys.argv[1:] or WORKLOADS
    out = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            out = json.load(f)
    threads = os.cpu_count() or 1
    for name in names:
        graph, types, desc = bench.make_workload(name)
        t0 = time.perf_counter()
        gen = O.Generator(graph.nodes(), graph.root_node_id)
        vg = O.VoxelGenerator(gen, 1.0, types)
        obj = O.Object.generate(vg, threads)
        m = obj.mesh(threads)
        info = dict(obj.info(), grid_shape=tuple(int(x) for x in vg.grid_shape))
        ch = obj.chunks()
        planes = digest_object(obj)
        entry = {
            "description": desc, "grid_shape": [int(x) for x in info["grid_shape"]],
            "chunk_counts": [int(x) for x in info["chunk_counts"]],
            "chunks": {"void": int((ch["kind"] == 0).sum()), "uniform": int((ch["kind"] == 1).sum()),
                       "non_uniform": int((ch["kind"] == 2).sum())},
            "object_planes": planes, "object": DG.combine(planes),
            "mesh": digest_mesh(m), "vertices": int(m.n_vertices), "indices": int(m.n_indices),
            "submeshes": int(m.n_submeshes),
        }
        print(f"{name}: grid {entry['grid_shape']} {entry['chunks']} V={entry['vertices']} I={entry['indices']} "
              f"({time.perf_counter() - t0:.1f}s)", flush=True)
        if name == "asteroid1024":
            # config 5 on the same object
            del m
            obj.clear_dirty()
            centers, radius = absorber_path(info["grid_shape"], FRACTURE_STEPS)
            steps = []
            t0 = time.perf_counter()
            for c in centers:
                st = obj.absorb_sphere(c, radius, radius + 2.0)
                d = np.sort(np.asarray(obj.dirty(), np.uint32).reshape(-1))
                steps.append({"dirty_chunks": int(len(d)), "dirty": hashlib.sha256(d.tobytes()).hexdigest(),
                              "stats": {k: int(v) for k, v in st.items()} if isinstance(st, dict) else None})
                obj.clear_dirty()
            planes5 = digest_object(obj)
            m5 = obj.mesh(threads)
            entry["fracture"] = {
                "steps": FRACTURE_STEPS, "absorber_radius": radius, "centers": [[float(x) for x in c] for c in centers],
                "per_step": steps, "object_planes": planes5, "object": DG.combine(planes5), "mesh": digest_mesh(m5),
                "vertices": int(m5.n_vertices), "indices": int(m5.n_indices),
            }
            print(f"  fracture: {FRACTURE_STEPS} steps, final V={m5.n_vertices} ({time.perf_counter() - t0:.1f}s)", flush=True)
        out[name] = entry
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
