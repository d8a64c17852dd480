"""Times ivx_object_inertial_moments on a bench workload (default asteroid1024): wall clock around the synchronous
call (three kernels + a 48-byte read-back), after warm-up. Prints one JSON line.
    python tools/bench_inertia.py [--workload asteroid1024] [--reps 10]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from bench import make_workload
from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="asteroid1024")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    graph, types, desc = make_workload(args.workload)
    ctx = Context(0)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    inf = obj.info()
    dens = [1.0, 2.7, 0.3, 5.5]
    for _ in range(3):
        m = obj.inertial_moments(dens)
    ctx.synchronize()
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        m2 = obj.inertial_moments(dens)
        ts.append((time.perf_counter() - t0) * 1e3)
    assert np.array_equal(m.view(np.uint32), m2.view(np.uint32)), "not deterministic"
    # per-kernel device times (CUDA events inside the library, separate passes so the events do not sit in `ts`)
    ctx.profile_enable(True)
    ctx.profile_reset()
    for _ in range(args.reps):
        obj.inertial_moments(dens)
    prof = {k: round(v[0] / args.reps, 4) for k, v in ctx.profile_get().items() if k.startswith("moments")}
    ctx.profile_enable(False)
    # for_each_surface_voxel over the whole object (ivx_object_surface_voxels_in_ranges), wall clock incl. the copy out
    tq = []
    for _ in range(4):
        t0 = time.perf_counter()
        sv = obj.surface_voxels_in_ranges()
        tq.append((time.perf_counter() - t0) * 1e3)
    n_chunks = int(np.prod(inf["chunk_counts"]))
    ms = float(np.median(ts))
    # algorithmic bytes: 2 B (type + flags) per voxel of every non-uniform chunk + 16 B descriptor per chunk
    bytes_alg = inf["n_non_uniform"] * 4096 * 2 + n_chunks * 16
    print(json.dumps({"workload": args.workload, "grid_shape": list(inf["grid_shape"]), "chunks": n_chunks,
                      "non_uniform": inf["n_non_uniform"], "uniform": inf["n_uniform"], "ms_median": ms,
                      "ms_min": float(min(ts)), "algorithmic_GB": bytes_alg / 1e9,
                      "GBps": bytes_alg / ms / 1e6, "kernel_ms": prof,
                      "surface_voxels": int(len(sv)), "surface_query_ms_min": float(min(tq)), "moments": [float(x) for x in m]}))


if __name__ == "__main__":
    main()
