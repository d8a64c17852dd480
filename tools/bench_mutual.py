"""The reference's `apply_mutual_voxel_absorption` benchmark (engine/src/benchmark/benchmarks/voxel_object.rs:226-268): two
balls of radius R (40 there), the second translated by 1.75 R along x, smoothness 2, unit densities, both inertial
updaters attached. Times ivx_objects_absorb_mutually (wall clock around the synchronous call, fresh objects every
repetition, generated outside the timed region) and the CPU oracle on one core. One JSON line.
    python tools/bench_mutual.py [--radius 40] [--reps 5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import numpy as np

import helpers as H
from impact_b200 import voxel as V
from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--radius", type=float, default=40.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    R = args.radius
    g = H.sphere_graph(R)
    ctx = Context(0)
    gen = SDFVoxelGenerator(1.0, ctx.build_generator(g), H.SAME0)
    dens = np.ones(256, np.float32)
    q, t = np.float32([0, 0, 0, 1]), np.float32([-1.75 * R, 0.0, 0.0])
    ts, stats = [], None
    for rep in range(args.reps + 2):
        a, b = VoxelObject.generate(gen), VoxelObject.generate(gen)
        ia, ib = a.info(), b.info()
        ranges = V.intersection_voxel_ranges(ia["occupied_voxel_ranges"], 1.0, ib["occupied_voxel_ranges"], 1.0, q, t)
        ma, mb = a.inertial_moments(dens).copy(), b.inertial_moments(dens).copy()
        ctx.synchronize()
        t0 = time.perf_counter()
        stats = V.absorb_mutually(a, b, q, t, 2.0, ranges[0], ranges[1], dens, ma, mb)
        dt = (time.perf_counter() - t0) * 1e3
        if rep >= 2:
            ts.append(dt)
    out = {"benchmark": "apply_mutual_voxel_absorption", "radius": R, "grid_shape": list(ia["grid_shape"]),
           "ranges_in_a": np.asarray(ranges[0]).tolist(), "gpu_ms_median": float(np.median(ts)), "gpu_ms_min": float(min(ts)),
           "touched_voxels": [stats[0]["touched_voxels"], stats[1]["touched_voxels"]],
           "emptied_voxels": [stats[0]["emptied_voxels"], stats[1]["emptied_voxels"]],
           "mass_after": [float(ma[0]), float(mb[0])]}
    if args.cpu:
        from oracle import oracle_lib as O

        og = O.VoxelGenerator(O.Generator(g.nodes(), g.root_node_id), 1.0, H.SAME0)
        ca, cb = O.Object.generate(og, 8), O.Object.generate(og, 8)
        mca, mcb = ca.inertial_moments(dens).copy(), cb.inertial_moments(dens).copy()
        t0 = time.perf_counter()
        sc = O.absorb_mutually(ca, cb, q, t, 2.0, ranges[0], ranges[1], dens, mca, mcb)
        out["cpu_port_ms_1_core"] = (time.perf_counter() - t0) * 1e3
        out["cpu_matches_gpu"] = bool(np.array_equal(mca.view(np.uint32), ma.view(np.uint32)) and
                                      sc[0]["emptied_voxels"] == stats[0]["emptied_voxels"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
