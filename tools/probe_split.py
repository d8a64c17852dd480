#!/usr/bin/env python
"""Where the time of one connected-region resolve goes (asteroid1024): wall, device and host parts, for a fresh object
(every chunk labelled), an unchanged one, and after each of a few absorption steps."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import bench
    from bench_fracture import absorber_path
    from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject

    wl = sys.argv[1] if len(sys.argv) > 1 else "asteroid1024"
    graph, types, _ = bench.make_workload(wl)
    ctx = Context(0)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    shape = obj.info()["grid_shape"]
    centers, radius = absorber_path(shape, 12)
    rows = []

    def resolve(tag):
        ctx.synchronize()
        t0 = time.perf_counter()
        r = obj.resolve_connected_regions()
        t1 = time.perf_counter()
        rows.append({"what": tag, "wall_ms": round(1e3 * (t1 - t0), 3), "device_ms": round(r["device_ms"], 3),
                     "host_ms": round(r["host_ms"], 3), "regions": r["n_regions"], "local_regions": r["n_local_regions"],
                     "connections": r["n_connections"], "relabelled": r["n_relabelled_chunks"]})

    resolve("fresh")
    resolve("unchanged")
    resolve("unchanged")
    for s, c in enumerate(centers):
        obj.absorb_sphere(c, radius, radius + 2.0)
        resolve(f"step{s}")
        while rows[-1]["regions"] > 1:
            t0 = time.perf_counter()
            xi, frag = obj.extract_any_disconnected_region()
            ctx.synchronize()
            rows.append({"what": f"step{s} extract", "wall_ms": round(1e3 * (time.perf_counter() - t0), 3),
                         "extracted": xi["extracted"], "regions_before": xi["n_regions_before"]})
            resolve(f"step{s} again")
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
