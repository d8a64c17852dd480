#!/usr/bin/env python
"""Wall time of one absorption call on the 1024^3 asteroid and what it is made of (launch list via IVX profile)."""
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
    from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

    graph, types, _ = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "asteroid1024")
    ctx = Context(0)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), types))
    VoxelObjectMesh.create(obj)
    centers, radius = absorber_path(obj.info()["grid_shape"], 14)
    for s, c in enumerate(centers):
        ctx.synchronize()
        l0 = ctx.kernel_launch_count
        t0 = time.perf_counter()
        st = obj.absorb_sphere(c, radius, radius + 2.0)
        t1 = time.perf_counter()
        d = obj.invalidated_mesh_chunk_indices()
        t2 = time.perf_counter()
        VoxelObjectMesh.sync(obj)
        t3 = time.perf_counter()
        print(f"step {s}: absorb {1e3*(t1-t0):.3f} ms ({ctx.kernel_launch_count - l0} launches)  dirty query {1e3*(t2-t1):.3f} ms  "
              f"mesh sync {1e3*(t3-t2):.3f} ms  touched {st['touched_chunks']} dirty {len(d)}")


if __name__ == "__main__":
    main()
