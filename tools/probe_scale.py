#!/usr/bin/env python
"""Beyond BASELINE's largest configuration: the asteroid graph compiled for a grid of up to `hi`^3 voxels (default 2048),
generated whole on one GPU and as 4 x-slabs with the halo protocol; both must give the same chunk planes and the same
mesh (there is no oracle digest at this size — the slab/whole identity is the size-independent property), and the step
time is reported next to the 1024^3 figure.

    python tools/probe_scale.py [hi] [steps] [time-only]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from impact_b200 import digests as DG
    from impact_b200 import distributed as D
    from impact_b200 import meta
    from impact_b200 import workloads as W
    from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject, VoxelObjectMesh, plane_work

    hi = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ctx = Context(0)
    t0 = time.perf_counter()
    graph = meta.asteroid_graph_scaled(hi - 32, hi, 0, ctx, use_cache=False)
    print(f"compiled for ({hi - 32},{hi}]: {len(graph)} atomic nodes, scale_factor {graph.scale_factor:.4f}, "
          f"{time.perf_counter() - t0:.2f} s", flush=True)
    vg = SDFVoxelGenerator(1.0, ctx.build_generator(graph), W.gradient_noise_types())

    def mesh_digest(m):
        return DG.mesh_digest(m["positions"], m["normals"], m["indices"], m["index_materials"], m["submeshes"],
                              m["vertex_ranges"])

    # timing: generate + mesh, free between steps (the same step bench.py times)
    ms = []
    for s in range(3 + steps):
        ctx.synchronize()
        t0 = time.perf_counter()
        obj = VoxelObject.generate(vg)
        mesh = VoxelObjectMesh.create(obj)
        ctx.synchronize()
        if s >= 3:
            ms.append(1e3 * (time.perf_counter() - t0))
        if s < 3 + steps - 1:
            obj.free()
    info = obj.info()
    n_vox = int(np.prod(info["grid_shape"]))
    print(f"grid {tuple(info['grid_shape'])}  chunks void/uniform/non-uniform {info['n_void']}/{info['n_uniform']}/"
          f"{info['n_non_uniform']}  mesh {mesh.n_vertices} vertices {mesh.n_indices} indices", flush=True)
    print(f"step (generate + mesh, device resident): median {np.median(ms):.2f} ms  min {min(ms):.2f}  "
          f"-> {n_vox / np.median(ms) / 1e6:.1f} G voxels/s", flush=True)
    print(f"device memory in use: {torch.cuda.mem_get_info(0)[1] / 2**30 - torch.cuda.mem_get_info(0)[0] / 2**30:.1f} GiB",
          flush=True)

    if len(sys.argv) > 3 and sys.argv[3] == "time-only":
        return
    chunks, voxels = obj.download()
    whole_planes = DG.object_plane_digests(chunks, voxels, info["chunk_counts"])
    del chunks, voxels
    whole_mesh = mesh_digest(mesh.download())
    whole_counts = (mesh.n_vertices, mesh.n_indices)
    obj.free()

    ranges = D.slab_ranges_weighted(plane_work(vg), 4)
    slabs = [VoxelObject.generate(vg, r) for r in ranges]
    D.exchange_halos_single_process(slabs)
    got = []
    for s in slabs:
        c, v = s.download()
        got += DG.object_plane_digests(c, v, s.info()["chunk_counts"])
    bad = [p for p, (a, b) in enumerate(zip(got, whole_planes)) if a != b]
    merged = D.merge_mesh_parts([VoxelObjectMesh.create(s).download() for s in slabs])
    ok_mesh = mesh_digest(merged) == whole_mesh and (len(merged["positions"]), len(merged["indices"])) == whole_counts
    print(f"4 x-slabs {ranges}: planes differing from the whole object: {bad or 'none'}; merged mesh identical: {ok_mesh}")
    if bad or not ok_mesh:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
