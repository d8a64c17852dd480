"""Per-kernel device times of ivx_object_generate vs ivx_object_generate_streamed (CUDA events inside the library)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from impact_b200.voxel import *
import bench
g, t, d = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "asteroid1024")
ctx = Context(0)
vg = SDFVoxelGenerator(1.0, ctx.build_generator(g), t)
obj = VoxelObject.generate(vg)
info = obj.info()
n = int(np.prod(info["chunk_counts"]))
h_chunks = torch.empty(n * 16, dtype=torch.uint8, pin_memory=True).numpy()
h_vox = torch.empty((info["n_non_uniform"] + 64) * 4096 * 3, dtype=torch.uint8, pin_memory=True).numpy()
obj.free()
for mode in ("plain", "streamed", "plain", "streamed"):
    ctx.profile_enable(True); ctx.profile_reset(); ctx.synchronize()
    t0 = time.perf_counter()
    if mode == "plain":
        o = VoxelObject.generate(vg)
    else:
        o, nnu = VoxelObject.generate_streamed(vg, h_chunks, h_vox)
    t1 = time.perf_counter()
    ctx.synchronize()
    t2 = time.perf_counter()
    prof = ctx.profile_get()
    print(mode, "call ms", round(1e3 * (t1 - t0), 2), "+sync", round(1e3 * (t2 - t1), 2),
          {k: (round(v[0], 2), v[1]) for k, v in prof.items() if v[1]}, "sum", round(sum(v[0] for v in prof.values()), 2))
    o.free()
