import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from impact_b200 import workloads as W
from impact_b200.voxel import *
import bench
t0=time.time()
g, t, d = bench.make_workload(sys.argv[1] if len(sys.argv)>1 else "asteroid1024")
print("workload", d, "graph nodes", len(g), "build s", time.time()-t0)
ctx = Context(0)
gen = ctx.build_generator(g)
print("program nodes", gen.node_count, "depth", gen.stack_depth)
obj = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, t))
print(obj.info())
if len(sys.argv) > 2 and sys.argv[2] == "mesh":
    m = VoxelObjectMesh.create(obj)
    print("mesh", m.info if hasattr(m, "info") else "")
