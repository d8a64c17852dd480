import sys, os
sys.path.insert(0, os.getcwd())
from impact_b200 import workloads as W
from impact_b200.voxel import *
import bench
g, t, d = bench.make_workload(sys.argv[1] if len(sys.argv)>1 else "asteroid1024")
ctx = Context(0)
gen = ctx.build_generator(g)
obj = VoxelObject.generate(SDFVoxelGenerator(1.0, gen, t))
print(obj.info())
