import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from impact_b200.voxel import *
from impact_b200 import distributed as D
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); lr=int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); torch.cuda.set_device(lr)
stream=torch.cuda.Stream(); ctx=Context(lr, stream=stream.cuda_stream)
g,t,d=bench.make_workload("asteroid1024"); gen=ctx.build_generator(g); vg=SDFVoxelGenerator(1.0, gen, t)
_,_,dlo,dhi=compile_program_host(g)
gs=[int(np.ceil(np.float32(h)-np.float32(l)))+2 for l,h in zip(dlo,dhi)]
planes=(gs[0]+15)//16; ranges=D.slab_ranges(planes, world); slab=ranges[rank]; dev=torch.device("cuda", lr)
import types
T={}
def timed(name, fn, *a, **k):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=fn(*a, **k); torch.cuda.synchronize(); T[name]=T.get(name,0)+time.perf_counter()-t0; return r
# monkeypatch pieces
orig=D._p2p
def p2p(ops, group): return timed("p2p", orig, ops, group)
D._p2p=p2p
for m in ["halo_export","halo_import","slab_classify","halo_kinds_export","halo_kinds_import","slab_finalize","halo_capacity"]:
    f=getattr(VoxelObject, m)
    def mk(f,m):
        def w(self,*a): return timed(m, f, self, *a)
        return w
    setattr(VoxelObject, m, mk(f,m))
with torch.cuda.stream(stream):
    for it in range(6):
        if it==3: T.clear()
        obj=VoxelObject.generate(vg, slab)
        timed("total_halo", D.exchange_halos_and_finalize, obj, ranges, rank, dev)
        mesh=VoxelObjectMesh.create(obj)
        timed("gather", D.gather_mesh, D.device_mesh_tensors(mesh, dev), rank, world, dev)
        obj.free()
if rank==0: print({k: round(1e3*v/3,3) for k,v in T.items()})
dist.destroy_process_group()
