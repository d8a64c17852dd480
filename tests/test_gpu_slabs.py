"""Multi-GPU slab path on the GPU: an object generated as x-slabs with the halo protocol of
include/impact_voxel_cuda.h must give, slab by slab, exactly the rows of the object generated whole — which
tests/test_gpu_parity.py pins to the oracle — and the concatenated slab meshes must equal the whole mesh bit
for bit, order included. The first test holds all slabs in one process (peer copy instead of NCCL); the
second runs one process per GPU over NCCL when the box has two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

import helpers as H
from impact_b200 import distributed as D
from impact_b200 import workloads as W
from impact_b200.voxel import SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

pytestmark = pytest.mark.gpu

GRAPHS = {
    "sphere_big_interior": (lambda: H.sphere_graph(70.0), H.SAME0),   # uniform chunks on both sides of a cut
    "box_types": (lambda: H.box_graph(45.0), H.GRADIENT4),
    "zoo": (H.csg_zoo_graph, H.GRADIENT4),
    "asteroid_like": (lambda: H.asteroid_like_graph(24, 40.0), H.GRADIENT4),
    "asteroid_stand_in": (lambda: W.asteroid_stand_in(0.6), H.GRADIENT4),
}


class _MeshLike:
    def __init__(self, m: dict):
        self.n_vertices, self.n_indices, self.n_submeshes = len(m["positions"]), len(m["indices"]), len(m["submeshes"])
        for k, v in m.items():
            setattr(self, k, v)


def _local_exchange(objs, ranges):
    """The slab protocol between slabs held by one process: device buffers handed over directly."""
    dev = torch.device("cuda:0")
    live = [r for r, (b, e) in enumerate(ranges) if b != e]
    for a, b in zip(live[:-1], live[1:]):  # exchange A across each cut
        for src, s_side, dst in ((a, 1, b), (b, 0, a)):
            cap = objs[src].halo_capacity()
            buf = torch.empty(cap, dtype=torch.uint8, device=dev)
            n = objs[src].halo_export(s_side, buf.data_ptr(), cap)
            objs[dst].halo_import(1 - s_side, buf.data_ptr(), n)
    for r in live:
        objs[r].slab_classify()
    for a, b in zip(live[:-1], live[1:]):  # exchange B: upper slab's lowest plane kinds → lower slab
        plane = objs[b].plane_chunks()
        buf = torch.empty(plane, dtype=torch.uint8, device=dev)
        objs[b].halo_kinds_export(0, buf.data_ptr(), plane)
        objs[a].halo_kinds_import(1, buf.data_ptr(), plane)
    for r in live:
        objs[r].slab_finalize()
    torch.cuda.synchronize()


def _merge_downloads(downloads):
    chunks = np.concatenate([c for c, _ in downloads])
    voxels = np.concatenate([v for _, v in downloads])
    off = 0
    pos = 0
    for c, v in downloads:  # data_offset is per slab: shift into the concatenated voxel array
        nu = c["kind"] == 2
        chunks["data_offset"][pos:pos + len(c)][nu] += off
        off += len(v) // 4096
        pos += len(c)
    return chunks, voxels


@pytest.mark.parametrize("name", sorted(GRAPHS))
@pytest.mark.parametrize("world", [2, 3, 5])
def test_slabs_in_one_process_equal_the_whole_object(ctx, name, world):
    make, types = GRAPHS[name]
    gen = ctx.build_generator(make())
    vg = SDFVoxelGenerator(1.0, gen, types)
    whole = VoxelObject.generate(vg)
    wi = whole.info()
    wc, wv = whole.download()
    wm = VoxelObjectMesh.create(whole).download()

    ranges = D.slab_ranges(wi["chunk_counts"][0], world)
    objs = [VoxelObject.generate(vg, r) for r in ranges]
    with pytest.raises(Exception, match="slab_finalize"):
        VoxelObjectMesh.create(objs[0])  # derived state is pending: meshing must refuse, not guess
    _local_exchange(objs, ranges)

    infos = [o.info() for o in objs]
    for key in ("n_void", "n_uniform", "n_non_uniform"):
        assert sum(i[key] for i in infos) == wi[key], key
    chunks, voxels = _merge_downloads([o.download() for o in objs])
    H.assert_objects_equal(chunks, voxels, wc, wv)

    dev = torch.device("cuda:0")
    meshes = [VoxelObjectMesh.create(o) for o in objs]
    merged = D.merged_mesh_to_numpy(D.concat_meshes([D.device_mesh_tensors(m, dev) for m in meshes]))
    H.assert_meshes_equal(merged, _MeshLike(wm))


def test_a_slab_that_covers_everything_needs_no_exchange(ctx):
    gen = ctx.build_generator(H.complex_graph(0.6))
    vg = SDFVoxelGenerator(1.0, gen, H.SAME0)
    whole = VoxelObject.generate(vg)
    n = whole.info()["chunk_counts"][0]
    slab = VoxelObject.generate(vg, (0, n))
    slab.slab_finalize()
    H.assert_objects_equal(*slab.download(), *whole.download())
    with pytest.raises(Exception, match="no neighbour"):
        slab.halo_export(0, 1, 1)


# ---- one process per GPU over NCCL ---------------------------------------------------------------

def _nccl_worker(rank, world, port, name, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch.distributed as dist

        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from impact_b200.voxel import Context

        make, types = GRAPHS[name]
        # the library works on torch's current stream, so NCCL's stream ordering covers its kernels
        stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(stream):
            c = Context(rank, stream.cuda_stream)
            vg = SDFVoxelGenerator(1.0, c.build_generator(make()), types)
            whole = VoxelObject.generate(vg) if rank == 0 else None
            n_planes = whole.info()["chunk_counts"][0] if rank == 0 else 0
            t = torch.tensor([n_planes], device=dev)
            dist.broadcast(t, 0)
            ranges = D.slab_ranges(int(t.item()), world)
            obj = VoxelObject.generate(vg, ranges[rank])
            stats = D.exchange_halos_and_finalize(obj, ranges, rank, dev)
            mesh = VoxelObjectMesh.create(obj)
            merged = D.gather_mesh(D.device_mesh_tensors(mesh, dev), rank, world, dev)
            # the same gather through peer memory (ivx_mesh_push over NVLink), twice: the second call reuses the block
            pg = D.PeerMeshGather(c, rank, world, dev)
            pushed = pg.gather(mesh)
            pushed = pg.gather(mesh)
            if rank == 0:
                assert stats["halo_bytes_received"] > 0
                wm = VoxelObjectMesh.create(whole).download()
                H.assert_meshes_equal(D.merged_mesh_to_numpy(merged), _MeshLike(wm))
                H.assert_meshes_equal(D.merged_mesh_to_numpy(pushed), _MeshLike(wm))
                wc, wv = whole.download()
                oc, ov = obj.download()
                H.assert_objects_equal(oc, ov, wc[: len(oc)], wv)
            pg.close()
            # the communicator of the C ABI (ivx_comm_*): halo exchange + gather over peer memory, three steps
            wm_counts = torch.tensor([merged["positions"].shape[0], merged["indices"].shape[0], merged["submeshes"].shape[0]]
                                     if rank == 0 else [0, 0, 0], dtype=torch.int64, device=dev)
            dist.broadcast(wm_counts, 0)
            cc = obj.info()["chunk_counts"]
            comm = D.PeerComm(c, rank, world, int(cc[1]) * int(cc[2]), [int(x) + 64 for x in wm_counts.tolist()], device=dev)
            for _ in range(3):
                o2 = VoxelObject.generate(vg, ranges[rank])
                comm.exchange_halos(o2, ranges)
                _, gm = comm.mesh_gather(o2)
                if rank == 0:
                    H.assert_meshes_equal(D.merged_mesh_to_numpy(D.PeerComm.merged_to_torch(gm, dev)), _MeshLike(wm))
                    oc2, ov2 = o2.download()
                    H.assert_objects_equal(oc2, ov2, wc[: len(oc2)], wv)
                o2.free()
            comm.close()
            stream.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        import traceback

        q.put((rank, traceback.format_exc() + repr(e)))


@pytest.mark.parametrize("name", ["asteroid_like", "sphere_big_interior"])
def test_slabs_over_nccl_equal_the_whole_object(name):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_nccl_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_plane_work_estimate_tracks_the_object(ctx):
    """ivx_program_plane_work: deterministic, one entry per chunk plane, zero-ish outside the object and largest
    through its middle; the weighted partition of an off-centre object differs from the equal-thickness one."""
    from impact_b200.voxel import plane_work
    from impact_b200.graph import SDFGraph

    g = SDFGraph()
    a = g.sphere(60.0)
    b = g.translation(g.sphere(12.0), [150.0, 0.0, 0.0])
    g.union(a, b, 1.0)
    vg = SDFVoxelGenerator(1.0, ctx.build_generator(g), H.GRADIENT4)
    w = plane_work(vg)
    w2 = plane_work(vg)
    obj = VoxelObject.generate(vg)
    assert len(w) == obj.info()["chunk_counts"][0] and np.array_equal(w, w2)
    assert w[: len(w) // 3].sum() > 3 * w[2 * len(w) // 3:].sum()  # the big sphere sits at low x
    r = D.slab_ranges_weighted(w, 4)
    assert r != D.slab_ranges(len(w), 4) and r[0][0] == 0 and r[-1][1] == len(w)
