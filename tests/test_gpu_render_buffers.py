"""`ivx_mesh_gpu_buffers_*` = `VoxelMeshGPUBuffers` (gpu_resource.rs:460-900) as exportable device allocations: an
independent consumer — the CUDA driver API through cuda-python, standing in for a Vulkan / wgpu external-memory import —
imports each buffer's file descriptor ONCE and must read, through that mapping, the object's mesh byte for byte: after
creation, after every mesh sync (only the updated ranges are copied, device to device), and through a new descriptor when
a buffer outgrew its allocation."""
import numpy as np
import pytest

import helpers as H
from impact_b200 import _lib as L
from impact_b200.graph import SDFGraph
from impact_b200.voxel import SDFVoxelGenerator, VoxelMeshGPUBuffers, VoxelObject, VoxelObjectMesh

pytestmark = pytest.mark.gpu

drv = pytest.importorskip("cuda.bindings.driver")


def _ok(result):
    err = result[0]
    assert err == drv.CUresult.CUDA_SUCCESS, err
    return result[1] if len(result) == 2 else result[1:]


class Importer:
    """What a renderer does with a descriptor: import the allocation, map it, read it."""

    def __init__(self):
        _ok(drv.cuInit(0))
        dev = _ok(drv.cuDeviceGet(0))
        _ok(drv.cuCtxSetCurrent(_ok(drv.cuDevicePrimaryCtxRetain(dev))))
        self.maps = {}

    def attach(self, name, fd, allocation_bytes):
        self.detach(name)
        handle = _ok(drv.cuMemImportFromShareableHandle(fd, drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR))
        ptr = _ok(drv.cuMemAddressReserve(allocation_bytes, 0, 0, 0))
        _ok(drv.cuMemMap(ptr, allocation_bytes, 0, handle, 0))
        acc = drv.CUmemAccessDesc()
        acc.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = 0
        acc.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        _ok(drv.cuMemSetAccess(ptr, allocation_bytes, [acc], 1))
        self.maps[name] = (ptr, handle, allocation_bytes)

    def detach(self, name):
        if name in self.maps:
            ptr, handle, size = self.maps.pop(name)
            _ok(drv.cuMemUnmap(ptr, size))
            _ok(drv.cuMemAddressFree(ptr, size))
            _ok(drv.cuMemRelease(handle))

    def read(self, name, nbytes) -> np.ndarray:
        out = np.zeros(max(1, nbytes), np.uint8)
        if nbytes:
            _ok(drv.cuMemcpyDtoH(out.ctypes.data, self.maps[name][0], nbytes))
        return out[:nbytes]

    def close(self):
        for name in list(self.maps):
            self.detach(name)


def _mesh_bytes(m: dict) -> dict:
    return {"positions": m["positions"].view(np.uint8).ravel(), "normals": m["normals"].view(np.uint8).ravel(),
            "index_materials": m["index_materials"].view(np.uint8).ravel(), "indices": m["indices"].view(np.uint8).ravel(),
            "chunk_submeshes": m["submeshes"].view(np.uint8).ravel()}


def _check(imp: Importer, bufs: VoxelMeshGPUBuffers, mesh: VoxelObjectMesh, ctx, what):
    ctx.synchronize()  # the copies run on the context's stream
    want = _mesh_bytes(mesh.download())
    assert (bufs.n_vertices, bufs.n_indices, bufs.n_chunks) == (mesh.n_vertices, mesh.n_indices, mesh.n_submeshes), what
    for name in L.MESH_BUFFER_NAMES:
        assert bufs.valid_bytes[name] == len(want[name]) <= bufs.allocation_bytes[name], f"{what}: {name} sizes"
        got = imp.read(name, bufs.valid_bytes[name])
        assert np.array_equal(got, want[name]), f"{what}: {name} differs through the imported mapping"


def _attach_all(imp, bufs, names=None):
    for name in names or L.MESH_BUFFER_NAMES:
        assert bufs.fds[name] >= 0
        imp.attach(name, bufs.fds[name], bufs.allocation_bytes[name])


def test_buffers_follow_the_synced_mesh_through_one_import(ctx):
    graph = H.asteroid_like_graph(24, 40.0)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), H.GRADIENT4))
    mesh = VoxelObjectMesh.create(obj)
    bufs = VoxelMeshGPUBuffers.for_voxel_object(obj)
    assert sorted(bufs.recreated) == sorted(L.MESH_BUFFER_NAMES) and bufs.bytes_copied == sum(bufs.valid_bytes.values())
    imp = Importer()
    try:
        _attach_all(imp, bufs)
        _check(imp, bufs, mesh, ctx, "created")
        # nothing modified: a sync moves nothing
        mesh = VoxelObjectMesh.sync(obj)
        bufs.sync_with_voxel_object()
        assert bufs.bytes_copied == 0 and not bufs.recreated
        rng = np.random.default_rng(5)
        shape = np.array(obj.info()["grid_shape"], np.float64)
        total = 0
        for step in range(8):
            c = (shape * rng.uniform(0.2, 0.8, 3)).astype(np.float32)
            r = float(rng.uniform(4, 9))
            obj.absorb_sphere(c, r, r + 2.0)
            mesh = VoxelObjectMesh.sync(obj)
            upd, removed = mesh.modifications()
            bufs.sync_with_voxel_object()
            assert bufs.n_updated_ranges == len(upd)
            # sync_with_voxel_object ends with report_gpu_resources_synchronized
            assert len(mesh.modifications()[0]) == 0 and not mesh.modifications()[1]
            if bufs.recreated:
                _attach_all(imp, bufs, bufs.recreated)
            else:
                ranges = 24 * int((upd[:, 1] - upd[:, 0]).sum()) + 12 * int((upd[:, 3] - upd[:, 2]).sum())
                assert bufs.bytes_copied == ranges + bufs.valid_bytes["chunk_submeshes"] if (len(upd) or removed) else bufs.bytes_copied == 0
                assert bufs.bytes_copied < sum(bufs.valid_bytes.values())
            total += len(upd)
            _check(imp, bufs, mesh, ctx, f"step {step}")
        assert total > 0
        # several mesh syncs between two buffer syncs: the modification list accumulates (mesh.rs:833-838)
        for _ in range(3):
            c = (shape * rng.uniform(0.2, 0.8, 3)).astype(np.float32)
            obj.absorb_sphere(c, 6.0, 8.0)
            mesh = VoxelObjectMesh.sync(obj)
        bufs.sync_with_voxel_object()
        if bufs.recreated:
            _attach_all(imp, bufs, bufs.recreated)
        _check(imp, bufs, mesh, ctx, "accumulated")
    finally:
        imp.close()
        bufs.close()


def test_outgrown_buffers_are_recreated_with_a_new_descriptor(ctx):
    # a solid box: carving spheres out of its inside multiplies the surface, and the slices outgrow their allocations
    # (25 % headroom, rounded up to the 2 MiB granularity of exportable memory: the object has to be this large)
    graph = SDFGraph()
    graph.box([200.0] * 3)
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(graph), H.SAME0))
    mesh = VoxelObjectMesh.create(obj)
    bufs = VoxelMeshGPUBuffers.for_voxel_object(obj)
    imp = Importer()
    try:
        _attach_all(imp, bufs)
        _check(imp, bufs, mesh, ctx, "created")
        first_fds = dict(bufs.fds)
        rng = np.random.default_rng(3)
        recreated = set()
        for step in range(14):
            c = rng.uniform(45, 155, 3).astype(np.float32)
            obj.absorb_sphere(c, 35.0, 37.0)
            mesh = VoxelObjectMesh.sync(obj)
            bufs.sync_with_voxel_object()
            if bufs.recreated:
                # vertex buffers and index buffers are re-created in pairs (gpu_resource.rs:748-830)
                r = set(bufs.recreated)
                assert ("positions" in r) == ("normals" in r) and ("indices" in r) == ("index_materials" in r)
                for name in bufs.recreated:
                    assert bufs.fds[name] >= 0 and bufs.allocation_bytes[name] >= bufs.valid_bytes[name]
                _attach_all(imp, bufs, bufs.recreated)
                recreated |= r
            _check(imp, bufs, mesh, ctx, f"step {step}")
        assert {"positions", "normals", "indices", "index_materials"} <= recreated, recreated
        assert first_fds  # (descriptors of replaced allocations were closed by the wrapper)
    finally:
        imp.close()
        bufs.close()


def test_a_mesh_created_anew_is_copied_whole_and_errors_are_reported(ctx):
    obj = VoxelObject.generate(SDFVoxelGenerator(1.0, ctx.build_generator(H.sphere_graph(20.0)), H.SAME0))
    with pytest.raises(Exception, match="ivx_object_mesh first"):
        VoxelMeshGPUBuffers.for_voxel_object(obj)
    mesh = VoxelObjectMesh.create(obj)
    bufs = VoxelMeshGPUBuffers.for_voxel_object(obj)
    imp = Importer()
    try:
        _attach_all(imp, bufs)
        obj.absorb_sphere(np.float32([20, 20, 38]), 5.0, 7.0)
        mesh = VoxelObjectMesh.create(obj)  # a full re-mesh instead of a sync: no modification list to follow
        bufs.sync_with_voxel_object()
        if bufs.recreated:
            _attach_all(imp, bufs, bufs.recreated)
        assert bufs.bytes_copied >= sum(bufs.valid_bytes.values()) - bufs.valid_bytes["chunk_submeshes"]
        _check(imp, bufs, mesh, ctx, "re-created mesh")
    finally:
        imp.close()
        bufs.close()
