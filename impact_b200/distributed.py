"""Multi-GPU orchestration of the voxel hot path: one process per GPU, x-slabs of chunk planes.

The reference parallelises generation over contiguous ranges of the x-major linear chunk index
(`VoxelObject::generate_without_derived_state_in_parallel`, object.rs:402-470) and meshing over exposed
chunks (mesh.rs:286-354). The same partition maps onto GPUs as slabs of whole chunk planes:

  * generation needs no communication;
  * derived state and Surface Nets read one chunk plane of the neighbouring slab (the 1-voxel brick padding,
    object/sdf.rs:35, and the face rules, object.rs:1682-1704) → ONE point-to-point halo exchange with each
    neighbour (NCCL send/recv over NVLink; only the touching voxel layer of each chunk travels), plus one byte per chunk of the upper neighbour's final chunk kinds
    (quad ownership, surface_nets.rs:252-261);
  * the per-slab meshes are concatenated on one rank in slab order, which IS the reference's order (chunks
    are meshed in linear chunk order), with vertex / index offsets rebased.

There is no other collective on the data path. `torch.distributed` is plumbing only: the functions take any
initialised process group (NCCL on the GPUs, gloo in the CPU tests) and objects exposing the slab protocol
of `impact_b200.voxel.VoxelObject` (raw-pointer methods), so the host logic is testable without a GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def slab_ranges(n_planes: int, world: int) -> list[tuple[int, int]]:
    """Contiguous chunk-plane ranges per rank: ceil split, like the reference's per-thread split of the
    linear chunk index (object.rs:423-427). Trailing ranks may be empty when n_planes < world."""
    per = (n_planes + world - 1) // world if world > 0 else 0
    return [(min(n_planes, r * per), min(n_planes, (r + 1) * per)) for r in range(world)]


def slab_ranges_weighted(work, world: int) -> list[tuple[int, int]]:
    """Contiguous chunk-plane ranges per rank with (nearly) equal summed `work` (one non-negative number per
    plane, e.g. `impact_b200.voxel.plane_work`): the r-th cut goes where the running sum is closest to r / world
    of the total. Deterministic in `work`; ranks must pass identical arrays. Falls back to `slab_ranges` for
    all-zero work."""
    w = np.asarray(work, np.float64)
    n = len(w)
    total = float(w.sum())
    if world <= 0 or n == 0 or not total > 0.0:
        return slab_ranges(n, world)
    prefix = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(prefix, target))
        i = min(max(i, 1), n)
        if abs(prefix[i - 1] - target) <= abs(prefix[i] - target):
            i -= 1
        cuts.append(min(n, max(cuts[-1], i)))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def slab_neighbours(ranges: list[tuple[int, int]], rank: int) -> tuple[int | None, int | None]:
    """Ranks owning the planes just below / above `rank`'s slab (empty slabs are skipped)."""
    b, e = ranges[rank]
    if b == e:
        return None, None
    lo = next((r for r in range(rank - 1, -1, -1) if ranges[r][0] < ranges[r][1]), None)
    hi = next((r for r in range(rank + 1, len(ranges)) if ranges[r][0] < ranges[r][1]), None)
    return lo, hi


def _p2p(ops, group):
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def exchange_halos_and_finalize(obj, ranges, rank: int, device, group=None) -> dict:
    """Runs the slab protocol of include/impact_voxel_cuda.h for this rank's slab object.

    → {"halo_bytes_sent": …, "halo_bytes_received": …}
    """
    lo, hi = slab_neighbours(ranges, rank)
    stats = {"halo_bytes_sent": 0, "halo_bytes_received": 0}
    if ranges[rank][0] == ranges[rank][1]:
        return stats
    peers = [(0, lo), (1, hi)]
    cap = obj.halo_capacity()

    # exchange A: boundary planes. A message has a fixed size (chunk descriptors + one 768-byte voxel layer per
    # chunk of the plane), so both directions of both neighbours go out in ONE batch with no size handshake.
    out_buf, in_buf = {}, {}
    ops = []
    for side, peer in peers:
        if peer is None:
            continue
        out_buf[side] = torch.empty(cap, dtype=torch.uint8, device=device)
        n = obj.halo_export(side, out_buf[side].data_ptr(), cap)
        assert n == cap, "halo messages have the fixed size ivx_object_halo_capacity reports"
        in_buf[side] = torch.empty(cap, dtype=torch.uint8, device=device)
        ops.append(dist.P2POp(dist.isend, out_buf[side], peer, group))
        ops.append(dist.P2POp(dist.irecv, in_buf[side], peer, group))
        stats["halo_bytes_sent"] += cap
        stats["halo_bytes_received"] += cap
    # halo_export / halo_import / kinds / slab_classify only ENQUEUE on the context's stream (include/impact_voxel_cuda.h);
    # NCCL runs on torch's current stream. Unless the two are the same stream, order them explicitly.
    same_stream = device is not None and torch.device(device).type == "cuda" and \
        getattr(obj.ctx, "stream_handle", None) == torch.cuda.current_stream().cuda_stream
    fence_lib = (lambda: None) if same_stream or not hasattr(obj, "ctx") else obj.ctx.synchronize
    fence_torch = (lambda: None) if same_stream or not torch.cuda.is_available() else (lambda: torch.cuda.current_stream().synchronize())
    fence_lib()
    _p2p(ops, group)
    fence_torch()
    for side, buf in in_buf.items():
        obj.halo_import(side, buf.data_ptr(), cap)

    obj.slab_classify()

    # exchange B: final non-uniform flags of my lowest plane → the rank below (its +x neighbour chunks)
    plane = obj.plane_chunks()
    ops = []
    k_out = k_in = None
    if lo is not None:
        k_out = torch.empty(plane, dtype=torch.uint8, device=device)
        obj.halo_kinds_export(0, k_out.data_ptr(), plane)
        ops.append(dist.P2POp(dist.isend, k_out, lo, group))
        stats["halo_bytes_sent"] += plane
    if hi is not None:
        k_in = torch.empty(plane, dtype=torch.uint8, device=device)
        ops.append(dist.P2POp(dist.irecv, k_in, hi, group))
        stats["halo_bytes_received"] += plane
    fence_lib()
    _p2p(ops, group)
    fence_torch()
    if k_in is not None:
        obj.halo_kinds_import(1, k_in.data_ptr(), plane)
    obj.slab_finalize()
    return stats


# ---- mesh gather -------------------------------------------------------------------------------------------
MESH_FIELDS = (  # name, trailing shape, torch dtype, "count" key
    ("positions", (3,), torch.float32, "v"),
    ("normals", (3,), torch.float32, "v"),
    ("indices", (), torch.int32, "i"),
    ("index_materials", (8,), torch.uint8, "i"),
    ("submeshes", (13,), torch.int32, "s"),
    ("vertex_ranges", (2,), torch.int32, "s"),
)
_SUBMESH_INDEX_OFFSET_COL = 3  # ivx_chunk_submesh.index_offset (include/impact_voxel_cuda.h)


class _DeviceArray:
    """Zero-copy view of a device buffer owned by the library (ivx_mesh_info.d_*) for torch.as_tensor."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_mesh_tensors(mesh, device) -> dict:
    """Wraps the device buffers of a `VoxelObjectMesh` as torch tensors (no copy)."""
    info = mesh.device_info
    nv, ni, ns = mesh.n_vertices, mesh.n_indices, mesh.n_submeshes

    def wrap(ptr, shape, typestr, dtype):
        n = int(np.prod(shape))
        if n == 0 or not ptr:
            return torch.empty(shape, dtype=dtype, device=device)
        return torch.as_tensor(_DeviceArray(ptr, shape, typestr), device=device)

    return {
        "positions": wrap(info.d_positions, (nv, 3), "<f4", torch.float32),
        "normals": wrap(info.d_normals, (nv, 3), "<f4", torch.float32),
        "indices": wrap(info.d_indices, (ni,), "<i4", torch.int32),
        "index_materials": wrap(info.d_index_materials, (ni, 8), "|u1", torch.uint8),
        "submeshes": wrap(info.d_submeshes, (ns, 13), "<i4", torch.int32),
        "vertex_ranges": wrap(info.d_vertex_ranges, (ns, 2), "<i4", torch.int32),
    }


def host_mesh_tensors(mesh: dict) -> dict:
    """The dict returned by `VoxelObjectMesh.download()` / the oracle as CPU tensors in the gather layout."""
    return {
        "positions": torch.from_numpy(np.ascontiguousarray(mesh["positions"], np.float32).reshape(-1, 3)),
        "normals": torch.from_numpy(np.ascontiguousarray(mesh["normals"], np.float32).reshape(-1, 3)),
        "indices": torch.from_numpy(np.ascontiguousarray(mesh["indices"]).view(np.int32).reshape(-1)),
        "index_materials": torch.from_numpy(np.ascontiguousarray(mesh["index_materials"]).view(np.uint8).reshape(-1, 8)),
        "submeshes": torch.from_numpy(np.ascontiguousarray(mesh["submeshes"]).view(np.int32).reshape(-1, 13)),
        "vertex_ranges": torch.from_numpy(np.ascontiguousarray(mesh["vertex_ranges"]).view(np.int32).reshape(-1, 2)),
    }


def gather_mesh(parts: dict, rank: int, world: int, device, dst: int = 0, group=None):
    """Concatenates the per-slab meshes on `dst` in slab (= linear chunk) order (mesh.rs:286-354).

    `parts`: this rank's tensors in the MESH_FIELDS layout. Indices are rebased by the vertex count of the
    lower slabs, submesh index offsets by their index count, vertex ranges by their vertex count.
    → the merged dict on `dst`, None elsewhere.
    """
    counts = torch.tensor([parts["positions"].shape[0], parts["indices"].shape[0], parts["submeshes"].shape[0]],
                          dtype=torch.int64, device=device)
    all_counts = [torch.zeros(3, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    table = torch.stack(all_counts).cpu().numpy()  # [world, 3]
    key_col = {"v": 0, "i": 1, "s": 2}
    if rank != dst:
        ops = [dist.P2POp(dist.isend, parts[name].contiguous(), dst, group)
               for name, _, _, key in MESH_FIELDS if table[rank, key_col[key]] > 0]
        _p2p(ops, group)
        return None
    totals = table.sum(axis=0)
    starts = np.concatenate([np.zeros((1, 3), np.int64), np.cumsum(table, axis=0)[:-1]])
    merged = {name: torch.empty((int(totals[key_col[key]]),) + shape, dtype=dtype, device=device)
              for name, shape, dtype, key in MESH_FIELDS}
    ops = []
    for r in range(world):
        for name, _, _, key in MESH_FIELDS:
            n, s = int(table[r, key_col[key]]), int(starts[r, key_col[key]])
            if n == 0:
                continue
            if r == rank:
                merged[name][s:s + n].copy_(parts[name])
            else:
                ops.append(dist.P2POp(dist.irecv, merged[name][s:s + n], r, group))
    _p2p(ops, group)
    _rebase(merged, table, starts)
    return merged


# ---- mesh gather over peer memory (NVLink / NVSwitch stores instead of NCCL messages) -------------------------
_FIELD_BYTES = (("positions", 12, "v"), ("normals", 12, "v"), ("indices", 4, "i"), ("index_materials", 8, "i"),
                ("submeshes", 52, "s"), ("vertex_ranges", 8, "s"))


class PeerMeshGather:
    """The merged mesh lives in one block of the gathering GPU's memory, exported once through CUDA IPC
    (`ivx_peer_alloc` / `ivx_peer_open`); every step each rank writes its slab's mesh straight into it with
    `ivx_mesh_push` — indices, submesh index offsets and vertex ranges rebased on the fly — so the only collective
    left is the all-gather of three counts per rank. The block grows (all ranks agree from the gathered counts)
    when a step needs more room."""

    def __init__(self, ctx, rank: int, world: int, device, dst: int = 0, group=None):
        self.ctx, self.rank, self.world, self.device, self.dst, self.group = ctx, rank, world, device, dst, group
        self.caps = np.zeros(3, np.int64)  # vertices, indices, submeshes
        self.base = None                   # pointer valid on this rank's device
        self.offsets = None
        self.available = True              # False once CUDA IPC turned out not to work on some rank

    def _layout(self, caps):
        key = {"v": int(caps[0]), "i": int(caps[1]), "s": int(caps[2])}
        offs, o = [], 0
        for _, nbytes, k in _FIELD_BYTES:
            offs.append(o)
            o += (key[k] * nbytes + 255) // 256 * 256
        return offs, max(o, 256)

    def _release(self):
        if self.base is not None:
            import ctypes as C
            lib = self.ctx._lib
            (lib.ivx_peer_free if self.rank == self.dst else lib.ivx_peer_close)(self.ctx.h, C.c_void_p(self.base))
            self.base = None

    def _ensure(self, totals):
        import ctypes as C
        if self.base is not None and np.all(totals <= self.caps):
            return
        self.ctx.synchronize()
        dist.barrier(group=self.group)  # nobody is still writing into the old block
        self._release()
        self.caps = np.maximum((totals * 1.25).astype(np.int64) + 1024, self.caps)
        self.offsets, nbytes = self._layout(self.caps)
        handle = torch.zeros(64, dtype=torch.uint8, device=self.device)
        lib = self.ctx._lib
        ok = 1
        if self.rank == self.dst:
            ptr = C.c_void_p()
            hbuf = (C.c_ubyte * 64)()
            if lib.ivx_peer_alloc(self.ctx.h, C.c_size_t(nbytes), C.byref(ptr), hbuf) == 0:
                self.base = ptr.value
                handle.copy_(torch.tensor(list(hbuf), dtype=torch.uint8))
            else:
                ok = 0
        dist.broadcast(handle, self.dst, group=self.group)
        if self.rank != self.dst:
            hb = (C.c_ubyte * 64)(*handle.cpu().tolist())
            ptr = C.c_void_p()
            if lib.ivx_peer_open(self.ctx.h, hb, C.byref(ptr)) == 0:
                self.base = ptr.value
            else:
                ok = 0
        # CUDA IPC can be unavailable (container / driver settings): all ranks agree, and the caller falls back to the
        # NCCL send/recv gather
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self._release()
            self.available = False

    def gather(self, mesh):
        """`mesh`: this rank's `VoxelObjectMesh`. → the merged dict of device tensors on `dst` (views into the
        block, valid until the next call), None elsewhere."""
        import ctypes as C
        counts = torch.tensor([mesh.n_vertices, mesh.n_indices, mesh.n_submeshes], dtype=torch.int64, device=self.device)
        all_counts = torch.zeros(self.world * 3, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(all_counts, counts, group=self.group)
        table = all_counts.cpu().numpy().reshape(self.world, 3)
        totals = table.sum(axis=0)
        starts = np.concatenate([np.zeros((1, 3), np.int64), np.cumsum(table, axis=0)[:-1]])
        if self.available:
            self._ensure(totals)
        if not self.available:
            return gather_mesh(device_mesh_tensors(mesh, self.device), self.rank, self.world, self.device, self.dst, self.group)
        v0, i0, s0 = (int(x) for x in starts[self.rank])
        offs = (C.c_uint64 * 6)(*self.offsets)
        self.ctx.check(self.ctx._lib.ivx_mesh_push(self.ctx.h, mesh.obj.h, C.c_void_p(self.base), offs, C.c_uint32(v0),
                                                   C.c_uint32(i0), C.c_uint32(s0)))
        dist.barrier(group=self.group)  # stream-ordered after the pushes: every part has landed
        if self.rank != self.dst:
            return None
        nv, ni, ns = (int(x) for x in totals)
        o = self.offsets

        def view(off, shape, typestr, dtype):
            if int(np.prod(shape)) == 0:
                return torch.empty(shape, dtype=dtype, device=self.device)
            return torch.as_tensor(_DeviceArray(self.base + off, shape, typestr), device=self.device)

        return {
            "positions": view(o[0], (nv, 3), "<f4", torch.float32), "normals": view(o[1], (nv, 3), "<f4", torch.float32),
            "indices": view(o[2], (ni,), "<i4", torch.int32), "index_materials": view(o[3], (ni, 8), "|u1", torch.uint8),
            "submeshes": view(o[4], (ns, 13), "<i4", torch.int32), "vertex_ranges": view(o[5], (ns, 2), "<i4", torch.int32),
        }

    def close(self):
        self.ctx.synchronize()
        dist.barrier(group=self.group)
        self._release()


def _rebase(merged: dict, table: np.ndarray, starts: np.ndarray) -> None:
    """Part r's indices / submesh index offsets / vertex ranges become offsets into the merged buffers."""
    for r in range(len(table)):
        v0, i0, s0 = (int(x) for x in starts[r])
        nv, ni, ns = (int(x) for x in table[r])
        if v0 and ni:
            merged["indices"][i0:i0 + ni] += v0
        if ns:
            if i0:
                merged["submeshes"][s0:s0 + ns, _SUBMESH_INDEX_OFFSET_COL] += i0
            if v0:
                merged["vertex_ranges"][s0:s0 + ns] += v0


def concat_meshes(parts_list: list[dict]) -> dict:
    """Single-process form of `gather_mesh` (several slabs held by one process)."""
    table = np.array([[p["positions"].shape[0], p["indices"].shape[0], p["submeshes"].shape[0]] for p in parts_list],
                     np.int64).reshape(-1, 3)
    starts = np.concatenate([np.zeros((1, 3), np.int64), np.cumsum(table, axis=0)[:-1]])
    merged = {name: torch.cat([p[name].reshape((-1,) + shape) for p in parts_list]).clone()
              for name, shape, _, _ in MESH_FIELDS}
    _rebase(merged, table, starts)
    return merged


def merged_mesh_to_numpy(merged: dict) -> dict:
    """→ the `VoxelObjectMesh.download()` layout (structured dtypes) for comparisons."""
    from . import _lib as L

    g = {k: v.cpu().numpy() for k, v in merged.items()}
    return {
        "positions": g["positions"], "normals": g["normals"], "indices": g["indices"].view(np.uint32),
        "index_materials": np.ascontiguousarray(g["index_materials"]).view(L.INDEX_MATERIALS_DTYPE).reshape(-1),
        "submeshes": np.ascontiguousarray(g["submeshes"]).view(L.SUBMESH_DTYPE).reshape(-1),
        "vertex_ranges": g["vertex_ranges"].view(np.uint32),
    }


# ---- the communicator of the C ABI (ivx_comm_*): peer-memory halo exchange and mesh gather ------------------------------
class PeerComm:
    """`ivx_comm`: every rank's window mapped on every rank; halo planes, quad-ownership bits and mesh parts are stored
    straight into the consumer's memory over NVLink and awaited on the device (include/impact_voxel_cuda.h
    "multi-GPU communicator over peer memory"). The process group is used ONCE, to exchange the 64-byte window handles."""

    def __init__(self, ctx, rank: int, world: int, plane_chunks: int, mesh_capacity, gather_rank: int = 0, group=None,
                 device=None, local_peers=None):
        import ctypes as C
        from . import _lib as L

        self.ctx, self.rank, self.world, self.gather_rank = ctx, rank, world, gather_rank
        cfg = L.CommConfig(rank, world, gather_rank, plane_chunks, int(mesh_capacity[0]), int(mesh_capacity[1]),
                           int(mesh_capacity[2]))
        self.h = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        ctx.check(ctx._lib.ivx_comm_create(ctx.h, C.byref(cfg), C.byref(self.h), handle))
        self.handle = bytes(handle)
        self._local = local_peers is not None
        if local_peers is None:
            self.connect(group, device)

    def connect(self, group=None, device=None):
        """Exchanges the window handles through the process group (any backend) and maps the peers' windows."""
        import ctypes as C

        mine = torch.tensor(list(self.handle), dtype=torch.uint8, device=device)
        allh = torch.zeros(self.world * 64, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        buf = (C.c_ubyte * (64 * self.world))(*allh.cpu().tolist())
        self.ctx.check(self.ctx._lib.ivx_comm_connect(self.ctx.h, self.h, buf))
        dist.barrier(group=group)  # every rank has mapped every window before anybody stores into one

    @staticmethod
    def connect_local(comms):
        """Ranks living in one process (one context each): plain pointers instead of CUDA IPC."""
        import ctypes as C

        arr = (C.c_void_p * len(comms))(*[c.h for c in comms])
        for c in comms:
            c.ctx.check(c.ctx._lib.ivx_comm_connect_local(c.ctx.h, c.h, arr))

    def exchange_halos(self, obj, ranges) -> None:
        """`ivx_object_exchange_halos` for this rank's slab of `ranges` (asynchronous on the context's stream)."""
        import ctypes as C

        lo, hi = slab_neighbours(ranges, self.rank)
        self.ctx.check(self.ctx._lib.ivx_object_exchange_halos(self.ctx.h, self.h, obj.h, C.c_int(-1 if lo is None else lo),
                                                               C.c_int(-1 if hi is None else hi)))

    def mesh_gather(self, obj):
        """`ivx_object_mesh_gather` → (this slab's `VoxelObjectMesh`, merged mesh as numpy-like device views on the gather
        rank / None elsewhere). Synchronises the context's stream."""
        import ctypes as C
        from . import _lib as L
        from .voxel import VoxelObjectMesh

        info, merged = L.MeshInfo(), L.GatheredMesh()
        self.ctx.check(self.ctx._lib.ivx_object_mesh_gather(self.ctx.h, self.h, obj.h, C.byref(info), C.byref(merged)))
        local = VoxelObjectMesh(obj, info)
        if self.rank != self.gather_rank:
            return local, None
        return local, merged

    def mesh_distributed(self, obj):
        """`ivx_object_mesh_distributed`: the step's mesh stays on its rank, rebased in place to the numbering of the whole
        job's mesh → (this slab's `VoxelObjectMesh`, (vertex base, index base, submesh base)). Takes the place of
        `mesh_gather` in a step; synchronises the context's stream."""
        import ctypes as C
        from . import _lib as L
        from .voxel import VoxelObjectMesh

        info = L.MeshInfo()
        bases = (C.c_uint64 * 3)()
        self.ctx.check(self.ctx._lib.ivx_object_mesh_distributed(self.ctx.h, self.h, obj.h, C.byref(info), bases))
        return VoxelObjectMesh(obj, info), (int(bases[0]), int(bases[1]), int(bases[2]))

    @staticmethod
    def merged_to_torch(merged, device) -> dict:
        """The `ivx_gathered_mesh` of the gather rank as torch tensors over the window (no copy)."""
        nv, ni, ns = int(merged.n_vertices), int(merged.n_indices), int(merged.n_submeshes)

        def view(ptr, shape, typestr, dtype):
            if int(np.prod(shape)) == 0 or not ptr:
                return torch.empty(shape, dtype=dtype, device=device)
            return torch.as_tensor(_DeviceArray(ptr, shape, typestr), device=device)

        return {
            "positions": view(merged.d_positions, (nv, 3), "<f4", torch.float32),
            "normals": view(merged.d_normals, (nv, 3), "<f4", torch.float32),
            "indices": view(merged.d_indices, (ni,), "<i4", torch.int32),
            "index_materials": view(merged.d_index_materials, (ni, 8), "|u1", torch.uint8),
            "submeshes": view(merged.d_submeshes, (ns, 13), "<i4", torch.int32),
            "vertex_ranges": view(merged.d_vertex_ranges, (ns, 2), "<i4", torch.int32),
        }

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx._lib.ivx_comm_destroy(self.ctx.h, self.h)
        self.h = None


# ---- several slabs held by ONE process (tests; a host that drives several objects on one device) ---------------
def exchange_halos_single_process(objs, device=None) -> None:
    """The slab protocol between the x-slab objects `objs` (in slab order, all on one device): the halo messages are
    handed over as device buffers, no process group involved."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    live = [o for o in objs if o.info()["chunk_i_begin"] != o.info()["chunk_i_end"]]
    for a, b in zip(live[:-1], live[1:]):  # exchange A across each cut
        for src, s_side, dst in ((a, 1, b), (b, 0, a)):
            cap = src.halo_capacity()
            buf = torch.empty(cap, dtype=torch.uint8, device=device)
            n = src.halo_export(s_side, buf.data_ptr(), cap)
            src.ctx.synchronize()  # the slabs may live on different contexts (streams)
            dst.halo_import(1 - s_side, buf.data_ptr(), n)
            dst.ctx.synchronize()
    for o in live:
        o.slab_classify()
    for a, b in zip(live[:-1], live[1:]):  # exchange B: upper slab's lowest plane kinds → lower slab
        plane = b.plane_chunks()
        buf = torch.empty(plane, dtype=torch.uint8, device=device)
        b.halo_kinds_export(0, buf.data_ptr(), plane)
        b.ctx.synchronize()
        a.halo_kinds_import(1, buf.data_ptr(), plane)
        a.ctx.synchronize()
    for o in live:
        o.slab_finalize()


def merge_mesh_parts(parts: list[dict]) -> dict:
    """`concat_meshes` for downloaded meshes (`VoxelObjectMesh.download()` dicts, numpy): slab order, rebased."""
    v0 = i0 = 0
    out = {k: [] for k in ("positions", "normals", "indices", "index_materials", "submeshes", "vertex_ranges")}
    for p in parts:
        out["positions"].append(p["positions"])
        out["normals"].append(p["normals"])
        out["indices"].append(p["indices"] + np.uint32(v0))
        out["index_materials"].append(p["index_materials"])
        sm = p["submeshes"].copy()
        sm["index_offset"] += np.uint32(i0)
        out["submeshes"].append(sm)
        out["vertex_ranges"].append(p["vertex_ranges"] + np.uint32(v0))
        v0 += len(p["positions"])
        i0 += len(p["indices"])
    return {k: np.concatenate(v) for k, v in out.items()}
