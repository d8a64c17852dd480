"""Host logic of the multi-GPU path on CPU: world_size 2 and 3 over `gloo`.

The slab protocol itself runs on the GPU (tests/test_gpu_slabs.py); here a recording stand-in for the slab
object checks that `exchange_halos_and_finalize` moves every boundary plane to the right neighbour's right
side in the right order, and `gather_mesh` is checked against the oracle: the oracle's mesh of a whole
object is cut into per-slab pieces, gathered over gloo, and must come back identical."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from impact_b200 import distributed as D


def test_weighted_slab_ranges_balance_the_work():
    rng = np.random.default_rng(0)
    for planes in (1, 7, 60, 64):
        for world in (1, 2, 3, 8):
            w = rng.integers(0, 100, planes)
            r = D.slab_ranges_weighted(w, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == planes
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1)) and all(a <= b for a, b in r)
    # an ellipsoid-like profile: the equal-thickness split leaves the end ranks idle, the weighted one does not
    x = np.linspace(-1, 1, 60)
    w = np.maximum(1.0 - x * x, 0.0) * 1000 + 1
    sums = [w[a:b].sum() for a, b in D.slab_ranges_weighted(w, 8)]
    even = [w[a:b].sum() for a, b in D.slab_ranges(60, 8)]
    assert max(sums) < 1.25 * w.sum() / 8 < max(even)
    assert D.slab_ranges_weighted(np.zeros(10), 4) == D.slab_ranges(10, 4)


def test_slab_ranges_cover_the_planes_without_overlap():
    for planes in (0, 1, 5, 60, 64):
        for world in (1, 2, 3, 8):
            r = D.slab_ranges(planes, world)
            assert len(r) == world and r[0][0] == 0 and max(e for _, e in r) == planes
            assert all(r[i][1] == r[i + 1][0] or r[i + 1][0] == r[i + 1][1] for i in range(world - 1))
            assert sum(e - b for b, e in r) == planes
    assert D.slab_ranges(60, 8)[0] == (0, 8) and D.slab_ranges(60, 8)[7] == (56, 60)
    # empty slabs are skipped when looking for neighbours
    r = D.slab_ranges(3, 8)
    assert D.slab_neighbours(r, 0) == (None, 1) and D.slab_neighbours(r, 2) == (1, None)
    assert D.slab_neighbours(r, 5) == (None, None)


class RecordingSlab:
    """Speaks the raw-pointer slab protocol of `VoxelObject`; payloads identify (rank, side)."""

    PLANE = 12

    def __init__(self, rank, ranges):
        self.rank, self.ranges, self.log = rank, ranges, []
        self.lo, self.hi = D.slab_neighbours(ranges, rank)

    @staticmethod
    def _view(ptr, n):
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr))

    def plane_chunks(self):
        return self.PLANE

    def halo_capacity(self):
        return 4096

    def halo_export(self, side, ptr, cap):
        assert cap == self.halo_capacity()  # fixed-size messages: no size handshake
        self._view(ptr, cap)[:] = (self.rank * 2 + side) & 0xFF
        self.log.append(("export", side))
        return cap

    def halo_import(self, side, ptr, nbytes):
        peer = self.lo if side == 0 else self.hi
        assert peer is not None
        assert nbytes == self.halo_capacity()
        assert (self._view(ptr, nbytes) == ((peer * 2 + (1 - side)) & 0xFF)).all()
        self.log.append(("import", side))

    def slab_classify(self):
        self.log.append(("classify",))

    def halo_kinds_export(self, side, ptr, cap):
        assert side == 0 and cap >= self.PLANE
        self._view(ptr, self.PLANE)[:] = 200 + self.rank
        self.log.append(("kinds_export", side))

    def halo_kinds_import(self, side, ptr, nbytes):
        assert side == 1 and nbytes == self.PLANE
        assert (self._view(ptr, nbytes) == 200 + self.hi).all()
        self.log.append(("kinds_import", side))

    def slab_finalize(self):
        self.log.append(("finalize",))


def _cut_mesh(om, world):
    """Cuts the oracle's whole-object mesh into `world` per-slab pieces with local offsets."""
    ns = om.n_submeshes
    cuts = [ns * r // world for r in range(world + 1)]
    parts = []
    for r in range(world):
        s0, s1 = cuts[r], cuts[r + 1]
        v0 = int(om.vertex_ranges[s0, 0]) if s0 < ns else om.n_vertices
        v1 = int(om.vertex_ranges[s1, 0]) if s1 < ns else om.n_vertices
        i0 = int(om.submeshes["index_offset"][s0]) if s0 < ns else om.n_indices
        i1 = int(om.submeshes["index_offset"][s1]) if s1 < ns else om.n_indices
        if r == 0:
            v0 = i0 = 0
        sm = om.submeshes[s0:s1].copy()
        sm["index_offset"] -= i0
        parts.append({
            "positions": om.positions[v0:v1], "normals": om.normals[v0:v1],
            "indices": (om.indices[i0:i1] - v0).astype(np.uint32), "index_materials": om.index_materials[i0:i1],
            "submeshes": sm, "vertex_ranges": (om.vertex_ranges[s0:s1] - v0).astype(np.uint32),
        })
    return parts


def _worker(rank, world, port, planes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ranges = D.slab_ranges(planes, world)
        slab = RecordingSlab(rank, ranges)
        stats = D.exchange_halos_and_finalize(slab, ranges, rank, torch.device("cpu"))
        lo, hi = D.slab_neighbours(ranges, rank)
        want = []
        if ranges[rank][0] != ranges[rank][1]:
            want = [("export", s) for s, p in ((0, lo), (1, hi)) if p is not None]
            want += [("import", s) for s, p in ((0, lo), (1, hi)) if p is not None]
            want += [("classify",)]
            want += [("kinds_export", 0)] if lo is not None else []
            want += [("kinds_import", 1)] if hi is not None else []
            want += [("finalize",)]
        assert slab.log == want, (slab.log, want)
        assert stats["halo_bytes_sent"] == sum(slab.halo_capacity() for s, p in ((0, lo), (1, hi)) if p is not None) + (
            slab.PLANE if lo is not None else 0)

        # mesh gather against the oracle
        from oracle import oracle_lib as O

        g = H.sphere_union_graph(0.5)
        vg = O.VoxelGenerator(O.Generator(g.nodes(), g.root_node_id), 1.0, H.SAME0)
        om = O.Object.generate(vg, 1).mesh(1)
        parts = _cut_mesh(om, world)
        merged = D.gather_mesh(D.host_mesh_tensors(parts[rank]), rank, world, torch.device("cpu"))
        if rank == 0:
            H.assert_meshes_equal(D.merged_mesh_to_numpy(merged), om)
            assert om.n_submeshes > world
        else:
            assert merged is None
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001 - reported to the parent
        import traceback

        q.put((rank, traceback.format_exc() + repr(e)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,planes", [(2, 7), (3, 8), (3, 2)])
def test_halo_exchange_and_mesh_gather_over_gloo(oracle, world, planes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, planes, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_concat_meshes_equals_the_whole_mesh(oracle):
    g = H.complex_graph(0.6)
    vg = oracle.VoxelGenerator(oracle.Generator(g.nodes(), g.root_node_id), 1.0, H.GRADIENT4)
    om = oracle.Object.generate(vg, 2).mesh(1)
    parts = [D.host_mesh_tensors(p) for p in _cut_mesh(om, 4)]
    H.assert_meshes_equal(D.merged_mesh_to_numpy(D.concat_meshes(parts)), om)
