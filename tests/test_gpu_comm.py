"""The multi-GPU communicator of the C ABI (ivx_comm_*: peer-memory halo exchange + mesh gather, csrc/comm.cu) with every
rank's context living in THIS process — one thread per rank, windows connected as plain pointers
(ivx_comm_connect_local) — so the whole device-side protocol (stores into the neighbour's window, flags, on-device
waits, epochs and double buffering over several steps, plans) runs on a one-GPU box too. With two or more GPUs the ranks
are spread over them (real NVLink peer stores). The result must equal the object generated whole, which
tests/test_gpu_parity.py and tests/test_gpu_baseline_sizes.py pin to the oracle: every slab's rows, and the merged mesh
bit for bit, order included."""
import threading

import numpy as np
import pytest
import torch

import helpers as H
from impact_b200 import distributed as D
from impact_b200 import workloads as W
from impact_b200.voxel import Context, SDFVoxelGenerator, VoxelObject, VoxelObjectMesh

pytestmark = pytest.mark.gpu

GRAPHS = {
    "sphere_big_interior": (lambda: H.sphere_graph(70.0), H.SAME0),   # uniform chunks on both sides of a cut
    "zoo": (H.csg_zoo_graph, H.GRADIENT4),
    "asteroid_stand_in": (lambda: W.asteroid_stand_in(0.6), H.GRADIENT4),
}


class _MeshLike:
    def __init__(self, m: dict):
        self.n_vertices, self.n_indices, self.n_submeshes = len(m["positions"]), len(m["indices"]), len(m["submeshes"])
        for k, v in m.items():
            setattr(self, k, v)


def _run(world, fn):
    errors = [None] * world

    def wrap(r):
        try:
            fn(r)
        except BaseException as e:  # noqa: BLE001
            import traceback

            errors[r] = traceback.format_exc() + repr(e)

    threads = [threading.Thread(target=wrap, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    return errors


def _setup(make, types, world, capacity_scale=1.25):
    n_dev = torch.cuda.device_count()
    ctxs = [Context(r % n_dev) for r in range(world)]
    graph = make()
    vgs = [SDFVoxelGenerator(1.0, c.build_generator(graph), types) for c in ctxs]
    whole = VoxelObject.generate(vgs[0])
    wi = whole.info()
    wmesh = VoxelObjectMesh.create(whole)
    cap = [int(x * capacity_scale) + 8 for x in (wmesh.n_vertices, wmesh.n_indices, wmesh.n_submeshes)]
    plane = int(wi["chunk_counts"][1]) * int(wi["chunk_counts"][2])
    comms = [D.PeerComm(ctxs[r], r, world, plane, cap, local_peers=True) for r in range(world)]
    D.PeerComm.connect_local(comms)
    # Ranks that SHARE a device (this test on a one-GPU box) must not allocate device memory while a peer's stream sits in
    # an on-device wait (a device memory allocation is an implicit synchronisation point between streams): one step over
    # the explicit slab protocol first, so that every context's pool already holds the blocks the timed steps need.
    ranges = D.slab_ranges(wi["chunk_counts"][0], world)
    for _ in range(2):
        slabs = [VoxelObject.generate(vgs[r], ranges[r]) for r in range(world)]
        D.exchange_halos_single_process(slabs, torch.device("cuda", 0))
        for sl in slabs:
            VoxelObjectMesh.create(sl)
            sl.download()
        for sl in slabs:
            sl.free()
    return ctxs, vgs, whole, wi, wmesh, comms


@pytest.mark.parametrize("name", sorted(GRAPHS))
@pytest.mark.parametrize("world", [2, 3])
def test_comm_steps_equal_the_whole_object(name, world):
    make, types = GRAPHS[name]
    ctxs, vgs, whole, wi, wmesh, comms = _setup(make, types, world)
    wc, wv = whole.download()
    wm = wmesh.download()
    ranges = D.slab_ranges(wi["chunk_counts"][0], world)
    per_plane = int(wi["chunk_counts"][1]) * int(wi["chunk_counts"][2])
    steps = 4  # both parities twice; from the second step on generation and meshing run from their plans

    def rank(r):
        dev = torch.device("cuda", ctxs[r].device if hasattr(ctxs[r], "device") else 0)
        for step in range(steps):
            obj = VoxelObject.generate(vgs[r], ranges[r])
            comms[r].exchange_halos(obj, ranges)
            local, merged = comms[r].mesh_gather(obj)
            oc, ov = obj.download()
            b, e = ranges[r]
            ref = wc[b * per_plane:e * per_plane]
            # rows of the whole object (its voxels are addressed through its own data_offset)
            assert np.array_equal(oc["kind"], ref["kind"]), f"rank {r} step {step}: chunk kinds"
            nu = ref["kind"] == 2
            assert np.array_equal(oc["flags"][nu], ref["flags"][nu]) and np.array_equal(oc["face"][nu], ref["face"][nu])
            if nu.any():
                gv = ov.reshape(-1, 4096)[oc["data_offset"][nu]]
                rv = wv.reshape(-1, 4096)[ref["data_offset"][nu]]
                assert np.array_equal(gv, rv), f"rank {r} step {step}: voxels"
            if r == 0:
                got = D.merged_mesh_to_numpy(D.PeerComm.merged_to_torch(merged, torch.device("cuda", 0)))
                H.assert_meshes_equal(got, _MeshLike(wm))
            obj.free()

    errors = _run(world, rank)
    for c in comms:
        c.close()
    for r, e in enumerate(errors):
        assert e is None, f"rank {r}: {e}"


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_mesh_parts_concatenate_to_the_whole_mesh(world):
    # ivx_object_mesh_distributed: the parts stay on their ranks, rebased in place; laid end to end at their bases they
    # are the whole object's mesh. Gathered and distributed steps alternate on the same communicator (epochs, parities).
    make, types = GRAPHS["zoo"]
    ctxs, vgs, whole, wi, wmesh, comms = _setup(make, types, world)
    wm = wmesh.download()
    ranges = D.slab_ranges(wi["chunk_counts"][0], world)
    steps = 5
    parts = [[None] * world for _ in range(steps)]

    def rank(r):
        for step in range(steps):
            obj = VoxelObject.generate(vgs[r], ranges[r])
            comms[r].exchange_halos(obj, ranges)
            if step % 2 == 0:
                local, bases = comms[r].mesh_distributed(obj)
                parts[step][r] = (local.download(), bases)
            else:
                local, merged = comms[r].mesh_gather(obj)
                if r == 0:
                    got = D.merged_mesh_to_numpy(D.PeerComm.merged_to_torch(merged, torch.device("cuda", 0)))
                    H.assert_meshes_equal(got, _MeshLike(wm))
            obj.free()

    errors = _run(world, rank)
    for c in comms:
        c.close()
    for r, e in enumerate(errors):
        assert e is None, f"rank {r}: {e}"
    for step in range(0, steps, 2):
        v = i = s = 0
        for r in range(world):
            m, bases = parts[step][r]
            assert bases == (v, i, s), (step, r, bases, (v, i, s))
            v, i, s = v + len(m["positions"]), i + len(m["indices"]), s + len(m["submeshes"])
        cat = {k: np.concatenate([parts[step][r][0][k] for r in range(world)]) for k in wm}
        H.assert_meshes_equal(cat, _MeshLike(wm))


def test_comm_reports_a_merged_mesh_that_does_not_fit():
    make, types = GRAPHS["zoo"]
    world = 2
    ctxs, vgs, whole, wi, wmesh, comms = _setup(make, types, world, capacity_scale=0.5)
    ranges = D.slab_ranges(wi["chunk_counts"][0], world)
    raised = [False] * world

    def rank(r):
        obj = VoxelObject.generate(vgs[r], ranges[r])
        comms[r].exchange_halos(obj, ranges)
        try:
            comms[r].mesh_gather(obj)
        except Exception as e:  # noqa: BLE001
            assert "IVX_ERR_CAPACITY" in str(e), str(e)
            raised[r] = True

    errors = _run(world, rank)
    for c in comms:
        c.close()
    assert all(e is None for e in errors), errors
    assert raised[0] and raised[world - 1]  # the gather rank and the rank whose part did not fit
