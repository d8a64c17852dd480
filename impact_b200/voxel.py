"""Host-side mirror of the reference's voxel object API over the C ABI.

Names follow the reference (engine/crates/impact_voxel/src):
  SDFGraph.build_in             → Context.build_generator       (generation/sdf/atomic.rs:1031)
  SDFVoxelGenerator::new        → SDFVoxelGenerator             (generation.rs:207)
  VoxelObject::generate         → VoxelObject.generate          (object.rs:239)
  VoxelObjectMesh::create       → VoxelObjectMesh.create        (mesh.rs:280)
  sync_with_voxel_object        → VoxelObjectMesh.sync_with_voxel_object (mesh.rs:360)
  apply_sphere_absorption       → VoxelObject.absorb_sphere     (interaction/absorption.rs:801)
  apply_capsule_absorption      → VoxelObject.absorb_capsule    (interaction/absorption.rs:846)
All compute happens in libimpact_voxel_cuda.so on the GPU; this module only
marshals POD buffers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .graph import PROG_NODE_DTYPE, SDFGraph, VoxelTypeGenerator


class Context:
    """One `ivx_ctx`: device, stream and device-memory pool."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = L.lib()
        self.device = device
        self.stream_handle = stream  # the caller's stream the library works on (None: a stream of its own)
        cfg = L.Config(self._lib.ivx_abi_version(), device, C.c_void_p(stream) if stream else None, 0)
        h = C.c_void_p()
        rc = self._lib.ivx_create(C.byref(cfg), C.byref(h))
        if rc != L.IVX_OK:
            raise L.IvxError(rc, "ivx_create failed (a CUDA device is required; there is no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._lib.ivx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def check(self, rc: int):
        if rc != L.IVX_OK:
            raise L.IvxError(rc, self._lib.ivx_last_error(self.h).decode())

    @property
    def kernel_launch_count(self) -> int:
        return int(self._lib.ivx_kernel_launch_count(self.h))

    def synchronize(self):
        self.check(self._lib.ivx_synchronize(self.h))

    KERNEL_IDS = {"fold_conservative": 0, "fold_exact": 1, "eval": 2, "boundary": 3, "mesh_count": 4,
                  "mesh_emit": 5, "absorb": 6, "types": 7, "moments_rows": 8, "moments_non_uniform": 9, "moments_sum": 10}

    def profile_enable(self, enabled: bool = True):
        self.check(self._lib.ivx_profile_enable(self.h, C.c_int(1 if enabled else 0)))

    def profile_reset(self):
        self.check(self._lib.ivx_profile_reset(self.h))

    def profile_get(self) -> dict:
        """→ {kernel name: (total device ms, launches)} measured with CUDA events on the ctx stream."""
        out = {}
        for name, kid in self.KERNEL_IDS.items():
            ms, n = C.c_double(), C.c_uint64()
            self.check(self._lib.ivx_profile_get(self.h, C.c_uint32(kid), C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def profile_counter(self, counter_id: int = 0) -> int:
        """Device-side work counters while profiling is on: 0 = 4-D simplex evaluations of the voxel type kernel."""
        v = C.c_uint64()
        self.check(self._lib.ivx_profile_counter(self.h, C.c_uint32(counter_id), C.byref(v)))
        return int(v.value)

    def build_generator(self, graph: SDFGraph) -> "SDFGenerator":
        """`SDFGraph::build_in` → `SDFGenerator::new_in` (atomic.rs:1031-1037, 228-493)."""
        return SDFGenerator.from_graph(self, graph.nodes(), graph.root_node_id)


class SDFGenerator:
    """`SDFGenerator` (atomic.rs:33-40): the compiled post-order program, resident on the device."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle
        info = L.ProgramInfo()
        ctx.check(ctx._lib.ivx_program_info_get(ctx.h, self.h, C.byref(info)))
        self.node_count, self.stack_depth = info.node_count, info.stack_depth
        self.domain = (np.array(info.domain_lo[:], np.float32), np.array(info.domain_hi[:], np.float32))

    @classmethod
    def from_graph(cls, ctx: Context, nodes: np.ndarray, root: int) -> "SDFGenerator":
        nodes = np.ascontiguousarray(nodes)
        h = C.c_void_p()
        ctx.check(ctx._lib.ivx_program_build(ctx.h, L.ptr(nodes), C.c_uint32(len(nodes)), C.c_uint32(root), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_processed_nodes(cls, ctx: Context, nodes: np.ndarray, stack_depth: int, domain_lo, domain_hi):
        nodes = np.ascontiguousarray(nodes)
        lo = np.asarray(domain_lo, np.float32)
        hi = np.asarray(domain_hi, np.float32)
        h = C.c_void_p()
        ctx.check(ctx._lib.ivx_program_upload(ctx.h, L.ptr(nodes), C.c_uint32(len(nodes)), C.c_uint32(stack_depth),
                                              L.ptr(lo), L.ptr(hi), C.byref(h)))
        return cls(ctx, h)

    def nodes(self) -> np.ndarray:
        out = np.zeros(self.node_count, PROG_NODE_DTYPE)
        self.ctx.check(self.ctx._lib.ivx_program_nodes(self.ctx.h, self.h, L.ptr(out), C.c_uint32(len(out))))
        return out

    def compute_signed_distances_for_chunks(self, chunk_origins_in_root_space) -> np.ndarray:
        """Batched `compute_signed_distances_for_chunk` (atomic.rs:207-216) → (n, 4096) f32."""
        org = np.ascontiguousarray(chunk_origins_in_root_space, np.float32).reshape(-1, 3)
        out = np.zeros((len(org), 4096), np.float32)
        self.ctx.check(self.ctx._lib.ivx_program_eval_chunks(self.ctx.h, self.h, L.ptr(org), C.c_uint32(len(org)), L.ptr(out)))
        return out

    def compute_signed_distances_for_blocks_preserving_gradients(self, block_origins, size: int) -> np.ndarray:
        """Batched `compute_signed_distances_for_block_preserving_gradients::<SIZE, _>` (atomic.rs:877-998),
        SIZE in {1, 2} → (n, SIZE**3) f32; the meta compiler's surface probes."""
        org = np.ascontiguousarray(block_origins, np.float32).reshape(-1, 3)
        out = np.zeros((len(org), size ** 3), np.float32)
        self.ctx.check(self.ctx._lib.ivx_program_eval_blocks(self.ctx.h, self.h, L.ptr(org), C.c_uint32(len(org)),
                                                             C.c_uint32(size), L.ptr(out)))
        return out

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx._lib.ivx_program_free(self.ctx.h, self.h)
            self.h = None


class SDFVoxelGenerator:
    """`SDFVoxelGenerator` (generation.rs:70-77): voxel extent + SDF generator + voxel type generator."""

    def __init__(self, voxel_extent: float, sdf_generator: SDFGenerator, voxel_type_generator: VoxelTypeGenerator):
        assert voxel_extent > 0.0
        self.voxel_extent = voxel_extent
        self.sdf_generator = sdf_generator
        self.voxel_type_generator = voxel_type_generator


def plane_work(generator: SDFVoxelGenerator) -> np.ndarray:
    """Estimated generation work per chunk plane (ivx_program_plane_work) for `distributed.slab_ranges_weighted`."""
    cached = getattr(generator, "_plane_work", None)
    if cached is not None:  # a property of the compiled program, the extent and the type generator, all immutable
        return cached
    ctx = generator.sdf_generator.ctx
    tg = generator.voxel_type_generator.pod()
    n = C.c_uint32()
    ctx.check(ctx._lib.ivx_program_plane_work(ctx.h, generator.sdf_generator.h, C.c_float(generator.voxel_extent), L.ptr(tg),
                                              None, C.c_uint32(0), C.byref(n)))
    out = np.zeros(max(1, n.value), np.uint32)
    ctx.check(ctx._lib.ivx_program_plane_work(ctx.h, generator.sdf_generator.h, C.c_float(generator.voxel_extent), L.ptr(tg),
                                              L.ptr(out), C.c_uint32(len(out)), C.byref(n)))
    generator._plane_work = out[: n.value]
    return generator._plane_work


class VoxelObject:
    """`VoxelObject` (object.rs:45-57), device resident."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    @classmethod
    def generate(cls, generator: SDFVoxelGenerator, chunk_i_range=None) -> "VoxelObject":
        """`VoxelObject::generate` (object.rs:239-244).

        `chunk_i_range` = this rank's x-slab of chunk planes (multi-GPU): the object's cross-chunk derived state
        then stays pending until `impact_b200.distributed.exchange_halos_and_finalize` (or the slab protocol
        calls below) has run."""
        ctx = generator.sdf_generator.ctx
        tg = generator.voxel_type_generator.pod()
        h = C.c_void_p()
        if chunk_i_range is None:
            ctx.check(ctx._lib.ivx_object_generate(ctx.h, generator.sdf_generator.h, C.c_float(generator.voxel_extent),
                                                   L.ptr(tg), C.byref(h)))
        else:
            ctx.check(ctx._lib.ivx_object_generate_slab(ctx.h, generator.sdf_generator.h, C.c_float(generator.voxel_extent),
                                                        L.ptr(tg), C.c_uint32(chunk_i_range[0]),
                                                        C.c_uint32(chunk_i_range[1]), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def generate_streamed(cls, generator: SDFVoxelGenerator, host_chunks: np.ndarray, host_voxels: np.ndarray):
        """`VoxelObject::generate` straight into host buffers (`ivx_object_generate_streamed`): the download of
        finished chunk planes overlaps the generation of the following ones. Returns (object, n_non_uniform); the
        buffers are complete after `ctx.synchronize()`."""
        ctx = generator.sdf_generator.ctx
        tg = generator.voxel_type_generator.pod()
        h = C.c_void_p()
        nnu = C.c_uint64()
        ctx.check(ctx._lib.ivx_object_generate_streamed(
            ctx.h, generator.sdf_generator.h, C.c_float(generator.voxel_extent), L.ptr(tg), L.ptr(host_chunks),
            C.c_size_t(host_chunks.nbytes // 16), L.ptr(host_voxels), C.c_size_t(host_voxels.nbytes // 3), C.byref(h),
            C.byref(nnu)))
        return cls(ctx, h), int(nnu.value)

    @classmethod
    def from_generated_chunks(cls, ctx: Context, voxel_extent: float, grid_shape, voxels: np.ndarray,
                              sparseness: np.ndarray) -> "VoxelObject":
        """`VoxelObject::generate` for a host-side `ChunkedVoxelGenerator` given as data (object.rs:361-404):
        `voxels` = (n_chunks, 4096) `Voxel`s in linear chunk order, `sparseness` = per chunk bit 0
        has_only_empty_voxels, bit 1 is_void. Classification and all derived state are computed on the device."""
        gs = np.asarray(grid_shape, np.uint32)
        voxels = np.ascontiguousarray(voxels, L.VOXEL_DTYPE)
        sparseness = np.ascontiguousarray(sparseness, np.uint8)
        n = int(np.prod((gs + 15) // 16))
        assert voxels.size == n * 4096 and sparseness.size == n
        h = C.c_void_p()
        ctx.check(ctx._lib.ivx_object_from_generated_chunks(ctx.h, C.c_float(voxel_extent), L.ptr(gs), L.ptr(voxels),
                                                            L.ptr(sparseness), C.byref(h)))
        return cls(ctx, h)

    def free(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx._lib.ivx_object_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        self.free()

    def info(self) -> dict:
        i = L.ObjectInfo()
        self.ctx.check(self.ctx._lib.ivx_object_info_get(self.ctx.h, self.h, C.byref(i)))
        return {
            "voxel_extent": i.voxel_extent, "grid_shape": tuple(i.grid_shape[:]), "chunk_counts": tuple(i.chunk_counts[:]),
            "chunk_i_begin": i.chunk_i_begin, "chunk_i_end": i.chunk_i_end, "n_void": i.n_void, "n_uniform": i.n_uniform,
            "n_non_uniform": i.n_non_uniform,
            "occupied_chunk_ranges": np.array(i.occupied_chunk_ranges[:], np.uint32).reshape(3, 2),
            "occupied_voxel_ranges": np.array(i.occupied_voxel_ranges[:], np.uint32).reshape(3, 2),
            "device_bytes": i.device_bytes,
        }

    def download(self, with_voxels: bool = True):
        """→ (chunks[C], voxels[4096 * n_non_uniform]) in the reference's layout."""
        inf = self.info()
        n = (inf["chunk_i_end"] - inf["chunk_i_begin"]) * inf["chunk_counts"][1] * inf["chunk_counts"][2]
        chunks = np.zeros(n, L.CHUNK_DTYPE)
        voxels = np.zeros(inf["n_non_uniform"] * 4096 if with_voxels else 0, L.VOXEL_DTYPE)
        self.ctx.check(self.ctx._lib.ivx_object_download(
            self.ctx.h, self.h, L.ptr(chunks), C.c_size_t(n), L.ptr(voxels) if with_voxels else None,
            C.c_size_t(len(voxels))))
        return chunks, voxels

    # ---- slab protocol (multi-GPU; include/impact_voxel_cuda.h "slab protocol") ----
    # Buffers are raw DEVICE pointers (int); impact_b200.distributed wraps them in torch tensors for NCCL.
    def plane_chunks(self) -> int:
        cc = self.info()["chunk_counts"]
        return int(cc[1]) * int(cc[2])

    def halo_capacity(self) -> int:
        n = C.c_size_t()
        self.ctx.check(self.ctx._lib.ivx_object_halo_capacity(self.ctx.h, self.h, C.byref(n)))
        return n.value

    def halo_export(self, side: int, device_ptr: int, capacity: int) -> int:
        n = C.c_size_t()
        self.ctx.check(self.ctx._lib.ivx_object_halo_export(self.ctx.h, self.h, C.c_int(side), C.c_void_p(device_ptr),
                                                            C.c_size_t(capacity), C.byref(n)))
        return n.value

    def halo_import(self, side: int, device_ptr: int, nbytes: int) -> None:
        self.ctx.check(self.ctx._lib.ivx_object_halo_import(self.ctx.h, self.h, C.c_int(side), C.c_void_p(device_ptr),
                                                            C.c_size_t(nbytes)))

    def slab_classify(self) -> None:
        self.ctx.check(self.ctx._lib.ivx_object_slab_classify(self.ctx.h, self.h))

    def halo_kinds_export(self, side: int, device_ptr: int, capacity: int) -> None:
        self.ctx.check(self.ctx._lib.ivx_object_halo_kinds_export(self.ctx.h, self.h, C.c_int(side), C.c_void_p(device_ptr),
                                                                  C.c_size_t(capacity)))

    def halo_kinds_import(self, side: int, device_ptr: int, nbytes: int) -> None:
        self.ctx.check(self.ctx._lib.ivx_object_halo_kinds_import(self.ctx.h, self.h, C.c_int(side), C.c_void_p(device_ptr),
                                                                  C.c_size_t(nbytes)))

    def slab_finalize(self) -> None:
        self.ctx.check(self.ctx._lib.ivx_object_slab_finalize(self.ctx.h, self.h))

    def absorb_sphere(self, center, radius: float, influence_radius: float) -> dict:
        """`apply_sphere_absorption` (absorption.rs:801-844) in normalized voxel space."""
        c = np.asarray(center, np.float32)
        st = L.AbsorbStats()
        self.ctx.check(self.ctx._lib.ivx_object_absorb_sphere(self.ctx.h, self.h, L.ptr(c), C.c_float(radius),
                                                              C.c_float(influence_radius), C.byref(st)))
        return {f: getattr(st, f) for f, _ in L.AbsorbStats._fields_}

    def absorb_capsule(self, segment_start, segment_vector, radius: float, influence_radius: float) -> dict:
        """`apply_capsule_absorption` (absorption.rs:846-889) in normalized voxel space."""
        a, v = np.asarray(segment_start, np.float32), np.asarray(segment_vector, np.float32)
        st = L.AbsorbStats()
        self.ctx.check(self.ctx._lib.ivx_object_absorb_capsule(self.ctx.h, self.h, L.ptr(a), L.ptr(v), C.c_float(radius),
                                                               C.c_float(influence_radius), C.byref(st)))
        return {f: getattr(st, f) for f, _ in L.AbsorbStats._fields_}

    def inertial_moments(self, voxel_type_densities, initial=None, per_chunk: bool = False):
        """`compute_inertial_property_moments_for_object` (object/inertia.rs:754-789) → 10 f32: mass, moments[3],
        moments_of_inertia[3], products_of_inertia[3] about the grid origin, bit for bit the reference's sums.
        `initial` continues another slab's sums (multi-GPU); `per_chunk` also returns the (n_owned_chunks, 10) terms."""
        dens = np.ascontiguousarray(voxel_type_densities, np.float32)
        out = np.zeros(10, np.float32)
        init = None if initial is None else np.ascontiguousarray(initial, np.float32)
        assert init is None or init.shape == (10,)
        pc = None
        if per_chunk:
            inf = self.info()
            n = (inf["chunk_i_end"] - inf["chunk_i_begin"]) * inf["chunk_counts"][1] * inf["chunk_counts"][2]
            pc = np.zeros((n, 10), np.float32)
        self.ctx.check(self.ctx._lib.ivx_object_inertial_moments(
            self.ctx.h, self.h, L.ptr(dens), C.c_uint32(len(dens)), L.ptr(init) if init is not None else None,
            L.ptr(out), L.ptr(pc) if per_chunk else None, C.c_size_t(len(pc) if per_chunk else 0)))
        return (out, pc) if per_chunk else out

    def absorb_sphere_inertial(self, center, radius: float, influence_radius: float, voxel_type_densities,
                               moments: np.ndarray) -> dict:
        """`apply_sphere_absorption` with its `VoxelObjectInertialPropertyUpdater` (absorption.rs:801-844): `moments`
        (10 f32, e.g. `VoxelObjectInertialPropertyManager.m`) is updated in place, bit for bit like the reference."""
        c = np.asarray(center, np.float32)
        dens = np.ascontiguousarray(voxel_type_densities, np.float32)
        assert moments.dtype == np.float32 and moments.shape == (10,) and moments.flags.c_contiguous
        st = L.AbsorbStats()
        self.ctx.check(self.ctx._lib.ivx_object_absorb_sphere_inertial(
            self.ctx.h, self.h, L.ptr(c), C.c_float(radius), C.c_float(influence_radius), L.ptr(dens),
            C.c_uint32(len(dens)), L.ptr(moments), C.byref(st)))
        return {f: getattr(st, f) for f, _ in L.AbsorbStats._fields_}

    def absorb_capsule_inertial(self, segment_start, segment_vector, radius: float, influence_radius: float,
                                voxel_type_densities, moments: np.ndarray) -> dict:
        """`apply_capsule_absorption` with its inertial-property updater (absorption.rs:846-889)."""
        a, v = np.asarray(segment_start, np.float32), np.asarray(segment_vector, np.float32)
        dens = np.ascontiguousarray(voxel_type_densities, np.float32)
        assert moments.dtype == np.float32 and moments.shape == (10,) and moments.flags.c_contiguous
        st = L.AbsorbStats()
        self.ctx.check(self.ctx._lib.ivx_object_absorb_capsule_inertial(
            self.ctx.h, self.h, L.ptr(a), L.ptr(v), C.c_float(radius), C.c_float(influence_radius), L.ptr(dens),
            C.c_uint32(len(dens)), L.ptr(moments), C.byref(st)))
        return {f: getattr(st, f) for f, _ in L.AbsorbStats._fields_}

    def sphere_contacts(self, rotation_xyzw, translation, center, radius: float) -> np.ndarray:
        """`for_each_sphere_voxel_object_contact` (collidable.rs:1097-1127): `transform_to_object_space` as (unit
        quaternion, translation) and the sphere in the space it starts from → contacts in the closure's call order."""
        iso = np.concatenate([np.asarray(rotation_xyzw, np.float32), np.asarray(translation, np.float32)]).astype(np.float32)
        ctr = np.asarray(center, np.float32)
        return self._surface_query(lambda o, c, n: self.ctx._lib.ivx_object_sphere_contacts(
            self.ctx.h, self.h, L.ptr(iso), L.ptr(ctr), C.c_float(radius), o, c, n), L.CONTACT_DTYPE)

    def plane_contacts(self, rotation_xyzw, translation, unit_normal, displacement: float) -> np.ndarray:
        """`for_each_voxel_object_plane_contact` (collidable.rs:1176-1209): contacts of the corner voxels with the plane
        { x : unit_normal . x = displacement } given in the space `transform_to_object_space` starts from."""
        iso = np.concatenate([np.asarray(rotation_xyzw, np.float32), np.asarray(translation, np.float32)]).astype(np.float32)
        nrm = np.asarray(unit_normal, np.float32)
        return self._surface_query(lambda o, c, n: self.ctx._lib.ivx_object_plane_contacts(
            self.ctx.h, self.h, L.ptr(iso), L.ptr(nrm), C.c_float(displacement), o, c, n), L.CONTACT_DTYPE)

    def capsule_contacts(self, rotation_xyzw, translation, segment_start, segment_vector, radius: float) -> np.ndarray:
        """`for_each_capsule_voxel_object_contact` (collidable.rs:1257-1288)."""
        iso = np.concatenate([np.asarray(rotation_xyzw, np.float32), np.asarray(translation, np.float32)]).astype(np.float32)
        a, v = np.asarray(segment_start, np.float32), np.asarray(segment_vector, np.float32)
        return self._surface_query(lambda o, c, n: self.ctx._lib.ivx_object_capsule_contacts(
            self.ctx.h, self.h, L.ptr(iso), L.ptr(a), L.ptr(v), C.c_float(radius), o, c, n), L.CONTACT_DTYPE)

    def _surface_query(self, call, dtype=None) -> np.ndarray:
        n = C.c_uint64()
        cap = 1 << 14
        while True:
            out = np.zeros(cap, dtype or L.SURFACE_VOXEL_DTYPE)
            rc = call(L.ptr(out), C.c_size_t(cap), C.byref(n))
            if rc == 5:  # IVX_ERR_CAPACITY: n holds the number found
                cap = int(n.value)
                continue
            self.ctx.check(rc)
            return out[: n.value]

    def surface_voxels_in_ranges(self, ranges=None) -> np.ndarray:
        """`for_each_surface_voxel_in_voxel_ranges` (intersection.rs:97-151) as an array in the closure's call order;
        `ranges` (3 x 2) defaults to the occupied voxel ranges = `for_each_surface_voxel`."""
        r = np.ascontiguousarray(self.info()["occupied_voxel_ranges"] if ranges is None else ranges, np.uint32).reshape(6)
        return self._surface_query(lambda o, c, n: self.ctx._lib.ivx_object_surface_voxels_in_ranges(self.ctx.h, self.h, L.ptr(r), o, c, n))

    def surface_voxels_touching_sphere(self, center, radius: float) -> np.ndarray:
        """`for_each_surface_voxel_maybe_intersecting_sphere` (intersection.rs:51-71), sphere in normalized voxel space."""
        ctr = np.asarray(center, np.float32)
        return self._surface_query(lambda o, c, n: self.ctx._lib.ivx_object_surface_voxels_touching_sphere(
            self.ctx.h, self.h, L.ptr(ctr), C.c_float(radius), o, c, n))

    def surface_voxels_touching_capsule(self, segment_start, segment_vector, radius: float) -> np.ndarray:
        """`for_each_surface_voxel_maybe_intersecting_capsule` (intersection.rs:73-85)."""
        a, v = np.asarray(segment_start, np.float32), np.asarray(segment_vector, np.float32)
        return self._surface_query(lambda o, c, n: self.ctx._lib.ivx_object_surface_voxels_touching_capsule(
            self.ctx.h, self.h, L.ptr(a), L.ptr(v), C.c_float(radius), o, c, n))

    def surface_voxels_within_plane(self, unit_normal, displacement: float) -> np.ndarray:
        """`for_each_surface_voxel_maybe_intersecting_negative_halfspace_of_plane` (intersection.rs:30-40), plane in
        normalized voxel space."""
        nrm = np.asarray(unit_normal, np.float32)
        return self._surface_query(lambda o, c, n: self.ctx._lib.ivx_object_surface_voxels_within_plane(
            self.ctx.h, self.h, L.ptr(nrm), C.c_float(displacement), o, c, n))

    def extract_any_disconnected_region(self):
        """`VoxelObject::extract_any_disconnected_region` (extraction.rs:78-113) → (info dict, extracted VoxelObject or
        None). This object is modified in place."""
        info = L.ExtractionInfo()
        h = C.c_void_p()
        self.ctx.check(self.ctx._lib.ivx_object_extract_disconnected_region(self.ctx.h, self.h, C.byref(info), C.byref(h)))
        d = {f: getattr(info, f) for f, _ in L.ExtractionInfo._fields_ if f != "origin_offset_in_parent"}
        d["origin_offset_in_parent"] = tuple(info.origin_offset_in_parent[:])
        return d, (VoxelObject(self.ctx, h) if h.value else None)

    def resolve_connected_regions(self, download: bool = False) -> dict:
        """`update_local_connected_regions_for_all_chunks` + `resolve_connected_regions_between_all_chunks` +
        `count_regions` / `find_two_disconnected_regions` (split_detection.rs:193-488) and the region
        `extract_smallest_region_with_property_transferrer` would pick (extraction.rs:121-281)."""
        si = L.SplitInfo()
        self.ctx.check(self.ctx._lib.ivx_object_resolve_connected_regions(self.ctx.h, self.h, C.byref(si)))
        cand = [{"label": int(c.label), "chunk_count": int(c.chunk_count),
                 "non_uniform_chunk_count": int(c.non_uniform_chunk_count),
                 "chunk_min": np.array(c.chunk_min[:], np.uint32), "chunk_max": np.array(c.chunk_max[:], np.uint32)}
                for c in si.candidates]
        out = {"n_regions": int(si.n_regions), "has_two": bool(si.has_two), "two": (cand[0]["label"], cand[1]["label"]),
               "smallest": int(si.smallest), "candidates": cand, "n_local_regions": int(si.n_local_regions),
               "n_connections": int(si.n_connections),
               "n_relabelled_chunks": int(si.n_relabelled_chunks), "device_ms": float(si.device_ms), "host_ms": float(si.host_ms)}
        if download:
            inf = self.info()
            n = int(np.prod(inf["chunk_counts"]))
            labels = np.zeros(inf["n_non_uniform"] * 4096, np.uint8)
            per_chunk = np.zeros(n, L.CHUNK_REGIONS_DTYPE)
            roots = np.zeros(max(1, si.n_local_regions), np.uint32)
            self.ctx.check(self.ctx._lib.ivx_object_split_detection_download(
                self.ctx.h, self.h, L.ptr(labels), C.c_size_t(len(labels)), L.ptr(per_chunk), C.c_size_t(n), L.ptr(roots),
                C.c_size_t(len(roots))))
            out.update(voxel_labels=labels, per_chunk=per_chunk, region_roots=roots[: si.n_local_regions])
        return out

    def count_regions(self) -> int:
        return self.resolve_connected_regions()["n_regions"]

    def invalidated_mesh_chunk_indices(self) -> np.ndarray:
        cnt = C.c_uint32()
        self.ctx.check(self.ctx._lib.ivx_object_dirty_chunks(self.ctx.h, self.h, None, C.c_uint32(0), C.byref(cnt)))
        out = np.zeros(max(1, cnt.value), np.uint32)
        self.ctx.check(self.ctx._lib.ivx_object_dirty_chunks(self.ctx.h, self.h, L.ptr(out), C.c_uint32(len(out)), C.byref(cnt)))
        return out[: cnt.value]


def voxel_ranges_within_plane(occupied_voxel_ranges, unit_normal, displacement: float) -> np.ndarray:
    """`voxel_ranges_within_plane` (object/intersection.rs:751-761), host only → 3 x 2 voxel ranges."""
    occ = np.ascontiguousarray(occupied_voxel_ranges, np.uint32).reshape(6)
    nrm = np.asarray(unit_normal, np.float32)
    out = np.zeros(6, np.uint32)
    rc = L.lib().ivx_voxel_ranges_within_plane(L.ptr(occ), L.ptr(nrm), C.c_float(displacement), L.ptr(out))
    if rc != L.IVX_OK:
        raise L.IvxError(rc, "ivx_voxel_ranges_within_plane")
    return out.reshape(3, 2)


def box_intersection_bounds(a_lower, a_upper, b_center, b_orientation_xyzw, b_half_extents):
    """`compute_box_intersection_bounds` (impact_geometry/src/oriented_box.rs:315-431), host only → None or
    ((lower, upper) in A's frame, (lower, upper) in B's frame relative to B's centre)."""
    f = lambda v: np.ascontiguousarray(v, np.float32)
    ia, ib = np.zeros(6, np.float32), np.zeros(6, np.float32)
    hit = C.c_int()
    rc = L.lib().ivx_box_intersection_bounds(L.ptr(f(a_lower)), L.ptr(f(a_upper)), L.ptr(f(b_center)),
                                             L.ptr(f(b_orientation_xyzw)), L.ptr(f(b_half_extents)), L.ptr(ia), L.ptr(ib),
                                             C.byref(hit))
    if rc != L.IVX_OK:
        raise L.IvxError(rc, "ivx_box_intersection_bounds")
    return ((ia[:3], ia[3:]), (ib[:3], ib[3:])) if hit.value else None


def intersection_voxel_ranges(occupied_a, voxel_extent_a: float, occupied_b, voxel_extent_b: float, rotation_xyzw,
                              translation):
    """`VoxelObject::determine_voxel_ranges_encompassing_intersection` (object/intersection.rs:707-745), host only:
    occupied voxel ranges (3 x 2) of the two objects, their voxel extents and `transform_from_b_to_a` → None or the two
    3 x 2 voxel ranges `absorb_mutually` takes."""
    oa = np.ascontiguousarray(occupied_a, np.uint32).reshape(6)
    ob = np.ascontiguousarray(occupied_b, np.uint32).reshape(6)
    iso = np.concatenate([np.asarray(rotation_xyzw, np.float32), np.asarray(translation, np.float32)]).astype(np.float32)
    ra, rb = np.zeros(6, np.uint32), np.zeros(6, np.uint32)
    hit = C.c_int()
    rc = L.lib().ivx_intersection_voxel_ranges(L.ptr(oa), C.c_float(voxel_extent_a), L.ptr(ob), C.c_float(voxel_extent_b),
                                               L.ptr(iso), L.ptr(ra), L.ptr(rb), C.byref(hit))
    if rc != L.IVX_OK:
        raise L.IvxError(rc, "ivx_intersection_voxel_ranges")
    return (ra.reshape(3, 2), rb.reshape(3, 2)) if hit.value else None


def mutual_contacts(a: VoxelObject, b: VoxelObject, world_to_a, world_to_b, ranges_in_a, ranges_in_b, moments_a, moments_b):
    """`for_each_mutual_voxel_object_contact` (collidable.rs:859-1050) for two objects with collision probes: isometries
    as 7 floats (unit quaternion x, y, z, w; translation), the two 3 x 2 voxel ranges from `intersection_voxel_ranges`, the
    inertial managers' 10-float moment arrays → (contacts of A's probes in B, contacts of B's probes in A)."""
    ctx = a.ctx
    wa, wb = np.ascontiguousarray(world_to_a, np.float32), np.ascontiguousarray(world_to_b, np.float32)
    ra, rb = np.ascontiguousarray(ranges_in_a, np.uint32).reshape(6), np.ascontiguousarray(ranges_in_b, np.uint32).reshape(6)
    ma, mb = np.ascontiguousarray(moments_a, np.float32), np.ascontiguousarray(moments_b, np.float32)
    assert wa.shape == (7,) and wb.shape == (7,) and ma.shape == (10,) and mb.shape == (10,)
    n_ab, n_ba = C.c_uint64(), C.c_uint64()
    capacity = 4096
    while True:
        out = np.zeros(capacity, L.CONTACT_DTYPE)
        rc = ctx._lib.ivx_objects_mutual_contacts(ctx.h, a.h, b.h, L.ptr(wa), L.ptr(wb), L.ptr(ra), L.ptr(rb), L.ptr(ma), L.ptr(mb),
                                                  L.ptr(out), C.c_size_t(capacity), C.byref(n_ab), C.byref(n_ba))
        if rc == 5 and n_ab.value + n_ba.value > capacity:  # IVX_ERR_CAPACITY
            capacity = int(n_ab.value + n_ba.value)
            continue
        ctx.check(rc)
        break
    return out[: n_ab.value].copy(), out[n_ab.value: n_ab.value + n_ba.value].copy()


def absorb_mutually(a: VoxelObject, b: VoxelObject, rotation_xyzw, translation, smoothness: float, ranges_in_a,
                    ranges_in_b, voxel_type_densities=None, moments_a=None, moments_b=None):
    """`apply_mutual_absorption` (interaction/absorption.rs:891-1080): `transform_from_b_to_a` = (unit quaternion
    x, y, z, w; translation), the two 3 x 2 voxel ranges from `determine_voxel_ranges_encompassing_intersection`. With
    densities, the two 10-float moment arrays are updated in place. → (stats_a, stats_b)."""
    ctx = a.ctx
    iso = np.concatenate([np.asarray(rotation_xyzw, np.float32), np.asarray(translation, np.float32)]).astype(np.float32)
    ra = np.ascontiguousarray(ranges_in_a, np.uint32).reshape(6)
    rb = np.ascontiguousarray(ranges_in_b, np.uint32).reshape(6)
    sa, sb = L.AbsorbStats(), L.AbsorbStats()
    dens = None if voxel_type_densities is None else np.ascontiguousarray(voxel_type_densities, np.float32)
    if dens is not None:
        for m in (moments_a, moments_b):
            assert m.dtype == np.float32 and m.shape == (10,) and m.flags.c_contiguous
    ctx.check(ctx._lib.ivx_objects_absorb_mutually(
        ctx.h, a.h, b.h, L.ptr(iso), C.c_float(smoothness), L.ptr(ra), L.ptr(rb),
        L.ptr(dens) if dens is not None else None, C.c_uint32(0 if dens is None else len(dens)),
        L.ptr(moments_a) if dens is not None else None, L.ptr(moments_b) if dens is not None else None,
        C.byref(sa), C.byref(sb)))
    return ({f: getattr(sa, f) for f, _ in L.AbsorbStats._fields_}, {f: getattr(sb, f) for f, _ in L.AbsorbStats._fields_})


class VoxelObjectMesh:
    """`VoxelObjectMesh` (mesh.rs:50-58): SoA buffers + chunk submesh table."""

    def __init__(self, obj: VoxelObject, info: L.MeshInfo):
        self.obj = obj
        self.n_vertices, self.n_indices, self.n_submeshes = info.n_vertices, info.n_indices, info.n_submeshes
        self.n_exposed_chunks = info.n_exposed_chunks
        self.device_info = info

    @classmethod
    def create(cls, obj: VoxelObject) -> "VoxelObjectMesh":
        info = L.MeshInfo()
        obj.ctx.check(obj.ctx._lib.ivx_object_mesh(obj.ctx.h, obj.h, C.byref(info)))
        return cls(obj, info)

    @classmethod
    def sync_with_voxel_object(cls, obj: VoxelObject) -> "VoxelObjectMesh":
        """Re-meshes the invalidated chunks only; returns their meshes as one compact patch."""
        info = L.MeshInfo()
        obj.ctx.check(obj.ctx._lib.ivx_object_remesh_dirty(obj.ctx.h, obj.h, C.byref(info)))
        return cls(obj, info)

    @classmethod
    def sync(cls, obj: VoxelObject) -> "VoxelObjectMesh":
        """`VoxelObjectMesh::sync_with_voxel_object` (mesh.rs:360-456) on the mesh `create` left on the device: the
        invalidated chunks are re-meshed into free ranges of the buffers (or appended), the submesh table follows."""
        info = L.MeshInfo()
        obj.ctx.check(obj.ctx._lib.ivx_object_mesh_sync(obj.ctx.h, obj.h, C.byref(info)))
        return cls(obj, info)

    def modifications(self):
        """`mesh_modifications` → ((n, 4) updated ranges: vertex start, end, index start, end; chunks_were_removed)."""
        ctx = self.obj.ctx
        cnt, removed = C.c_uint64(), C.c_int()
        ctx.check(ctx._lib.ivx_mesh_modifications(ctx.h, self.obj.h, None, C.c_size_t(0), C.byref(cnt), C.byref(removed)))
        out = np.zeros((max(1, cnt.value), 4), np.uint32)
        ctx.check(ctx._lib.ivx_mesh_modifications(ctx.h, self.obj.h, L.ptr(out), C.c_size_t(len(out)), C.byref(cnt),
                                                  C.byref(removed)))
        return out[: cnt.value], bool(removed.value)

    def report_synchronized(self):
        self.obj.ctx.check(self.obj.ctx._lib.ivx_mesh_report_synchronized(self.obj.ctx.h, self.obj.h))

    def collision_probes(self, sync: bool = False) -> dict:
        """`VoxelObjectCollisionProbes` beside this mesh (collidable.rs:346-780): all chunks (`compute_for_all_chunks`,
        MeshedVoxelObject::create) or, with sync=True right after `VoxelObjectMesh.sync`, the chunks that sync visited
        (`sync_with_voxel_object_and_mesh`). → log2_block_size, points (buffer incl. freed ranges), ranges per chunk."""
        ctx = self.obj.ctx
        info = L.ProbesInfo()
        fn = ctx._lib.ivx_object_collision_probes_sync if sync else ctx._lib.ivx_object_collision_probes
        ctx.check(fn(ctx.h, self.obj.h, C.byref(info)))
        points = np.zeros((max(1, info.n_points), 3), np.float32)
        ranges = np.zeros(max(1, info.n_chunks), L.PROBE_RANGE_DTYPE)
        ctx.check(ctx._lib.ivx_collision_probes_download(ctx.h, self.obj.h, L.ptr(points), C.c_size_t(len(points)), L.ptr(ranges),
                                                         C.c_size_t(len(ranges))))
        return {"log2_block_size": int(info.log2_block_size), "points": points[: info.n_points], "ranges": ranges[: info.n_chunks]}

    def download(self) -> dict:
        pos = np.zeros((self.n_vertices, 3), np.float32)
        nrm = np.zeros((self.n_vertices, 3), np.float32)
        im = np.zeros(self.n_indices, L.INDEX_MATERIALS_DTYPE)
        idx = np.zeros(self.n_indices, np.uint32)
        sm = np.zeros(self.n_submeshes, L.SUBMESH_DTYPE)
        vr = np.zeros((self.n_submeshes, 2), np.uint32)
        ctx = self.obj.ctx
        # (the checked form: a handle made before a later create / sync of the same object must not be read with its sizes)
        ctx.check(ctx._lib.ivx_mesh_download_checked(ctx.h, self.obj.h, C.c_uint32(self.n_vertices), C.c_uint32(self.n_indices),
                                                     C.c_uint32(self.n_submeshes), L.ptr(pos), L.ptr(nrm), L.ptr(im), L.ptr(idx),
                                                     L.ptr(sm), L.ptr(vr)))
        return {"positions": pos, "normals": nrm, "index_materials": im, "indices": idx, "submeshes": sm,
                "vertex_ranges": vr}


class VoxelMeshGPUBuffers:
    """`VoxelMeshGPUBuffers` (gpu_resource.rs:460-900): the five buffers a renderer draws a meshed voxel object from, as
    exportable device allocations. `fds[name]` is the POSIX file descriptor of the current allocation behind a buffer
    (import it as external memory; this object closes it when the buffer is re-created and on `close`), `device_ptrs`
    the same buffers for a CUDA consumer in this process."""

    def __init__(self, obj: VoxelObject, handle, info):
        self.obj, self.h = obj, handle
        self.fds = {n: -1 for n in L.MESH_BUFFER_NAMES}
        self._take(info)

    def _take(self, info):
        self.recreated = []
        for b, name in enumerate(L.MESH_BUFFER_NAMES):
            bi = info.buffer[b]
            if bi.fd >= 0:
                if self.fds[name] >= 0:
                    os.close(self.fds[name])
                self.fds[name] = int(bi.fd)
            if bi.recreated:
                self.recreated.append(name)
        self.allocation_bytes = {n: int(info.buffer[b].allocation_bytes) for b, n in enumerate(L.MESH_BUFFER_NAMES)}
        self.valid_bytes = {n: int(info.buffer[b].valid_bytes) for b, n in enumerate(L.MESH_BUFFER_NAMES)}
        self.device_ptrs = {n: int(info.buffer[b].device_ptr or 0) for b, n in enumerate(L.MESH_BUFFER_NAMES)}
        self.n_vertices, self.n_indices, self.n_chunks = int(info.n_vertices), int(info.n_indices), int(info.n_chunks)
        self.bytes_copied, self.n_updated_ranges = int(info.bytes_copied), int(info.n_updated_ranges)

    @classmethod
    def for_voxel_object(cls, obj: VoxelObject) -> "VoxelMeshGPUBuffers":
        """`for_voxel_object` (gpu_resource.rs:484-600): buffers holding the object's current mesh."""
        ctx = obj.ctx
        h, info = C.c_void_p(), L.MeshGpuBuffersInfo()
        ctx.check(ctx._lib.ivx_mesh_gpu_buffers_create(ctx.h, obj.h, C.byref(h), C.byref(info)))
        return cls(obj, h, info)

    def sync_with_voxel_object(self) -> "VoxelMeshGPUBuffers":
        """`sync_with_voxel_object` (gpu_resource.rs:714-900), after `VoxelObjectMesh.sync`: the updated ranges device to
        device; `recreated` names the buffers that outgrew their allocation (new `fds`)."""
        ctx = self.obj.ctx
        info = L.MeshGpuBuffersInfo()
        ctx.check(ctx._lib.ivx_mesh_gpu_buffers_sync(ctx.h, self.obj.h, self.h, C.byref(info)))
        self._take(info)
        return self

    def close(self):
        if getattr(self, "h", None) and getattr(self.obj.ctx, "h", None):
            self.obj.ctx._lib.ivx_mesh_gpu_buffers_destroy(self.obj.ctx.h, self.h)
        self.h = None
        for n, fd in getattr(self, "fds", {}).items():
            if fd >= 0:
                os.close(fd)
            self.fds[n] = -1

    def __del__(self):
        self.close()


def compile_program_host(graph: SDFGraph):
    """Host-only `SDFGraph::build_in` (no device needed): → (nodes, stack_depth, domain_lo, domain_hi).

    Raises ValueError with the reference's message for cycles / missing nodes (atomic.rs:263-268).
    """
    lib = L.lib()
    nodes = np.ascontiguousarray(graph.nodes())
    count = C.c_uint32()
    info = L.ProgramInfo()
    err = C.create_string_buffer(256)
    cap = max(16, 4 * len(nodes))
    while True:
        out = np.zeros(cap, PROG_NODE_DTYPE)
        rc = lib.ivx_program_compile_host(L.ptr(nodes), C.c_uint32(len(nodes)), C.c_uint32(graph.root_node_id), L.ptr(out),
                                          C.c_uint32(cap), C.byref(count), C.byref(info), err, C.c_size_t(256))
        if rc == 5:  # IVX_ERR_CAPACITY
            cap = count.value
            continue
        if rc == 2:
            raise ValueError(err.value.decode())
        if rc != 0:
            raise L.IvxError(rc, "ivx_program_compile_host")
        return (out[: count.value].copy(), info.stack_depth, np.array(info.domain_lo[:], np.float32),
                np.array(info.domain_hi[:], np.float32))
