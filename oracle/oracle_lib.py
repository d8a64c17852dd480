"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/liboracle.so (the CPU restatement of the
reference's voxel hot path). Importable only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg. The product package
(impact_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

VOXEL_DTYPE = np.dtype([("type", "u1"), ("sd", "i1"), ("flags", "u1")])
CHUNK_DTYPE = np.dtype(
    [("kind", "u1"), ("flags", "u1"), ("face", "u1", (6,)), ("uniform_type", "u1"), ("uniform_sd", "i1"),
     ("uniform_flags", "u1"), ("_pad", "u1"), ("data_offset", "<u4")]
)
assert CHUNK_DTYPE.itemsize == 16
SUBMESH_DTYPE = np.dtype(
    [("chunk_indices", "<u4", (3,)), ("index_offset", "<u4"), ("index_count", "<u4"), ("obscured", "<u4", (8,))]
)
assert SUBMESH_DTYPE.itemsize == 52
INDEX_MATERIALS_DTYPE = np.dtype([("indices", "u1", (4,)), ("weights", "u1", (4,))])


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_build_generator.restype = C.c_int
        _lib.orc_generator_node_count.restype = C.c_uint32
        _lib.orc_generator_stack_size.restype = C.c_uint32
        _lib.orc_voxel_generator_create.restype = C.c_void_p
        _lib.orc_object_generate.restype = C.c_void_p
        _lib.orc_object_from_dense.restype = C.c_void_p
        _lib.orc_object_generate_slab.restype = C.c_void_p
        _lib.orc_object_dirty.restype = C.c_uint32
        _lib.orc_mesh_create.restype = C.c_void_p
        _lib.orc_fill_brick.restype = C.c_int
        _lib.orc_mesh_chunk.restype = C.c_int
        for f in ("orc_simplex3", "orc_fbm3", "orc_simplex4", "orc_sd_decode"):
            getattr(_lib, f).restype = C.c_float
        _lib.orc_sd_encode.restype = C.c_int8
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Generator:
    """SDFGenerator (atomic.rs:33-40) built from an atomic node array."""

    def __init__(self, nodes: np.ndarray, root: int):
        from impact_b200.graph import PROG_NODE_DTYPE

        self._prog_dtype = PROG_NODE_DTYPE
        nodes = np.ascontiguousarray(nodes)
        h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = lib().orc_build_generator(_p(nodes), C.c_uint32(len(nodes)), C.c_uint32(root), C.byref(h), err,
                                       C.c_size_t(256))
        if rc != 0:
            raise ValueError(err.value.decode())
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_generator_free(self.h)
            self.h = None

    @property
    def node_count(self) -> int:
        return lib().orc_generator_node_count(self.h)

    @property
    def stack_size(self) -> int:
        return lib().orc_generator_stack_size(self.h)

    def nodes(self) -> np.ndarray:
        out = np.zeros(self.node_count, dtype=self._prog_dtype)
        if len(out):
            lib().orc_generator_nodes(self.h, _p(out))
        return out

    def domain(self):
        lo = np.zeros(3, np.float32)
        hi = np.zeros(3, np.float32)
        lib().orc_generator_domain(self.h, _p(lo), _p(hi))
        return lo, hi

    def eval_chunk(self, lo):
        lo = np.asarray(lo, np.float32)
        out = np.zeros(4096, np.float32)
        dec = np.zeros(max(1, self.node_count), np.uint8)
        lib().orc_eval_chunk(self.h, _p(lo), _p(out), _p(dec))
        return out, dec[: self.node_count]

    def eval_block_preserving_gradients(self, origin, size: int):
        origin = np.asarray(origin, np.float32)
        out = np.zeros(size**3, np.float32)
        lib().orc_eval_block_preserving_gradients(self.h, _p(origin), C.c_int(size), _p(out))
        return out


class VoxelGenerator:
    """SDFVoxelGenerator (generation.rs:70-77)."""

    def __init__(self, gen: Generator, voxel_extent: float, type_gen):
        self.gen = gen
        pod = type_gen.pod()
        self.h = C.c_void_p(lib().orc_voxel_generator_create(gen.h, C.c_float(voxel_extent), _p(pod)))
        gs = np.zeros(3, np.uint32)
        sc = np.zeros(3, np.float32)
        lib().orc_voxel_generator_info(self.h, _p(gs), _p(sc))
        self.grid_shape = tuple(int(x) for x in gs)
        self.shifted_center = sc

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_voxel_generator_free(self.h)
            self.h = None

    def generate_chunk(self, origin):
        origin = np.asarray(origin, np.uint32)
        out = np.zeros(4096, VOXEL_DTYPE)
        sp = np.zeros(2, np.uint8)
        lib().orc_generate_chunk(self.h, _p(origin), _p(out), _p(sp))
        return out, bool(sp[0]), bool(sp[1])


class Object:
    """VoxelObject (object.rs:45-57) minus split detection."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)
        self.t_generate = self.t_derive = 0.0

    @classmethod
    def generate(cls, vg: VoxelGenerator, n_threads: int = 1) -> "Object":
        tg, td = C.c_double(), C.c_double()
        o = cls(lib().orc_object_generate(vg.h, C.c_int(n_threads), C.byref(tg), C.byref(td)))
        o.t_generate, o.t_derive = tg.value, td.value
        return o

    @classmethod
    def generate_slab(cls, vg: VoxelGenerator, plane_begin: int, plane_end: int, n_threads: int = 1) -> "Object":
        """Chunk planes [begin, end) of the object as a stand-alone slab (bench.py's bounded CPU sample)."""
        tg, td = C.c_double(), C.c_double()
        o = cls(lib().orc_object_generate_slab(vg.h, C.c_int(n_threads), C.c_uint32(plane_begin), C.c_uint32(plane_end),
                                               C.byref(tg), C.byref(td)))
        o.t_generate, o.t_derive = tg.value, td.value
        return o

    @classmethod
    def from_dense(cls, sd: np.ndarray, types: np.ndarray, voxel_extent: float = 1.0) -> "Object":
        sd = np.ascontiguousarray(sd, np.int8)
        types = np.ascontiguousarray(types, np.uint8)
        shape = np.asarray(sd.shape, np.uint32)
        return cls(lib().orc_object_from_dense(_p(sd), _p(types), _p(shape), C.c_float(voxel_extent)))

    @classmethod
    def from_generated_chunks(cls, voxels: np.ndarray, sparseness: np.ndarray, grid_shape, voxel_extent: float = 1.0,
                              derive: bool = True) -> "Object":
        """`VoxelObject::generate` (derive=True) / `generate_without_derived_state` for a `ChunkedVoxelGenerator` given
        as data: `voxels` = (n_chunks, 4096) VOXEL_DTYPE in linear chunk order, `sparseness` = per chunk bit 0
        has_only_empty_voxels, bit 1 is_void (generation.rs:41-67, object.rs:361-404)."""
        voxels = np.ascontiguousarray(voxels, VOXEL_DTYPE)
        sparseness = np.ascontiguousarray(sparseness, np.uint8)
        gs = np.asarray(grid_shape, np.uint32)
        lib().orc_object_from_generated_chunks.restype = C.c_void_p
        return cls(lib().orc_object_from_generated_chunks(_p(voxels), _p(sparseness), _p(gs), C.c_float(voxel_extent),
                                                          C.c_int(1 if derive else 0)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_object_free(self.h)
            self.h = None

    def info(self):
        cc = np.zeros(3, np.uint32)
        nv = C.c_uint64()
        oc = np.zeros(6, np.uint32)
        ov = np.zeros(6, np.uint32)
        lib().orc_object_info(self.h, _p(cc), C.byref(nv), _p(oc), _p(ov))
        return {"chunk_counts": tuple(int(x) for x in cc), "n_voxels": nv.value,
                "occupied_chunk_ranges": oc.reshape(3, 2).copy(), "occupied_voxel_ranges": ov.reshape(3, 2).copy()}

    def chunks(self) -> np.ndarray:
        n = int(np.prod(self.info()["chunk_counts"]))
        out = np.zeros(n, CHUNK_DTYPE)
        if n:
            lib().orc_object_chunks(self.h, _p(out))
        return out

    def voxels(self) -> np.ndarray:
        n = self.info()["n_voxels"]
        out = np.zeros(n, VOXEL_DTYPE)
        if n:
            lib().orc_object_voxels(self.h, _p(out))
        return out

    def dirty(self) -> np.ndarray:
        n = lib().orc_object_dirty(self.h, None, C.c_uint32(0))
        out = np.zeros(max(n, 1), np.uint32)
        lib().orc_object_dirty(self.h, _p(out), C.c_uint32(n))
        return out[:n]

    def clear_dirty(self):
        lib().orc_object_clear_dirty(self.h)

    def fill_brick(self, ci, cj, ck):
        values = np.zeros(5832, np.float32)
        types = np.zeros(5832, np.uint8)
        adj = np.zeros(6, np.uint8)
        flags = C.c_uint8()
        ok = lib().orc_fill_brick(self.h, C.c_uint32(ci), C.c_uint32(cj), C.c_uint32(ck), _p(values), _p(types),
                                  _p(adj), C.byref(flags))
        return (values, types, adj, flags.value) if ok else None

    def mesh(self, n_threads: int = 1) -> "Mesh":
        t = C.c_double()
        m = Mesh(lib().orc_mesh_create(self.h, C.c_int(n_threads), C.byref(t)))
        m.t_mesh = t.value
        return m

    def mesh_chunk(self, ci, cj, ck):
        nv, ni = C.c_uint32(), C.c_uint32()
        pos = np.zeros(4913 * 3, np.float32)
        nrm = np.zeros(4913 * 3, np.float32)
        im = np.zeros(4913 * 18, INDEX_MATERIALS_DTYPE)
        idx = np.zeros(4913 * 18, np.uint16)
        flags = C.c_uint8()
        ok = lib().orc_mesh_chunk(self.h, C.c_uint32(ci), C.c_uint32(cj), C.c_uint32(ck), C.byref(nv), C.byref(ni),
                                  _p(pos), _p(nrm), _p(im), _p(idx), C.byref(flags))
        if not ok:
            return None
        v, i = nv.value, ni.value
        return {"positions": pos[: 3 * v].reshape(-1, 3).copy(), "normals": nrm[: 3 * v].reshape(-1, 3).copy(),
                "index_materials": im[:i].copy(), "indices": idx[:i].copy(), "flags": flags.value}

    def absorb_sphere(self, center, radius: float, influence_radius: float):
        center = np.asarray(center, np.float32)
        st = np.zeros(4, np.uint32)
        lib().orc_absorb_sphere(self.h, _p(center), C.c_float(radius), C.c_float(influence_radius), _p(st))
        return {"touched_chunks": int(st[0]), "touched_voxels": int(st[1]), "emptied_voxels": int(st[2]),
                "removed_chunks": int(st[3])}

    def absorb_capsule(self, segment_start, segment_vector, radius: float, influence_radius: float):
        """apply_capsule_absorption (absorption.rs:846-889) with the influence capsule given in voxel space."""
        a, v = np.asarray(segment_start, np.float32), np.asarray(segment_vector, np.float32)
        st = np.zeros(4, np.uint32)
        lib().orc_absorb_capsule(self.h, _p(a), _p(v), C.c_float(radius), C.c_float(influence_radius), _p(st))
        return {"touched_chunks": int(st[0]), "touched_voxels": int(st[1]), "emptied_voxels": int(st[2]),
                "removed_chunks": int(st[3])}

    def inertial_moments(self, densities, per_chunk: bool = False):
        """VoxelObjectInertialPropertyManager::initialized_from (inertia.rs:125-137) → 10 f32: mass, moments[3],
        moments_of_inertia[3], products_of_inertia[3] (and the per-chunk terms, one row per chunk of the grid)."""
        dens = _densities(densities)
        out = np.zeros(10, np.float32)
        pc = np.zeros((len(self.chunks()), 10), np.float32) if per_chunk else None
        lib().orc_object_inertial_moments(self.h, _p(dens), _p(out), _p(pc) if per_chunk else None)
        return (out, pc) if per_chunk else out

    def absorb_sphere_inertial(self, center, radius: float, influence_radius: float, densities, moments):
        """apply_sphere_absorption with its VoxelObjectInertialPropertyUpdater (absorption.rs:801-844): `moments`
        (10 f32) is updated in place, voxel by voxel in the reference's visiting order."""
        center = np.asarray(center, np.float32)
        dens = _densities(densities)
        assert moments.dtype == np.float32 and moments.shape == (10,)
        st = np.zeros(4, np.uint32)
        lib().orc_absorb_sphere_inertial(self.h, _p(center), C.c_float(radius), C.c_float(influence_radius), _p(st),
                                         _p(dens), _p(moments))
        return {"touched_chunks": int(st[0]), "touched_voxels": int(st[1]), "emptied_voxels": int(st[2]),
                "removed_chunks": int(st[3])}

    def absorb_capsule_inertial(self, segment_start, segment_vector, radius: float, influence_radius: float, densities,
                                moments):
        a, v = np.asarray(segment_start, np.float32), np.asarray(segment_vector, np.float32)
        dens = _densities(densities)
        assert moments.dtype == np.float32 and moments.shape == (10,)
        st = np.zeros(4, np.uint32)
        lib().orc_absorb_capsule_inertial(self.h, _p(a), _p(v), C.c_float(radius), C.c_float(influence_radius), _p(st),
                                          _p(dens), _p(moments))
        return {"touched_chunks": int(st[0]), "touched_voxels": int(st[1]), "emptied_voxels": int(st[2]),
                "removed_chunks": int(st[3])}

    SURFACE_VOXEL_DTYPE = np.dtype([("indices", "<u4", (3,)), ("type", "u1"), ("sd", "i1"), ("flags", "u1"), ("placement", "u1")])

    def surface_voxels_in_ranges(self, ranges=None) -> np.ndarray:
        """`for_each_surface_voxel_in_voxel_ranges` (intersection.rs:97-151) in the closure's call order; `ranges` defaults
        to the occupied voxel ranges (`for_each_surface_voxel`)."""
        r = np.ascontiguousarray(self.info()["occupied_voxel_ranges"] if ranges is None else ranges, np.uint32).reshape(6)
        lib().orc_surface_voxels_in_ranges.restype = C.c_uint64
        n = lib().orc_surface_voxels_in_ranges(self.h, _p(r), None, C.c_uint64(0))
        out = np.zeros(max(1, n), self.SURFACE_VOXEL_DTYPE)
        lib().orc_surface_voxels_in_ranges(self.h, _p(r), _p(out), C.c_uint64(n))
        return out[:n]

    CONTACT_DTYPE = np.dtype([("indices", "<u4", (3,)), ("position", "<f4", (3,)), ("normal", "<f4", (3,)), ("depth", "<f4")])

    def sphere_contacts(self, rotation_xyzw, translation, center, radius: float) -> np.ndarray:
        """`for_each_sphere_voxel_object_contact` (collidable.rs:1097-1127): `transform_to_object_space` as (unit
        quaternion, translation), the sphere in the space that transform starts from; contacts in call order."""
        q, t, c = (np.asarray(x, np.float32) for x in (rotation_xyzw, translation, center))
        lib().orc_sphere_contacts.restype = C.c_uint64
        n = lib().orc_sphere_contacts(self.h, _p(q), _p(t), _p(c), C.c_float(radius), None, C.c_uint64(0))
        out = np.zeros(max(1, n), self.CONTACT_DTYPE)
        lib().orc_sphere_contacts(self.h, _p(q), _p(t), _p(c), C.c_float(radius), _p(out), C.c_uint64(n))
        return out[:n]

    def plane_contacts(self, rotation_xyzw, translation, unit_normal, displacement: float) -> np.ndarray:
        """`for_each_voxel_object_plane_contact` (collidable.rs:1176-1209): corner voxels against the plane
        { x : unit_normal . x = displacement } given in the space `transform_to_object_space` starts from."""
        q, t, n = (np.asarray(x, np.float32) for x in (rotation_xyzw, translation, unit_normal))
        lib().orc_plane_contacts.restype = C.c_uint64
        cnt = lib().orc_plane_contacts(self.h, _p(q), _p(t), _p(n), C.c_float(displacement), None, C.c_uint64(0))
        out = np.zeros(max(1, cnt), self.CONTACT_DTYPE)
        lib().orc_plane_contacts(self.h, _p(q), _p(t), _p(n), C.c_float(displacement), _p(out), C.c_uint64(cnt))
        return out[:cnt]

    def capsule_contacts(self, rotation_xyzw, translation, segment_start, segment_vector, radius: float) -> np.ndarray:
        """`for_each_capsule_voxel_object_contact` (collidable.rs:1257-1288): the capsule in the space
        `transform_to_object_space` starts from."""
        q, t, a, v = (np.asarray(x, np.float32) for x in (rotation_xyzw, translation, segment_start, segment_vector))
        lib().orc_capsule_contacts.restype = C.c_uint64
        cnt = lib().orc_capsule_contacts(self.h, _p(q), _p(t), _p(a), _p(v), C.c_float(radius), None, C.c_uint64(0))
        out = np.zeros(max(1, cnt), self.CONTACT_DTYPE)
        lib().orc_capsule_contacts(self.h, _p(q), _p(t), _p(a), _p(v), C.c_float(radius), _p(out), C.c_uint64(cnt))
        return out[:cnt]

    def extract_any_disconnected_region(self):
        """`VoxelObject::extract_any_disconnected_region` (extraction.rs:78-113): → (info, extracted Object or None).
        This object is modified in place (the region's voxels leave it)."""
        info = np.zeros(8, np.uint32)
        lib().orc_extract_any_disconnected_region.restype = C.c_void_p
        h = lib().orc_extract_any_disconnected_region(self.h, _p(info))
        d = {"found_two": bool(info[0]), "extracted": bool(info[1]), "discarded": bool(info[2]),
             "single_chunk": bool(info[3]), "region_label": int(info[4]),
             "origin_offset_in_parent": tuple(int(x) for x in info[5:8])}
        return d, (Object(h) if h else None)

    REGIONS_DTYPE = np.dtype([("region_count", "<u2"), ("boundary_region_count", "<u2"), ("first_region", "<u4")])

    def split_detection(self) -> dict:
        """Connected regions of the object's current state (split_detection.rs): local labels per voxel,
        region counts per chunk, the resolved root of every local region, count_regions,
        find_two_disconnected_regions and the region extract_smallest_region would pick."""
        info = np.zeros(24, np.uint32)
        lib().orc_split_detect.restype = C.c_void_p
        h = C.c_void_p(lib().orc_split_detect(self.h, _p(info)))
        labels = np.zeros(int(info[7]), np.uint8)
        per_chunk = np.zeros(len(self.chunks()), self.REGIONS_DTYPE)
        roots = np.zeros(int(info[6]), np.uint32)
        lib().orc_split_copy(h, _p(labels), _p(per_chunk), _p(roots))
        lib().orc_split_free(h)
        cand = [{"chunk_count": int(info[8 + 8 * q]), "non_uniform_chunk_count": int(info[9 + 8 * q]),
                 "chunk_min": info[10 + 8 * q: 13 + 8 * q].copy(), "chunk_max": info[13 + 8 * q: 16 + 8 * q].copy()}
                for q in range(2)]
        return {"n_regions": int(info[0]), "has_two": bool(info[1]), "two": (int(info[2]), int(info[3])),
                "smallest": int(info[4]), "overflow": bool(info[5]), "voxel_labels": labels, "per_chunk": per_chunk,
                "region_roots": roots, "candidates": cand}

    def count_regions_brute_force(self) -> int:
        lib().orc_count_regions_brute_force.restype = C.c_uint32
        return int(lib().orc_count_regions_brute_force(self.h))


class Mesh:
    """VoxelObjectMesh (mesh.rs:50-58)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)
        self.t_mesh = 0.0
        nv, ni, ns = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().orc_mesh_sizes(self.h, C.byref(nv), C.byref(ni), C.byref(ns))
        self.n_vertices, self.n_indices, self.n_submeshes = nv.value, ni.value, ns.value
        self.positions = np.zeros((self.n_vertices, 3), np.float32)
        self.normals = np.zeros((self.n_vertices, 3), np.float32)
        self.index_materials = np.zeros(self.n_indices, INDEX_MATERIALS_DTYPE)
        self.indices = np.zeros(self.n_indices, np.uint32)
        self.submeshes = np.zeros(self.n_submeshes, SUBMESH_DTYPE)
        self.vertex_ranges = np.zeros((self.n_submeshes, 2), np.uint32)
        lib().orc_mesh_copy(self.h, _p(self.positions), _p(self.normals), _p(self.index_materials),
                            _p(self.indices), _p(self.submeshes), _p(self.vertex_ranges))
        lib().orc_mesh_free(self.h)
        self.h = None


class SyncedMesh:
    """`VoxelObjectMesh` kept in sync with a modified object (mesh.rs:360-456): `create` = recreate, `sync` =
    `sync_with_voxel_object` over the given chunks in the given order. Arrays are read back after every call."""

    def __init__(self, obj: "Object", n_threads: int = 1):
        lib().orc_synced_mesh_create.restype = C.c_void_p
        lib().orc_synced_mesh_mesh.restype = C.c_void_p
        self.h = C.c_void_p(lib().orc_synced_mesh_create(obj.h, C.c_int(n_threads)))
        self._read()

    def _read(self):
        m = C.c_void_p(lib().orc_synced_mesh_mesh(self.h))
        nv, ni, ns = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().orc_mesh_sizes(m, C.byref(nv), C.byref(ni), C.byref(ns))
        self.n_vertices, self.n_indices, self.n_submeshes = nv.value, ni.value, ns.value
        self.positions = np.zeros((self.n_vertices, 3), np.float32)
        self.normals = np.zeros((self.n_vertices, 3), np.float32)
        self.index_materials = np.zeros(self.n_indices, INDEX_MATERIALS_DTYPE)
        self.indices = np.zeros(self.n_indices, np.uint32)
        self.submeshes = np.zeros(self.n_submeshes, SUBMESH_DTYPE)
        self.vertex_ranges = np.zeros((self.n_submeshes, 2), np.uint32)
        lib().orc_mesh_copy(m, _p(self.positions), _p(self.normals), _p(self.index_materials), _p(self.indices),
                            _p(self.submeshes), _p(self.vertex_ranges))

    def sync(self, obj: "Object", dirty_chunks) -> None:
        d = np.ascontiguousarray(dirty_chunks, np.uint32)
        lib().orc_synced_mesh_sync(self.h, obj.h, _p(d), C.c_uint32(len(d)))
        self._read()

    def modifications(self):
        """→ (updated ranges: (n, 4) vertex start, end, index start, end; chunks_were_removed)"""
        lib().orc_synced_mesh_modifications.restype = C.c_uint32
        removed = C.c_int()
        n = lib().orc_synced_mesh_modifications(self.h, None, C.c_uint32(0), C.byref(removed))
        out = np.zeros((max(n, 1), 4), np.uint32)
        lib().orc_synced_mesh_modifications(self.h, _p(out), C.c_uint32(n), C.byref(removed))
        return out[:n], bool(removed.value)

    def report_synchronized(self):
        lib().orc_synced_mesh_report_synchronized(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_synced_mesh_free(self.h)
            self.h = None


class RangeAllocator:
    """`RangeAllocator` (impact_containers/src/range_allocator.rs)."""

    def __init__(self):
        lib().orc_range_allocator_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().orc_range_allocator_create())

    def free_range(self, start, end):
        lib().orc_range_allocator_free_range(self.h, C.c_uint64(start), C.c_uint64(end))

    def allocate_range(self, length):
        s = C.c_uint64()
        ok = lib().orc_range_allocator_allocate(self.h, C.c_uint64(length), C.byref(s))
        return (s.value, s.value + length) if ok else None

    def merge_consecutive_ranges(self):
        lib().orc_range_allocator_merge(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_range_allocator_destroy(self.h)
            self.h = None


def vertex_materials(has_voxel, materials):
    has = np.asarray(has_voxel, np.uint8)
    mat = np.asarray(materials, np.uint8)
    out = np.zeros(16, np.uint8)
    lib().orc_vertex_materials(_p(has), _p(mat), _p(out))
    return out[:8].copy(), out[8:].copy()


def index_materials(vms):
    """vms: three (indices[8], weights[8]) pairs → three (indices[4], weights[4])."""
    buf = np.zeros(48, np.uint8)
    for v, (ind, w) in enumerate(vms):
        buf[16 * v: 16 * v + 8] = ind
        buf[16 * v + 8: 16 * v + 16] = w
    out = np.zeros(24, np.uint8)
    lib().orc_index_materials(_p(buf), _p(out))
    return [(out[8 * v: 8 * v + 4].copy(), out[8 * v + 4: 8 * v + 8].copy()) for v in range(3)]


def absorb_mutually(a: "Object", b: "Object", rotation_xyzw, translation, smoothness: float, ranges_in_a, ranges_in_b,
                    densities=None, moments_a=None, moments_b=None):
    """`apply_mutual_absorption` (absorption.rs:891-1080) with the intersection voxel ranges given (3 x 2 each) and
    `transform_from_b_to_a` = (unit quaternion x, y, z, w; translation). With `densities`, the two 10-float moment arrays
    are updated in place by the inertial-property updaters. → (stats_a, stats_b)."""
    q, t = np.asarray(rotation_xyzw, np.float32), np.asarray(translation, np.float32)
    ra = np.ascontiguousarray(ranges_in_a, np.uint32).reshape(6)
    rb = np.ascontiguousarray(ranges_in_b, np.uint32).reshape(6)
    sa, sb = np.zeros(4, np.uint32), np.zeros(4, np.uint32)
    dens = _densities(densities) if densities is not None else None
    lib().orc_absorb_mutually(a.h, b.h, _p(q), _p(t), C.c_float(smoothness), _p(ra), _p(rb),
                              _p(dens) if dens is not None else None,
                              _p(moments_a) if dens is not None else None, _p(moments_b) if dens is not None else None,
                              _p(sa), _p(sb))
    keys = ("touched_chunks", "touched_voxels", "emptied_voxels", "removed_chunks")
    return dict(zip(keys, map(int, sa))), dict(zip(keys, map(int, sb)))


def voxel_ranges_within_plane(occupied_voxel_ranges, unit_normal, displacement: float) -> np.ndarray:
    """`voxel_ranges_within_plane` (object/intersection.rs:751-761) → 3 x 2."""
    occ = np.ascontiguousarray(occupied_voxel_ranges, np.uint32).reshape(6)
    n = np.asarray(unit_normal, np.float32)
    out = np.zeros(6, np.uint32)
    lib().orc_voxel_ranges_within_plane(_p(occ), _p(n), C.c_float(displacement), _p(out))
    return out.reshape(3, 2)


def _densities(densities) -> np.ndarray:
    """voxel_type_densities padded to 256 entries so that any u8 type indexes it (the reference would panic)."""
    d = np.zeros(256, np.float32)
    d[: len(densities)] = np.asarray(densities, np.float32)
    return d


def moments_for_voxel(voxel_extent: float, densities, ijk, voxel_type: int) -> np.ndarray:
    """compute_moments_for_voxel (inertia.rs:591-625)."""
    out = np.zeros(10, np.float32)
    lib().orc_moments_for_voxel(C.c_float(voxel_extent), _p(_densities(densities)), _p(np.asarray(ijk, np.uint32)),
                                C.c_uint8(voxel_type), _p(out))
    return out


def moments_for_non_uniform_chunk(voxel_extent: float, voxels: np.ndarray, densities, chunk_indices) -> np.ndarray:
    """compute_moments_for_non_uniform_chunk (inertia.rs:629-706); voxels: 4096 x VOXEL_DTYPE."""
    assert voxels.dtype == VOXEL_DTYPE and len(voxels) == 4096
    out = np.zeros(10, np.float32)
    lib().orc_moments_for_non_uniform_chunk(C.c_float(voxel_extent), _p(np.ascontiguousarray(voxels)),
                                            _p(_densities(densities)), _p(np.asarray(chunk_indices, np.uint32)),
                                            _p(out))
    return out


def moments_for_uniform_chunk(voxel_extent: float, densities, voxel_type: int, chunk_indices) -> np.ndarray:
    """compute_moments_for_uniform_chunk (inertia.rs:710-752)."""
    out = np.zeros(10, np.float32)
    lib().orc_moments_for_uniform_chunk(C.c_float(voxel_extent), _p(_densities(densities)), C.c_uint8(voxel_type),
                                        _p(np.asarray(chunk_indices, np.uint32)), _p(out))
    return out


def sd_encode(v: float) -> int:
    return int(lib().orc_sd_encode(C.c_float(v)))


def simplex3(x, y, z, seed):
    return float(lib().orc_simplex3(C.c_float(x), C.c_float(y), C.c_float(z), C.c_int32(seed)))


def fbm3(x, y, z, lac, gain, octaves, seed):
    return float(lib().orc_fbm3(C.c_float(x), C.c_float(y), C.c_float(z), C.c_float(lac), C.c_float(gain),
                                C.c_uint32(octaves), C.c_int32(seed)))


def simplex4(x, y, z, w, seed):
    return float(lib().orc_simplex4(C.c_float(x), C.c_float(y), C.c_float(z), C.c_float(w), C.c_int32(seed)))


class CollisionProbes:
    """`VoxelObjectCollisionProbes` (collidable.rs:346-780): `CollisionProbes(obj, mesh)` = compute_for_all_chunks on a
    `Mesh` or on a `SyncedMesh`'s mesh; `sync` = sync_with_voxel_object_and_mesh over the given chunks in the given order.
    `points` (n, 3) keeps obsolete points in freed ranges; `ranges` = (linear chunk index, start, end) sorted by chunk."""

    def __init__(self, obj: "Object", mesh):
        lib().orc_probes_create.restype = C.c_void_p
        lib().orc_probes_log2_block_size.restype = C.c_uint32
        if isinstance(mesh, SyncedMesh):
            self.h = C.c_void_p(lib().orc_probes_create(obj.h, C.c_void_p(lib().orc_synced_mesh_mesh(mesh.h))))
        else:
            lib().orc_probes_create_arrays.restype = C.c_void_p
            self.h = C.c_void_p(lib().orc_probes_create_arrays(
                obj.h, _p(mesh.positions), _p(mesh.normals), _p(mesh.indices), _p(mesh.submeshes), _p(mesh.vertex_ranges),
                C.c_uint32(mesh.n_vertices), C.c_uint32(mesh.n_indices), C.c_uint32(mesh.n_submeshes)))
        self.log2_block_size = int(lib().orc_probes_log2_block_size(obj.h))
        self._read()

    def _read(self):
        n_points, n_chunks = C.c_uint64(), C.c_uint64()
        lib().orc_probes_sizes(self.h, C.byref(n_points), C.byref(n_chunks))
        self.points = np.zeros((n_points.value, 3), np.float32)
        self.ranges = np.zeros((n_chunks.value, 3), np.uint32)
        lib().orc_probes_copy(self.h, _p(self.points), _p(self.ranges))

    def sync(self, obj: "Object", synced_mesh: "SyncedMesh", dirty_chunks) -> None:
        d = np.ascontiguousarray(dirty_chunks, np.uint32)
        lib().orc_probes_sync(self.h, obj.h, synced_mesh.h, _p(d), C.c_uint32(len(d)))
        self.log2_block_size = int(lib().orc_probes_log2_block_size(obj.h))
        self._read()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_probes_free(self.h)
            self.h = None


def probes_points_for_chunk(log2_block_size, chunk_indices, positions, normals, indices, start_index=0,
                            inverse_voxel_extent=1.0) -> np.ndarray:
    """`add_points_for_vertices_in_blocks` (collidable.rs:614-731) on the given chunk mesh → (n, 3) points."""
    pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, np.uint32)
    ci = np.ascontiguousarray(chunk_indices, np.uint32)
    out = np.zeros((4096, 3), np.float32)
    lib().orc_probes_points_for_chunk.restype = C.c_uint32
    n = lib().orc_probes_points_for_chunk(C.c_uint32(log2_block_size), _p(ci), _p(pos), _p(nrm), C.c_uint32(len(pos)), _p(idx),
                                          C.c_uint32(len(idx)), C.c_uint32(start_index), C.c_float(inverse_voxel_extent),
                                          _p(out), C.c_uint32(len(out)))
    return out[:n].copy()


def mutual_contacts(obj_a: "Object", probes_a: "CollisionProbes", moments_a, world_to_a, obj_b: "Object",
                    probes_b: "CollisionProbes", moments_b, world_to_b, ranges_in_a, ranges_in_b):
    """`for_each_mutual_voxel_object_contact` (collidable.rs:859-1050) given the voxel ranges encompassing the intersection:
    isometries as 7 floats (unit quaternion x, y, z, w; translation), moments as the managers' 10 floats →
    (contacts of A's probes in B, contacts of B's probes in A), each in ascending chunk / point order."""
    ma, mb = np.ascontiguousarray(moments_a, np.float32), np.ascontiguousarray(moments_b, np.float32)
    wa, wb = np.ascontiguousarray(world_to_a, np.float32), np.ascontiguousarray(world_to_b, np.float32)
    ra, rb = np.ascontiguousarray(ranges_in_a, np.uint32).reshape(6), np.ascontiguousarray(ranges_in_b, np.uint32).reshape(6)
    assert ma.shape == (10,) and mb.shape == (10,) and wa.shape == (7,) and wb.shape == (7,)
    counts = np.zeros(2, np.uint32)
    lib().orc_mutual_contacts.restype = C.c_uint32
    args = (obj_a.h, probes_a.h, _p(ma), _p(wa), obj_b.h, probes_b.h, _p(mb), _p(wb), _p(ra), _p(rb))
    n = lib().orc_mutual_contacts(*args, None, C.c_uint32(0), _p(counts))
    out = np.zeros(max(1, n), Object.CONTACT_DTYPE)
    lib().orc_mutual_contacts(*args, _p(out), C.c_uint32(len(out)), _p(counts))
    return out[: counts[0]].copy(), out[counts[0]: n].copy()
