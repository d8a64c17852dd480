"""ctypes binding of impact_b200/libimpact_voxel_cuda.so (include/impact_voxel_cuda.h).

The library is the product; there is no Python or CPU fallback. If the shared
object is missing this module raises at import of `lib()`, and `ivx_create`
fails with IVX_ERR_NO_DEVICE on a machine without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("IMPACT_VOXEL_CUDA_LIB", os.path.join(_HERE, "libimpact_voxel_cuda.so"))

IVX_OK = 0
STATUS_NAMES = {
    0: "IVX_OK", 1: "IVX_ERR_INVALID_ARGUMENT", 2: "IVX_ERR_GRAPH", 3: "IVX_ERR_CUDA", 4: "IVX_ERR_OUT_OF_MEMORY",
    5: "IVX_ERR_CAPACITY", 6: "IVX_ERR_UNSUPPORTED", 7: "IVX_ERR_NO_DEVICE",
}

# every symbol include/impact_voxel_cuda.h declares
EXPORTED_SYMBOLS = [
    "ivx_create", "ivx_destroy", "ivx_last_error", "ivx_abi_version", "ivx_kernel_launch_count", "ivx_synchronize", "ivx_profile_enable", "ivx_profile_reset", "ivx_profile_get", "ivx_profile_counter",
    "ivx_program_build", "ivx_program_upload", "ivx_program_compile_host", "ivx_program_info_get", "ivx_program_nodes", "ivx_program_free",
    "ivx_meta_compile", "ivx_program_eval_chunks", "ivx_program_eval_blocks", "ivx_object_generate", "ivx_object_generate_streamed", "ivx_object_generate_slab", "ivx_program_plane_work", "ivx_object_halo_capacity", "ivx_object_halo_export",
    "ivx_object_halo_import", "ivx_object_slab_classify", "ivx_object_halo_kinds_export", "ivx_object_halo_kinds_import",
    "ivx_object_slab_finalize", "ivx_object_info_get",
    "ivx_object_download", "ivx_object_download_async", "ivx_object_free", "ivx_object_mesh", "ivx_mesh_download", "ivx_mesh_download_checked", "ivx_peer_alloc", "ivx_peer_free", "ivx_peer_open", "ivx_peer_close", "ivx_mesh_push", "ivx_object_absorb_sphere", "ivx_object_absorb_capsule",
    "ivx_object_dirty_chunks", "ivx_object_remesh_dirty",
    "ivx_object_resolve_connected_regions", "ivx_object_split_detection_download", "ivx_object_extract_disconnected_region",
    "ivx_object_from_generated_chunks", "ivx_object_inertial_moments", "ivx_object_absorb_sphere_inertial", "ivx_object_absorb_capsule_inertial",
    "ivx_objects_absorb_mutually", "ivx_intersection_voxel_ranges", "ivx_box_intersection_bounds",
    "ivx_object_surface_voxels_in_ranges", "ivx_object_surface_voxels_touching_sphere", "ivx_object_surface_voxels_touching_capsule",
    "ivx_object_surface_voxels_within_plane", "ivx_voxel_ranges_within_plane", "ivx_object_sphere_contacts", "ivx_object_plane_contacts", "ivx_object_capsule_contacts",
    "ivx_comm_create", "ivx_comm_connect", "ivx_comm_connect_local", "ivx_comm_destroy", "ivx_object_exchange_halos",
    "ivx_object_mesh_gather", "ivx_object_mesh_distributed", "ivx_object_mesh_sync", "ivx_mesh_modifications", "ivx_mesh_report_synchronized", "ivx_object_collision_probes", "ivx_object_collision_probes_sync", "ivx_collision_probes_download", "ivx_objects_mutual_contacts",
    "ivx_mesh_gpu_buffers_create", "ivx_mesh_gpu_buffers_sync", "ivx_mesh_gpu_buffers_destroy",
]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("stream", C.c_void_p), ("flags", C.c_uint32)]


class ProgramInfo(C.Structure):
    _fields_ = [("node_count", C.c_uint32), ("stack_depth", C.c_uint32), ("domain_lo", C.c_float * 3),
                ("domain_hi", C.c_float * 3)]


class ObjectInfo(C.Structure):
    _fields_ = [("voxel_extent", C.c_float), ("grid_shape", C.c_uint32 * 3), ("chunk_counts", C.c_uint32 * 3),
                ("chunk_i_begin", C.c_uint32), ("chunk_i_end", C.c_uint32), ("n_void", C.c_uint32),
                ("n_uniform", C.c_uint32), ("n_non_uniform", C.c_uint32), ("occupied_chunk_ranges", C.c_uint32 * 6),
                ("occupied_voxel_ranges", C.c_uint32 * 6), ("device_bytes", C.c_uint64)]


class MeshInfo(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("n_indices", C.c_uint32), ("n_submeshes", C.c_uint32),
                ("n_exposed_chunks", C.c_uint32), ("d_positions", C.c_void_p), ("d_normals", C.c_void_p),
                ("d_index_materials", C.c_void_p), ("d_indices", C.c_void_p), ("d_submeshes", C.c_void_p),
                ("d_vertex_ranges", C.c_void_p)]


class CommConfig(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("gather_rank", C.c_uint32), ("plane_chunks", C.c_uint32),
                ("mesh_vertices", C.c_uint64), ("mesh_indices", C.c_uint64), ("mesh_submeshes", C.c_uint64)]


class ProbesInfo(C.Structure):
    _fields_ = [("log2_block_size", C.c_uint32), ("_pad", C.c_uint32), ("n_points", C.c_uint64), ("n_chunks", C.c_uint64),
                ("d_points", C.c_void_p)]


class MeshGpuBufferInfo(C.Structure):
    _fields_ = [("fd", C.c_int32), ("recreated", C.c_uint32), ("allocation_bytes", C.c_uint64), ("valid_bytes", C.c_uint64),
                ("device_ptr", C.c_void_p)]


MESH_BUFFER_NAMES = ("positions", "normals", "index_materials", "indices", "chunk_submeshes")


class MeshGpuBuffersInfo(C.Structure):
    _fields_ = [("buffer", MeshGpuBufferInfo * 5), ("n_vertices", C.c_uint64), ("n_indices", C.c_uint64),
                ("n_chunks", C.c_uint64), ("bytes_copied", C.c_uint64), ("n_updated_ranges", C.c_uint32), ("reserved", C.c_uint32)]


PROBE_RANGE_DTYPE = np.dtype([("chunk_indices", "<u4", (3,)), ("point_start", "<u4"), ("point_end", "<u4")])


class MetaSource(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("idx", C.c_uint32), ("value", C.c_float), ("scale", C.c_float)]


class MetaParam(C.Structure):
    _fields_ = [("dist", C.c_uint32), ("src", MetaSource * 3)]


META_MAX_PARAMS = 8


class MetaNode(C.Structure):
    """`ivx_meta_node` (include/impact_voxel_cuda.h)."""
    _fields_ = [("kind", C.c_uint32), ("child", C.c_uint32 * 2), ("count", C.c_uint32), ("seed", C.c_uint32),
                ("sampling", C.c_uint32), ("composition", C.c_uint32), ("rotation", C.c_uint32), ("anchor", C.c_uint32),
                ("min_pick_count", C.c_uint32), ("max_pick_count", C.c_uint32), ("pick_probability", C.c_float),
                ("smoothness", C.c_float), ("params", MetaParam * META_MAX_PARAMS)]


class GatheredMesh(C.Structure):
    _fields_ = [("n_vertices", C.c_uint64), ("n_indices", C.c_uint64), ("n_submeshes", C.c_uint64),
                ("d_positions", C.c_void_p), ("d_normals", C.c_void_p), ("d_indices", C.c_void_p),
                ("d_index_materials", C.c_void_p), ("d_submeshes", C.c_void_p), ("d_vertex_ranges", C.c_void_p)]


class ExtractionInfo(C.Structure):
    _fields_ = [("n_regions_before", C.c_uint32), ("found_two", C.c_uint32), ("extracted", C.c_uint32),
                ("discarded", C.c_uint32), ("single_chunk", C.c_uint32), ("region_label", C.c_uint32),
                ("region_chunks", C.c_uint32), ("moved_non_empty_voxels", C.c_uint32),
                ("origin_offset_in_parent", C.c_uint32 * 3)]


class AbsorbStats(C.Structure):
    _fields_ = [("touched_chunks", C.c_uint32), ("touched_voxels", C.c_uint32), ("emptied_voxels", C.c_uint32),
                ("removed_chunks", C.c_uint32), ("dirty_chunks", C.c_uint32)]


class RegionCandidate(C.Structure):
    _fields_ = [("label", C.c_uint32), ("chunk_count", C.c_uint32), ("non_uniform_chunk_count", C.c_uint32),
                ("chunk_min", C.c_uint32 * 3), ("chunk_max", C.c_uint32 * 3)]


class SplitInfo(C.Structure):
    _fields_ = [("n_regions", C.c_uint32), ("has_two", C.c_uint32), ("candidates", RegionCandidate * 2),
                ("smallest", C.c_uint32), ("n_local_regions", C.c_uint32), ("n_connections", C.c_uint32),
                ("n_relabelled_chunks", C.c_uint32), ("device_ms", C.c_float), ("host_ms", C.c_float)]


SURFACE_VOXEL_DTYPE = np.dtype([("indices", "<u4", (3,)), ("type", "u1"), ("sd", "i1"), ("flags", "u1"), ("placement", "u1")])
assert SURFACE_VOXEL_DTYPE.itemsize == 16
CONTACT_DTYPE = np.dtype([("indices", "<u4", (3,)), ("position", "<f4", (3,)), ("normal", "<f4", (3,)), ("depth", "<f4")])
assert CONTACT_DTYPE.itemsize == 40
CHUNK_REGIONS_DTYPE = np.dtype([("region_count", "<u2"), ("boundary_region_count", "<u2"), ("first_region", "<u4")])
VOXEL_DTYPE = np.dtype([("type", "u1"), ("sd", "i1"), ("flags", "u1")])
CHUNK_DTYPE = np.dtype(
    [("kind", "u1"), ("flags", "u1"), ("face", "u1", (6,)), ("uniform_type", "u1"), ("uniform_sd", "i1"),
     ("uniform_flags", "u1"), ("_pad", "u1"), ("data_offset", "<u4")]
)
SUBMESH_DTYPE = np.dtype(
    [("chunk_indices", "<u4", (3,)), ("index_offset", "<u4"), ("index_count", "<u4"), ("obscured", "<u4", (8,))]
)
INDEX_MATERIALS_DTYPE = np.dtype([("indices", "u1", (4,)), ("weights", "u1", (4,))])
assert CHUNK_DTYPE.itemsize == 16 and SUBMESH_DTYPE.itemsize == 52 and VOXEL_DTYPE.itemsize == 3

_lib = None


def lib():
    """Loads the shared library (once). Raises OSError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise OSError(
                f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the voxel hot path)"
            )
        L = C.CDLL(SO_PATH)
        L.ivx_last_error.restype = C.c_char_p
        L.ivx_last_error.argtypes = [C.c_void_p]
        L.ivx_abi_version.restype = C.c_uint32
        L.ivx_kernel_launch_count.restype = C.c_uint64
        L.ivx_kernel_launch_count.argtypes = [C.c_void_p]
        L.ivx_destroy.argtypes = [C.c_void_p]
        L.ivx_destroy.restype = None
        L.ivx_program_free.argtypes = [C.c_void_p, C.c_void_p]
        L.ivx_program_free.restype = None
        L.ivx_object_free.argtypes = [C.c_void_p, C.c_void_p]
        L.ivx_object_free.restype = None
        L.ivx_comm_destroy.argtypes = [C.c_void_p, C.c_void_p]
        L.ivx_comm_destroy.restype = None
        L.ivx_mesh_gpu_buffers_destroy.argtypes = [C.c_void_p, C.c_void_p]
        L.ivx_mesh_gpu_buffers_destroy.restype = None
        _lib = L
    return _lib


class IvxError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)
