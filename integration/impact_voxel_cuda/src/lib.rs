//! `impact_voxel_cuda`: the FFI table of `libimpact_voxel_cuda.so` (include/impact_voxel_cuda.h) for the Impact engine,
//! loaded with the engine's own `dynamic_lib::define_lib!` (interop/dynamic_lib/src/macros.rs:17-29). Drop this crate
//! under `engine/crates/impact_voxel_cuda/`; INTEGRATION.md lists the call sites of `impact_voxel` to swap.
//! NOT compiled in the build container of this repository (no Rust toolchain there): it is the source a maintainer adds.
//! `Voxel`, `VoxelMeshIndexMaterials`, `ChunkSubmesh` come from `impact_voxel` (already `#[repr(C)]` + `Pod`);
//! `IvxSplitInfo`, `IvxChunkRegions`, `IvxExtractionInfo`, `IvxObjectInfo`, `IvxMeshInfo` mirror the structs of the same
//! names (`ivx_split_info`, ...) in the header field for field.
#![allow(clippy::missing_safety_doc)]
use impact_voxel::{Voxel, mesh::{ChunkSubmesh, VoxelMeshIndexMaterials}};
use std::ffi::{c_char, c_void};

#[repr(C)] pub struct IvxConfig { pub abi_version: u32, pub device: i32, pub stream: *mut c_void, pub flags: u32 }
#[repr(C)] pub struct IvxSdfNode { pub kind: u32, pub child: [u32; 2], pub octaves: u32, pub seed: u32, pub p: [f32; 8] }
#[repr(C)] pub struct IvxTypeGenerator { pub kind: u32, pub same_type: u32, pub n_types: u32,
                                         pub noise_frequency: f32, pub voxel_type_frequency: f32, pub seed: u32 }
#[repr(C)] pub struct IvxChunkDesc { pub kind: u8, pub flags: u8, pub face: [u8; 6], pub uniform_voxel: [u8; 3],
                                     pub _pad: u8, pub data_offset: u32 }
#[repr(C)] pub struct IvxObjectInfo { /* include/impact_voxel_cuda.h: ivx_object_info */ }
#[repr(C)] pub struct IvxNode        { /* include/impact_voxel_cuda.h: ivx_node (ProcessedSDFNode, 144 bytes) */ }
#[repr(C)] pub struct IvxProgramInfo { /* include/impact_voxel_cuda.h: ivx_program_info */ }
#[repr(C)] pub struct IvxMeshInfo   { /* include/impact_voxel_cuda.h: ivx_mesh_info   */ }
#[repr(C)] pub struct IvxAbsorbStats { pub touched_chunks: u32, pub touched_voxels: u32, pub emptied_voxels: u32,
                                       pub removed_chunks: u32, pub dirty_chunks: u32 }
pub enum IvxCtx {} pub enum IvxProgram {} pub enum IvxObject {} pub enum IvxComm {} pub enum IvxMeshGpuBuffers {}
#[repr(C)] pub struct IvxCommConfig { pub rank: u32, pub world: u32, pub gather_rank: u32, pub plane_chunks: u32,
                                      pub mesh_vertices: u64, pub mesh_indices: u64, pub mesh_submeshes: u64 }
#[repr(C)] pub struct IvxGatheredMesh { pub n_vertices: u64, pub n_indices: u64, pub n_submeshes: u64,
                                        pub d_positions: *mut c_void, pub d_normals: *mut c_void, pub d_indices: *mut c_void,
                                        pub d_index_materials: *mut c_void, pub d_submeshes: *mut c_void,
                                        pub d_vertex_ranges: *mut c_void }

/// `ParamSource` after name resolution (meta/params.rs:40-62): kind 0 Fixed(value), 1 FromParam { idx, Linear { offset: value, scale } }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct IvxMetaSource { pub kind: u32, pub idx: u32, pub value: f32, pub scale: f32 }
/// `ContParamSpec` / `DiscreteParamSpec`: dist 0 Constant, 1 Uniform, 2 UniformCosAngle, 3 PowerLaw
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct IvxMetaParam { pub dist: u32, pub src: [IvxMetaSource; 3] }
/// one `MetaSDFNode` (meta.rs:55-80); `kind` = the variant's position in the enum, `params` in field order
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct IvxMetaNode { pub kind: u32, pub child: [u32; 2], pub count: u32, pub seed: u32, pub sampling: u32,
                         pub composition: u32, pub rotation: u32, pub anchor: u32, pub min_pick_count: u32,
                         pub max_pick_count: u32, pub pick_probability: f32, pub smoothness: f32,
                         pub params: [IvxMetaParam; 8] }

/// `ivx_probes_info` / `ivx_probe_range`: VoxelObjectCollisionProbes as the library keeps them (collidable.rs:97-101)
#[repr(C)] pub struct IvxProbesInfo { pub log2_block_size: u32, pub _pad: u32, pub n_points: u64, pub n_chunks: u64,
                                      pub d_points: *const f32 }
#[repr(C)] pub struct IvxProbeRange { pub chunk_indices: [u32; 3], pub point_start: u32, pub point_end: u32 }

/// `ivx_mesh_gpu_buffers_info`: VoxelMeshGPUBuffers (gpu_resource.rs:460-900) as exportable device allocations — one
/// descriptor per buffer (positions, normals, index materials, indices, chunk submeshes) for
/// `wgpu::hal` / Vulkan external-memory import instead of the staging-belt upload.
#[repr(C)] #[derive(Clone, Copy)] pub struct IvxMeshGpuBufferInfo { pub fd: i32, pub recreated: u32, pub allocation_bytes: u64,
                                                                   pub valid_bytes: u64, pub device_ptr: *mut c_void }
#[repr(C)] pub struct IvxMeshGpuBuffersInfo { pub buffer: [IvxMeshGpuBufferInfo; 5], pub n_vertices: u64, pub n_indices: u64,
                                              pub n_chunks: u64, pub bytes_copied: u64, pub n_updated_ranges: u32, pub reserved: u32 }

dynamic_lib::define_lib! {
    name = VoxelCudaLib,
    path_env_var = "IMPACT_VOXEL_CUDA_LIB",
    fallback_path = "./libimpact_voxel_cuda";

    unsafe fn ivx_create(config: *const IvxConfig, out_ctx: *mut *mut IvxCtx) -> i32;
    unsafe fn ivx_destroy(ctx: *mut IvxCtx) -> ();
    unsafe fn ivx_last_error(ctx: *const IvxCtx) -> *const c_char;
    unsafe fn ivx_program_build(ctx: *mut IvxCtx, nodes: *const IvxSdfNode, n_nodes: u32, root: u32,
                                out: *mut *mut IvxProgram) -> i32;
    unsafe fn ivx_program_free(ctx: *mut IvxCtx, program: *mut IvxProgram) -> ();
    unsafe fn ivx_meta_compile(ctx: *mut IvxCtx, nodes: *const IvxMetaNode, n_nodes: u32, scale_factor: f32, seed: u64,
                               out_nodes: *mut IvxSdfNode, capacity: u32, out_count: *mut u32, out_root: *mut u32,
                               out_empty: *mut i32, err: *mut c_char, err_capacity: usize) -> i32;
    unsafe fn ivx_object_generate(ctx: *mut IvxCtx, program: *const IvxProgram, voxel_extent: f32,
                                  types: *const IvxTypeGenerator, out: *mut *mut IvxObject) -> i32;
    unsafe fn ivx_object_info_get(ctx: *mut IvxCtx, object: *const IvxObject, out: *mut IvxObjectInfo) -> i32;
    unsafe fn ivx_object_download(ctx: *mut IvxCtx, object: *const IvxObject, chunks: *mut IvxChunkDesc,
                                  chunk_capacity: usize, voxels: *mut Voxel, voxel_capacity: usize) -> i32;
    unsafe fn ivx_object_mesh(ctx: *mut IvxCtx, object: *mut IvxObject, out: *mut IvxMeshInfo) -> i32;
    unsafe fn ivx_mesh_download(ctx: *mut IvxCtx, object: *const IvxObject, positions: *mut f32, normals: *mut f32,
                                index_materials: *mut VoxelMeshIndexMaterials, indices: *mut u32,
                                submeshes: *mut ChunkSubmesh, vertex_ranges: *mut u32) -> i32;
    unsafe fn ivx_mesh_download_checked(ctx: *mut IvxCtx, object: *const IvxObject, n_vertices: u32, n_indices: u32,
                                        n_submeshes: u32, positions: *mut f32, normals: *mut f32,
                                        index_materials: *mut VoxelMeshIndexMaterials, indices: *mut u32,
                                        submeshes: *mut ChunkSubmesh, vertex_ranges: *mut u32) -> i32;
    unsafe fn ivx_object_absorb_sphere(ctx: *mut IvxCtx, object: *mut IvxObject, center: *const f32, radius: f32,
                                       influence_radius: f32, stats: *mut IvxAbsorbStats) -> i32;
    unsafe fn ivx_object_absorb_capsule(ctx: *mut IvxCtx, object: *mut IvxObject, segment_start: *const f32,
                                        segment_vector: *const f32, radius: f32, influence_radius: f32,
                                        stats: *mut IvxAbsorbStats) -> i32;
    unsafe fn ivx_object_generate_streamed(ctx: *mut IvxCtx, program: *const IvxProgram, voxel_extent: f32,
                                           types: *const IvxTypeGenerator, chunks: *mut IvxChunkDesc,
                                           chunk_capacity: usize, voxels: *mut Voxel, voxel_capacity: usize,
                                           out: *mut *mut IvxObject, out_non_uniform_chunks: *mut u64) -> i32;
    unsafe fn ivx_synchronize(ctx: *mut IvxCtx) -> i32;
    unsafe fn ivx_object_remesh_dirty(ctx: *mut IvxCtx, object: *mut IvxObject, out: *mut IvxMeshInfo) -> i32;
    unsafe fn ivx_object_resolve_connected_regions(ctx: *mut IvxCtx, object: *mut IvxObject, out: *mut IvxSplitInfo) -> i32;
    unsafe fn ivx_object_split_detection_download(ctx: *mut IvxCtx, object: *const IvxObject, voxel_labels: *mut u8,
                                                  label_capacity: usize, per_chunk: *mut IvxChunkRegions,
                                                  chunk_capacity: usize, region_roots: *mut u32,
                                                  region_capacity: usize) -> i32;
    unsafe fn ivx_object_extract_disconnected_region(ctx: *mut IvxCtx, object: *mut IvxObject, info: *mut IvxExtractionInfo,
                                                     out_extracted: *mut *mut IvxObject) -> i32;
    unsafe fn ivx_object_from_generated_chunks(ctx: *mut IvxCtx, voxel_extent: f32, grid_shape: *const u32,
                                               voxels: *const Voxel, sparseness: *const u8,
                                               out: *mut *mut IvxObject) -> i32;
    unsafe fn ivx_object_inertial_moments(ctx: *mut IvxCtx, object: *const IvxObject, voxel_type_densities: *const f32,
                                          n_densities: u32, initial: *const IvxInertialMoments,
                                          out: *mut IvxInertialMoments, per_chunk_terms: *mut f32,
                                          per_chunk_capacity: usize) -> i32;
    unsafe fn ivx_object_absorb_sphere_inertial(ctx: *mut IvxCtx, object: *mut IvxObject, center: *const f32, radius: f32,
                                                influence_radius: f32, voxel_type_densities: *const f32, n_densities: u32,
                                                inout_moments: *mut IvxInertialMoments, stats: *mut IvxAbsorbStats) -> i32;
    unsafe fn ivx_object_absorb_capsule_inertial(ctx: *mut IvxCtx, object: *mut IvxObject, segment_start: *const f32,
                                                 segment_vector: *const f32, radius: f32, influence_radius: f32,
                                                 voxel_type_densities: *const f32, n_densities: u32,
                                                 inout_moments: *mut IvxInertialMoments, stats: *mut IvxAbsorbStats) -> i32;
    unsafe fn ivx_objects_absorb_mutually(ctx: *mut IvxCtx, object_a: *mut IvxObject, object_b: *mut IvxObject,
                                          transform_from_b_to_a: *const IvxIsometry, smoothness: f32,
                                          ranges_in_a: *const u32, ranges_in_b: *const u32,
                                          voxel_type_densities: *const f32, n_densities: u32,
                                          inout_a: *mut IvxInertialMoments, inout_b: *mut IvxInertialMoments,
                                          stats_a: *mut IvxAbsorbStats, stats_b: *mut IvxAbsorbStats) -> i32;
    unsafe fn ivx_object_surface_voxels_in_ranges(ctx: *mut IvxCtx, object: *const IvxObject, ranges: *const u32,
                                                  out: *mut IvxSurfaceVoxel, capacity: usize, out_count: *mut u64) -> i32;
    unsafe fn ivx_object_surface_voxels_touching_sphere(ctx: *mut IvxCtx, object: *const IvxObject, center: *const f32,
                                                        radius: f32, out: *mut IvxSurfaceVoxel, capacity: usize,
                                                        out_count: *mut u64) -> i32;
    unsafe fn ivx_object_surface_voxels_within_plane(ctx: *mut IvxCtx, object: *const IvxObject, unit_normal: *const f32,
                                                     displacement: f32, out: *mut IvxSurfaceVoxel, capacity: usize,
                                                     out_count: *mut u64) -> i32;
    unsafe fn ivx_object_sphere_contacts(ctx: *mut IvxCtx, object: *const IvxObject, transform_to_object_space: *const IvxIsometry,
                                         center: *const f32, radius: f32, out: *mut IvxVoxelContact, capacity: usize,
                                         out_count: *mut u64) -> i32;
    unsafe fn ivx_intersection_voxel_ranges(occupied_a: *const u32, voxel_extent_a: f32, occupied_b: *const u32,
                                            voxel_extent_b: f32, transform_from_b_to_a: *const IvxIsometry,
                                            out_ranges_in_a: *mut u32, out_ranges_in_b: *mut u32,
                                            out_intersect: *mut i32) -> i32;

    // ---- multi-GPU communicator over peer memory (include/impact_voxel_cuda.h "multi-GPU communicator") ----
    unsafe fn ivx_object_generate_slab(ctx: *mut IvxCtx, program: *const IvxProgram, voxel_extent: f32,
                                       types: *const IvxTypeGenerator, chunk_i_begin: u32, chunk_i_end: u32,
                                       out: *mut *mut IvxObject) -> i32;
    unsafe fn ivx_program_plane_work(ctx: *mut IvxCtx, program: *const IvxProgram, voxel_extent: f32,
                                     types: *const IvxTypeGenerator, out_work: *mut u32, capacity: u32,
                                     out_planes: *mut u32) -> i32;
    unsafe fn ivx_comm_create(ctx: *mut IvxCtx, config: *const IvxCommConfig, out_comm: *mut *mut IvxComm,
                              out_handle: *mut u8) -> i32;
    unsafe fn ivx_comm_connect(ctx: *mut IvxCtx, comm: *mut IvxComm, all_handles: *const u8) -> i32;
    unsafe fn ivx_comm_destroy(ctx: *mut IvxCtx, comm: *mut IvxComm) -> ();
    unsafe fn ivx_object_exchange_halos(ctx: *mut IvxCtx, comm: *mut IvxComm, object: *mut IvxObject, lower_rank: i32,
                                        upper_rank: i32) -> i32;
    unsafe fn ivx_object_mesh_gather(ctx: *mut IvxCtx, comm: *mut IvxComm, object: *mut IvxObject,
                                     out_local: *mut IvxMeshInfo, out_merged: *mut IvxGatheredMesh) -> i32;
    unsafe fn ivx_object_download_async(ctx: *mut IvxCtx, object: *mut IvxObject, chunks: *mut IvxChunkDesc,
                                        chunk_capacity: usize, voxels: *mut Voxel, voxel_capacity: usize,
                                        out_non_uniform_chunks: *mut u64) -> i32;
    unsafe fn ivx_object_collision_probes(ctx: *mut IvxCtx, object: *mut IvxObject, out: *mut IvxProbesInfo) -> i32;
    unsafe fn ivx_object_collision_probes_sync(ctx: *mut IvxCtx, object: *mut IvxObject, out: *mut IvxProbesInfo) -> i32;
    unsafe fn ivx_collision_probes_download(ctx: *mut IvxCtx, object: *const IvxObject, points: *mut f32, capacity_points: usize,
                                            ranges: *mut IvxProbeRange, capacity_ranges: usize) -> i32;
    unsafe fn ivx_objects_mutual_contacts(ctx: *mut IvxCtx, object_a: *const IvxObject, object_b: *const IvxObject,
                                          world_to_a: *const IvxIsometry, world_to_b: *const IvxIsometry,
                                          ranges_in_a: *const u32, ranges_in_b: *const u32,
                                          inertial_a: *const IvxInertialMoments, inertial_b: *const IvxInertialMoments,
                                          out: *mut IvxVoxelContact, capacity: usize, out_count_a_against_b: *mut u64,
                                          out_count_b_against_a: *mut u64) -> i32;
    unsafe fn ivx_mesh_gpu_buffers_create(ctx: *mut IvxCtx, object: *mut IvxObject, out: *mut *mut IvxMeshGpuBuffers, info: *mut IvxMeshGpuBuffersInfo) -> i32;
    unsafe fn ivx_mesh_gpu_buffers_sync(ctx: *mut IvxCtx, object: *mut IvxObject, buffers: *mut IvxMeshGpuBuffers, info: *mut IvxMeshGpuBuffersInfo) -> i32;
    unsafe fn ivx_mesh_gpu_buffers_destroy(ctx: *mut IvxCtx, buffers: *mut IvxMeshGpuBuffers) -> ();
    // ---- the rest of include/impact_voxel_cuda.h (generated by tools/gen_rust_bindings.py from the prototypes) ----
    unsafe fn ivx_abi_version() -> u32;
    unsafe fn ivx_kernel_launch_count(ctx: *const IvxCtx) -> u64;
    unsafe fn ivx_profile_enable(ctx: *mut IvxCtx, enabled: i32) -> i32;
    unsafe fn ivx_profile_reset(ctx: *mut IvxCtx) -> i32;
    unsafe fn ivx_profile_get(ctx: *mut IvxCtx, kernel_id: u32, out_total_ms: *mut f64, out_launches: *mut u64) -> i32;
    unsafe fn ivx_profile_counter(ctx: *mut IvxCtx, counter_id: u32, out_value: *mut u64) -> i32;
    unsafe fn ivx_program_upload(ctx: *mut IvxCtx, nodes: *const IvxNode, n_nodes: u32, stack_depth: u32,
                                 domain_lo: *const f32, domain_hi: *const f32, out_program: *mut *mut IvxProgram) -> i32;
    unsafe fn ivx_program_compile_host(nodes: *const IvxSdfNode, n_nodes: u32, root_node_id: u32, out_nodes: *mut IvxNode,
                                       capacity: u32, out_count: *mut u32, out_info: *mut IvxProgramInfo, err: *mut c_char,
                                       err_capacity: usize) -> i32;
    unsafe fn ivx_program_info_get(ctx: *mut IvxCtx, program: *const IvxProgram, out: *mut IvxProgramInfo) -> i32;
    unsafe fn ivx_program_nodes(ctx: *mut IvxCtx, program: *const IvxProgram, out: *mut IvxNode, capacity: u32) -> i32;
    unsafe fn ivx_program_eval_chunks(ctx: *mut IvxCtx, program: *const IvxProgram, chunk_origins: *const f32,
                                      n_chunks: u32, out_signed_distances: *mut f32) -> i32;
    unsafe fn ivx_program_eval_blocks(ctx: *mut IvxCtx, program: *const IvxProgram, block_origins: *const f32,
                                      n_blocks: u32, size: u32, out_signed_distances: *mut f32) -> i32;
    unsafe fn ivx_object_halo_capacity(ctx: *mut IvxCtx, object: *const IvxObject, out_bytes: *mut usize) -> i32;
    unsafe fn ivx_object_halo_export(ctx: *mut IvxCtx, object: *const IvxObject, side: i32, device_buffer: *mut c_void,
                                     capacity: usize, out_bytes: *mut usize) -> i32;
    unsafe fn ivx_object_halo_import(ctx: *mut IvxCtx, object: *mut IvxObject, side: i32, device_buffer: *const c_void,
                                     bytes: usize) -> i32;
    unsafe fn ivx_object_slab_classify(ctx: *mut IvxCtx, object: *mut IvxObject) -> i32;
    unsafe fn ivx_object_halo_kinds_export(ctx: *mut IvxCtx, object: *const IvxObject, side: i32,
                                           device_buffer: *mut c_void, capacity: usize) -> i32;
    unsafe fn ivx_object_halo_kinds_import(ctx: *mut IvxCtx, object: *mut IvxObject, side: i32,
                                           device_buffer: *const c_void, bytes: usize) -> i32;
    unsafe fn ivx_object_slab_finalize(ctx: *mut IvxCtx, object: *mut IvxObject) -> i32;
    unsafe fn ivx_peer_alloc(ctx: *mut IvxCtx, bytes: usize, out_device_ptr: *mut *mut c_void, out_handle: *mut u8) -> i32;
    unsafe fn ivx_peer_free(ctx: *mut IvxCtx, device_ptr: *mut c_void) -> i32;
    unsafe fn ivx_peer_open(ctx: *mut IvxCtx, handle: *const u8, out_device_ptr: *mut *mut c_void) -> i32;
    unsafe fn ivx_peer_close(ctx: *mut IvxCtx, device_ptr: *mut c_void) -> i32;
    unsafe fn ivx_mesh_push(ctx: *mut IvxCtx, object: *const IvxObject, merged_base: *mut c_void,
                            field_offsets: *const u64, vertex_base: u32, index_base: u32, submesh_base: u32) -> i32;
    unsafe fn ivx_object_mesh_sync(ctx: *mut IvxCtx, object: *mut IvxObject, out: *mut IvxMeshInfo) -> i32;
    unsafe fn ivx_mesh_modifications(ctx: *mut IvxCtx, object: *const IvxObject, out_ranges: *mut u32,
                                     capacity_records: usize, out_count: *mut u64,
                                     out_chunks_were_removed: *mut i32) -> i32;
    unsafe fn ivx_mesh_report_synchronized(ctx: *mut IvxCtx, object: *mut IvxObject) -> i32;
    unsafe fn ivx_comm_connect_local(ctx: *mut IvxCtx, comm: *mut IvxComm, all_comms: *mut *mut IvxComm) -> i32;
    unsafe fn ivx_object_mesh_distributed(ctx: *mut IvxCtx, comm: *mut IvxComm, object: *mut IvxObject,
                                          out_local: *mut IvxMeshInfo, out_bases: *mut u64) -> i32;
    unsafe fn ivx_box_intersection_bounds(a_lower: *const f32, a_upper: *const f32, b_center: *const f32,
                                          b_orientation: *const f32, b_half_extents: *const f32, out_in_a: *mut f32,
                                          out_in_b: *mut f32, out_intersect: *mut i32) -> i32;
    unsafe fn ivx_object_dirty_chunks(ctx: *mut IvxCtx, object: *const IvxObject, out_linear_indices: *mut u32,
                                      capacity: u32, out_count: *mut u32) -> i32;
    unsafe fn ivx_object_plane_contacts(ctx: *mut IvxCtx, object: *const IvxObject,
                                        transform_to_object_space: *const IvxIsometry, unit_normal: *const f32,
                                        displacement: f32, out: *mut IvxVoxelContact, capacity: usize,
                                        out_count: *mut u64) -> i32;
    unsafe fn ivx_object_capsule_contacts(ctx: *mut IvxCtx, object: *const IvxObject,
                                          transform_to_object_space: *const IvxIsometry, segment_start: *const f32,
                                          segment_vector: *const f32, radius: f32, out: *mut IvxVoxelContact,
                                          capacity: usize, out_count: *mut u64) -> i32;
    unsafe fn ivx_object_surface_voxels_touching_capsule(ctx: *mut IvxCtx, object: *const IvxObject,
                                                         segment_start: *const f32, segment_vector: *const f32,
                                                         radius: f32, out: *mut IvxSurfaceVoxel, capacity: usize,
                                                         out_count: *mut u64) -> i32;
    unsafe fn ivx_voxel_ranges_within_plane(occupied: *const u32, unit_normal: *const f32, displacement: f32,
                                            out_ranges: *mut u32) -> i32;
    unsafe fn ivx_object_free(ctx: *mut IvxCtx, object: *mut IvxObject) -> ();
}
/// one call of the closures of `for_each_surface_voxel_*`: indices, the voxel, `VoxelSurfacePlacement` as u8
#[repr(C)] pub struct IvxSurfaceVoxel { pub indices: [u32; 3], pub voxel: Voxel, pub placement: u8 }
/// `([usize; 3], ContactGeometry)` as `for_each_sphere_voxel_object_contact` hands them to its closure
#[repr(C)] pub struct IvxVoxelContact { pub indices: [u32; 3], pub position: [f32; 3], pub surface_normal: [f32; 3],
                                        pub penetration_depth: f32 }
/// `Isometry3` as unit quaternion (x, y, z, w) + translation (impact_math/src/transform/isometry.rs)
#[repr(C)] pub struct IvxIsometry { pub rotation: [f32; 4], pub translation: [f32; 3] }

/// `VoxelObjectInertialPropertyManager`'s four fields in declaration order (V/object/inertia.rs:19-25); the manager
/// itself only needs `#[repr(C)]` (Vector3 = three f32) to be passed as it is.
#[repr(C)] pub struct IvxInertialMoments { pub mass: f32, pub moments: [f32; 3], pub moments_of_inertia: [f32; 3],
                                           pub products_of_inertia: [f32; 3] }
