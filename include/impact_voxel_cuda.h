/*
 * impact_voxel_cuda.h — C ABI of libimpact_voxel_cuda.so
 *
 * A B200 (sm_100a) implementation of the `impact_voxel` hot path of
 * lars-frogner/Impact: compile an atomic SDF graph, generate a chunked voxel
 * object from it, derive adjacency / obscuredness state, mesh it with Surface
 * Nets, and re-run the same stages on dirty chunks after sphere absorption.
 *
 * The reference has no FFI seam around this path (it is called through Rust
 * generics); each entry point below names the reference function it replaces.
 * Reference paths are relative to engine/crates/impact_voxel/src/.
 *
 * Conventions
 *  - every function returns an ivx_status (0 = OK) and never throws / aborts;
 *    ivx_last_error() gives the message of the last failing call on that ctx;
 *  - the caller owns all host memory, the library owns all device memory;
 *  - calls are synchronous on return (work is enqueued on the ctx stream and
 *    the stream is synchronised) unless the name ends in `_async`;
 *  - one ivx_ctx is used from one thread at a time (the engine already holds
 *    the voxel manager write lock around these calls).
 *  - there is NO CPU fallback: without a CUDA device ivx_create fails.
 */
#ifndef IMPACT_VOXEL_CUDA_H
#define IMPACT_VOXEL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IVX_CHUNK_SIZE 16u        /* object.rs:199-210 CHUNK_SIZE */
#define IVX_CHUNK_VOXELS 4096u
#define IVX_ABI_VERSION 1u

typedef enum ivx_status {
    IVX_OK = 0,
    IVX_ERR_INVALID_ARGUMENT = 1,
    IVX_ERR_GRAPH = 2,        /* cycle / missing node / bad kind: anyhow errors of atomic.rs:263-268 */
    IVX_ERR_CUDA = 3,
    IVX_ERR_OUT_OF_MEMORY = 4,
    IVX_ERR_CAPACITY = 5,     /* caller buffer too small */
    IVX_ERR_UNSUPPORTED = 6,
    IVX_ERR_NO_DEVICE = 7
} ivx_status;

/* SDFNode variants in declaration order (generation/sdf/atomic.rs:63-81). */
typedef enum ivx_node_kind {
    IVX_SPHERE = 0,
    IVX_CAPSULE = 1,
    IVX_BOX = 2,
    IVX_TRANSLATION = 3,
    IVX_ROTATION = 4,
    IVX_SCALING = 5,
    IVX_MULTIFRACTAL_NOISE = 6,
    IVX_UNION = 7,
    IVX_SUBTRACTION = 8,
    IVX_INTERSECTION = 9
} ivx_node_kind;

/* One atomic graph node = `SDFNode` (atomic.rs:63-181) as the host would
 * construct it with SDFNode::new_* (atomic.rs:1060-1128):
 *   sphere       p[0]=radius
 *   capsule      p[0]=segment_length p[1]=radius
 *   box          p[0..3]=extents
 *   translation  child[0], p[0..3]=translation
 *   rotation     child[0], p[0..4]=unit quaternion (x,y,z,w)
 *   scaling      child[0], p[0]=scaling
 *   noise        child[0], octaves, seed, p[0]=frequency p[1]=lacunarity
 *                p[2]=persistence p[3]=amplitude
 *   union/subtraction/intersection  child[0], child[1], p[0]=smoothness      */
typedef struct ivx_sdf_node {
    uint32_t kind;
    uint32_t child[2];
    uint32_t octaves;
    uint32_t seed;
    float p[8];
} ivx_sdf_node;

/* One compiled program node = `ProcessedSDFNode` (atomic.rs:83-102) after
 * SDFGenerator::new_in + determine_transforms_and_margins (atomic.rs:228-596).
 *   sphere p[0]=radius | capsule p[0]=half_segment_length p[1]=radius |
 *   box p[0..3]=half_extents | scaling p[0] | noise p[0..4] as above, p[4]=noise_scale |
 *   combine p[0]=smoothness p[1]=0.25/smoothness                              */
typedef struct ivx_node {
    uint32_t kind;
    uint32_t octaves;
    uint32_t seed;
    uint32_t leaf_count;
    float p[8];
    float transform_to_node_space[16]; /* column-major Matrix4 */
    float domain_lo[3];                /* domain_with_margin */
    float domain_hi[3];
    float margin;                      /* domain_margin */
    uint32_t _pad;
} ivx_node;

/* `Voxel` (lib.rs:60-66), repr(C), 3 bytes. */
typedef struct ivx_voxel {
    uint8_t voxel_type;
    int8_t signed_distance; /* VoxelSignedDistance: trunc_sat(f32 * 50) (lib.rs:154-201) */
    uint8_t flags;          /* VoxelFlags bits (lib.rs:75-101) */
} ivx_voxel;

/* `VoxelTypeGenerator` (generation/voxel_type.rs:9-36). */
typedef struct ivx_type_generator {
    uint32_t kind;       /* 0 = Same, 1 = GradientNoise */
    uint32_t same_type;  /* Same: the voxel type */
    uint32_t n_types;    /* GradientNoise: voxel_types.len() */
    float noise_frequency;
    float voxel_type_frequency;
    uint32_t seed;
} ivx_type_generator;

/* `VoxelChunk` (object.rs:96-126) flattened. kind: 0 Void, 1 Uniform, 2 NonUniform.
 * face[dim*2+side]: FaceVoxelDistribution 0 Empty, 1 Full, 2 Mixed.
 * flags: VoxelChunkFlags (object.rs:158-182). */
typedef struct ivx_chunk_desc {
    uint8_t kind;
    uint8_t flags;
    uint8_t face[6];
    ivx_voxel uniform_voxel;
    uint8_t _pad;
    uint32_t data_offset; /* units of 4096 voxels into the downloaded voxel array */
} ivx_chunk_desc;

/* `ChunkSubmesh` (mesh.rs:92-103). */
typedef struct ivx_chunk_submesh {
    uint32_t chunk_indices[3];
    uint32_t index_offset;
    uint32_t index_count;
    uint32_t is_obscured_from_direction[8];
} ivx_chunk_submesh;

/* `VoxelMeshIndexMaterials` (mesh.rs:79-84). */
typedef struct ivx_index_materials {
    uint8_t indices[4];
    uint8_t weights[4];
} ivx_index_materials;

typedef struct ivx_config {
    uint32_t abi_version;  /* IVX_ABI_VERSION */
    int32_t device;        /* CUDA device ordinal */
    void* stream;          /* cudaStream_t to enqueue on; NULL = library-owned stream */
    uint32_t flags;        /* reserved, 0 */
} ivx_config;

typedef struct ivx_program_info {
    uint32_t node_count;
    uint32_t stack_depth;  /* required_forward_stack_size */
    float domain_lo[3];
    float domain_hi[3];
} ivx_program_info;

typedef struct ivx_object_info {
    float voxel_extent;
    uint32_t grid_shape[3];
    uint32_t chunk_counts[3];
    uint32_t chunk_i_begin;   /* x-slab owned by this object (multi-GPU); [0, chunk_counts[0]) otherwise */
    uint32_t chunk_i_end;
    uint32_t n_void, n_uniform, n_non_uniform;
    uint32_t occupied_chunk_ranges[6]; /* [dim*2 + {start,end}] */
    uint32_t occupied_voxel_ranges[6];
    uint64_t device_bytes;
} ivx_object_info;

typedef struct ivx_mesh_info {
    uint32_t n_vertices;
    uint32_t n_indices;
    uint32_t n_submeshes;
    uint32_t n_exposed_chunks;
    /* device pointers, valid until the next mesh / remesh / free call on the object */
    const float* d_positions;              /* 3 * n_vertices */
    const float* d_normals;                /* 3 * n_vertices */
    const ivx_index_materials* d_index_materials; /* n_indices */
    const uint32_t* d_indices;             /* n_indices */
    const ivx_chunk_submesh* d_submeshes;  /* n_submeshes */
    const uint32_t* d_vertex_ranges;       /* 2 * n_submeshes */
} ivx_mesh_info;

typedef struct ivx_absorb_stats {
    uint32_t touched_chunks;
    uint32_t touched_voxels;
    uint32_t emptied_voxels;
    uint32_t removed_chunks;
    uint32_t dirty_chunks; /* size of invalidated_mesh_chunk_indices after the call */
} ivx_absorb_stats;

/* `VoxelObjectInertialPropertyManager` (object/inertia.rs:19-25), see "inertial properties" below */
typedef struct ivx_inertial_moments {
    float mass;
    float moments[3];
    float moments_of_inertia[3];
    float products_of_inertia[3];
} ivx_inertial_moments;

typedef struct ivx_ctx ivx_ctx;
typedef struct ivx_program ivx_program;
typedef struct ivx_object ivx_object;

/* ---- context ------------------------------------------------------------ */
int ivx_create(const ivx_config* config, ivx_ctx** out_ctx);
void ivx_destroy(ivx_ctx* ctx);
const char* ivx_last_error(const ivx_ctx* ctx);
uint32_t ivx_abi_version(void);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t ivx_kernel_launch_count(const ivx_ctx* ctx);
int ivx_synchronize(ivx_ctx* ctx);
/* Per-kernel device timing with CUDA events on the ctx stream (for bench.py's
 * roofline object). kernel ids: 0 fold (conservative), 1 fold (exact), 2 eval,
 * 3 boundary (classify+apply), 4 mesh count, 5 mesh emit, 6 absorb, 7 voxel types + flags (k_types),
 * 8 inertial moments: rows + uniform chunks, 9 inertial moments: non-uniform chunks, 10 inertial moments: ordered sum.
 * ivx_profile_get synchronises, then returns the accumulated milliseconds and
 * launch count of that kernel since the last reset. */
int ivx_profile_enable(ivx_ctx* ctx, int enabled);
int ivx_profile_reset(ivx_ctx* ctx);
int ivx_profile_get(ivx_ctx* ctx, uint32_t kernel_id, double* out_total_ms, uint64_t* out_launches);
/* work counters accumulated on the device while profiling is enabled (reset by ivx_profile_reset):
 * 0 = 4-D simplex noise evaluations performed by the voxel type kernel (voxel_type.rs:125-168). */
int ivx_profile_counter(ivx_ctx* ctx, uint32_t counter_id, uint64_t* out_value);

/* ---- graph compile ------------------------------------------------------
 * ivx_program_build  replaces SDFGraph::build_in → SDFGenerator::new_in
 *                    (atomic.rs:1031-1037, 228-493, 495-596).
 * ivx_program_upload takes an already compiled ProcessedSDFNode list (a Rust
 *                    host that keeps SDFGenerator::new_in on its side). */
int ivx_program_build(ivx_ctx* ctx, const ivx_sdf_node* nodes, uint32_t n_nodes, uint32_t root_node_id,
                      ivx_program** out_program);
int ivx_program_upload(ivx_ctx* ctx, const ivx_node* nodes, uint32_t n_nodes, uint32_t stack_depth,
                       const float domain_lo[3], const float domain_hi[3], ivx_program** out_program);
/* Host-only form of ivx_program_build (no device, no ctx): writes the compiled
 * ProcessedSDFNode list into out_nodes (capacity in nodes; the unrolled tree can
 * be larger than n_nodes) and returns IVX_ERR_CAPACITY with *out_count set when
 * it does not fit. err receives the reference's error text on IVX_ERR_GRAPH. */
int ivx_program_compile_host(const ivx_sdf_node* nodes, uint32_t n_nodes, uint32_t root_node_id,
                             ivx_node* out_nodes, uint32_t capacity, uint32_t* out_count,
                             ivx_program_info* out_info, char* err, size_t err_capacity);
int ivx_program_info_get(ivx_ctx* ctx, const ivx_program* program, ivx_program_info* out);
int ivx_program_nodes(ivx_ctx* ctx, const ivx_program* program, ivx_node* out, uint32_t capacity);
void ivx_program_free(ivx_ctx* ctx, ivx_program* program);

/* ---- meta graph compile ---------------------------------------------------
 * ivx_meta_compile replaces MetaSDFGraph::build_in (generation/sdf/meta.rs:741-896): it resolves the meta nodes
 * (instances, transforms, placements, selections, SDF instantiation and combination; all 21 `MetaSDFNode` kinds,
 * meta.rs:55-80) into the atomic `SDFNode` list that ivx_program_build takes. The host parses its file format
 * (RON in the engine) into these PODs; sampling, seeding and the emitted node order are the reference's.
 *
 * ivx_meta_source  = `ParamSource` after name resolution (meta/params.rs:40-62): kind 0 a fixed value, kind 1
 *                    `offset + scale * value of parameter idx of the same node` (value = offset).
 * ivx_meta_param   = `ContParamSpec` / `DiscreteParamSpec` (params.rs:20-38): dist 0 Constant(src[0]),
 *                    1 Uniform{min,max}, 2 UniformCosAngle{min_angle,max_angle} (degrees), 3 PowerLaw{min,max,exponent}.
 * ivx_meta_node    = one `MetaSDFNode`; params[] holds the kind's parameters in the declaration order of the reference
 *                    struct:
 *   Points                    count
 *   Spheres                   count, sampling, seed; radius, center_x, center_y, center_z
 *   Capsules                  count, sampling, seed; segment_length, radius, center_x, center_y, center_z
 *   Boxes                     count, sampling, seed; extent_x, extent_y, extent_z, center_x, center_y, center_z
 *   Translation               child[0], composition, sampling, seed; translation_x, _y, _z
 *   Rotation                  child[0], composition, sampling, seed; tilt_angle, turn_angle, roll_angle (degrees)
 *   Scaling                   child[0], composition, sampling, seed; scaling
 *   Similarity                child[0], composition, sampling, seed; scale, tilt_angle, turn_angle, roll_angle,
 *                             translation_x, _y, _z
 *   StratifiedGridTransforms  child[0], seed; shape_x, shape_y, shape_z (discrete), cell_extent_x, _y, _z, jitter_fraction
 *   SphereSurfaceTransforms   child[0], rotation, seed; radius, jitter_fraction
 *   ClosestTranslationToSurface / RotationToGradient   child[0] = surface SDF, child[1] = subject instances
 *   RayTranslationToSurface   child[0] = surface SDF, child[1] = subject instances, anchor
 *   StochasticSelection       child[0], min_pick_count, max_pick_count, pick_probability, seed
 *   SDFInstantiation          child[0]
 *   TransformApplication      child[0] = SDF or group, child[1] = instances
 *   MultifractalNoiseSDFModifier  child[0], sampling, seed; octaves (discrete), frequency, lacunarity, persistence,
 *                             amplitude
 *   SDFUnion / SDFSubtraction / SDFIntersection   child[0], child[1], smoothness
 *   SDFGroupUnion             child[0], smoothness
 * The root is the last node (meta.rs:766). */
#define IVX_META_MAX_PARAMS 8
typedef enum ivx_meta_kind {
    IVX_META_POINTS = 0,
    IVX_META_SPHERES = 1,
    IVX_META_CAPSULES = 2,
    IVX_META_BOXES = 3,
    IVX_META_TRANSLATION = 4,
    IVX_META_ROTATION = 5,
    IVX_META_SCALING = 6,
    IVX_META_SIMILARITY = 7,
    IVX_META_STRATIFIED_GRID_TRANSFORMS = 8,
    IVX_META_SPHERE_SURFACE_TRANSFORMS = 9,
    IVX_META_CLOSEST_TRANSLATION_TO_SURFACE = 10,
    IVX_META_RAY_TRANSLATION_TO_SURFACE = 11,
    IVX_META_ROTATION_TO_GRADIENT = 12,
    IVX_META_STOCHASTIC_SELECTION = 13,
    IVX_META_SDF_INSTANTIATION = 14,
    IVX_META_TRANSFORM_APPLICATION = 15,
    IVX_META_MULTIFRACTAL_NOISE = 16,
    IVX_META_SDF_UNION = 17,
    IVX_META_SDF_SUBTRACTION = 18,
    IVX_META_SDF_INTERSECTION = 19,
    IVX_META_SDF_GROUP_UNION = 20
} ivx_meta_kind;

typedef struct ivx_meta_source {
    uint32_t kind;  /* 0 Fixed, 1 FromParam */
    uint32_t idx;   /* FromParam: index into the node's params */
    float value;    /* Fixed: the value (discrete: the integer); FromParam: offset */
    float scale;    /* FromParam: scale */
} ivx_meta_source;

typedef struct ivx_meta_param {
    uint32_t dist;  /* 0 Constant, 1 Uniform, 2 UniformCosAngle, 3 PowerLaw */
    ivx_meta_source src[3];
} ivx_meta_param;

typedef struct ivx_meta_node {
    uint32_t kind;            /* ivx_meta_kind */
    uint32_t child[2];
    uint32_t count;
    uint32_t seed;
    uint32_t sampling;        /* ParameterSamplingMode: 0 OnlyOnce, 1 PerInstance */
    uint32_t composition;     /* CompositionMode: 0 Post, 1 Pre */
    uint32_t rotation;        /* SphereSurfaceRotation: 0 Identity, 1 RadialOutwards, 2 RadialInwards */
    uint32_t anchor;          /* RayTranslationAnchor: 0 Origin, 1 ShapeBoundaryAtOrigin */
    uint32_t min_pick_count;
    uint32_t max_pick_count;
    float pick_probability;
    float smoothness;
    ivx_meta_param params[IVX_META_MAX_PARAMS];
} ivx_meta_node;

/* Writes the atomic nodes to out_nodes and the root id to *out_root. *out_empty = 1 when the graph resolves to nothing
 * (`SDFGraph` without root; no nodes are written). IVX_ERR_CAPACITY with *out_count set when `capacity` is too small.
 * ctx may be NULL for graphs without ClosestTranslationToSurface / RayTranslationToSurface / RotationToGradient nodes;
 * those probe the partial SDF on the device (ivx_program_eval_blocks, all instances of a node per launch).
 * err receives the reference's error text on IVX_ERR_GRAPH. */
int ivx_meta_compile(ivx_ctx* ctx, const ivx_meta_node* nodes, uint32_t n_nodes, float scale_factor, uint64_t seed,
                     ivx_sdf_node* out_nodes, uint32_t capacity, uint32_t* out_count, uint32_t* out_root,
                     int* out_empty, char* err, size_t err_capacity);

/* SDFGenerator::compute_signed_distances_for_chunk (atomic.rs:207-216) for a
 * batch of chunks: origins = n_chunks x 3 chunk lower corners in root space;
 * out = n_chunks x 4096 f32 (host). Used by the parity tests and by the
 * meta-graph compiler's probes. */
int ivx_program_eval_chunks(ivx_ctx* ctx, const ivx_program* program, const float* chunk_origins,
                            uint32_t n_chunks, float* out_signed_distances);

/* SDFGenerator::compute_signed_distances_for_block_preserving_gradients<SIZE, SIZE^3>
 * (atomic.rs:877-998; no culling) for a batch of small blocks, SIZE = 1 or 2: origins = n_blocks x 3 block
 * lower corners in root space, out = n_blocks x SIZE^3 f32 (host). These are the probes the meta-graph compiler's
 * sphere-cast / closest-point / gradient nodes issue (meta.rs:2705-2748), batched over instances. */
int ivx_program_eval_blocks(ivx_ctx* ctx, const ivx_program* program, const float* block_origins,
                            uint32_t n_blocks, uint32_t size, float* out_signed_distances);

/* ---- object generation --------------------------------------------------
 * Replaces SDFVoxelGenerator::new (generation.rs:207-258) +
 * VoxelObject::generate / generate_in_parallel (object.rs:239-263):
 * generate_without_derived_state, update_occupied_voxel_ranges,
 * compute_all_derived_state (split detection excluded). */
int ivx_object_generate(ivx_ctx* ctx, const ivx_program* program, float voxel_extent,
                        const ivx_type_generator* type_generator, ivx_object** out_object);
/* ivx_object_generate followed by ivx_object_download, overlapped: the chunk planes are generated in parts and each
 * finished part is packed to the reference's layout (`Voxel` AoS, `data_offset` = ordinal among NonUniform chunks,
 * object.rs:574-577) and copied to the host buffers on a second stream while the next parts are still being
 * computed. `host_chunks` needs room for every chunk of the grid, `host_voxels` for 4096 voxels per NonUniform chunk
 * (IVX_ERR_CAPACITY otherwise); page-locked buffers are needed for the copies to overlap. The call returns when the
 * object is complete on the device; the host buffers are complete after ivx_synchronize (or ivx_object_free), so a
 * following ivx_object_mesh runs under the tail of the transfer. */
int ivx_object_generate_streamed(ivx_ctx* ctx, const ivx_program* program, float voxel_extent,
                                 const ivx_type_generator* types, ivx_chunk_desc* host_chunks, size_t chunk_capacity,
                                 ivx_voxel* host_voxels, size_t voxel_capacity, ivx_object** out_object,
                                 uint64_t* out_non_uniform_chunks);

/* VoxelObject::generate (object.rs:239-244 → generate_voxels_for_chunks :361-404) for a host-side
 * `ChunkedVoxelGenerator` (generation.rs:41-67) — any generator other than the SDF one, e.g. the reference's
 * OffsetBoxVoxelGenerator / ManualVoxelGenerator fixtures (object.rs:3387-3561) or voxels loaded from elsewhere: the host
 * calls its generator's generate_chunk for every chunk of the grid and hands over what it returned,
 *   voxels      4096 `Voxel`s per chunk, chunks in x-major linear order, voxels in the order i*256 + j*16 + k
 *   sparseness  one byte per chunk: bit 0 ChunkSparseness::has_only_empty_voxels, bit 1 ::is_void
 * and the device does the rest of generate(): VoxelChunk::create_for_generated_voxels (object.rs:1890-1964),
 * update_occupied_voxel_ranges and compute_all_derived_state. chunk_counts = ceil(grid_shape / 16). The result is an
 * ordinary object (mesh, absorb, connected regions, extraction, inertial moments, download). Voxels of non-void chunks
 * must satisfy the invariant of the reference's `Voxel` constructors (EMPTY flag ⇔ signed-distance code >= 0, lib.rs:
 * 300-348); IVX_ERR_INVALID_ARGUMENT otherwise. */
int ivx_object_from_generated_chunks(ivx_ctx* ctx, float voxel_extent, const uint32_t grid_shape[3],
                                     const ivx_voxel* voxels, const uint8_t* sparseness, ivx_object** out_object);

/* Multi-GPU: generate only chunk planes [chunk_i_begin, chunk_i_end) of the
 * x-major chunk grid (the reference's thread split of the linear chunk index,
 * object.rs:423-427). The object also reserves one halo chunk plane on each
 * side that has a neighbouring slab; its cross-chunk derived state
 * (object.rs:1659-1785) stays pending until the slab protocol below has run:
 *
 *   every rank:  generate_slab
 *   exchange A:  halo_export(side) → send → neighbour's halo_import(1 - side)
 *   every rank:  slab_classify           (which uniform chunks convert)
 *   exchange B:  halo_kinds_export(0) → send → lower neighbour's halo_kinds_import(1)
 *   every rank:  slab_finalize           (adjacency bits, obscuredness)
 *
 * after which ivx_object_mesh / ivx_object_download act on the owned planes
 * and give exactly the rows of the whole object. Exchange B carries the one
 * bit per chunk that the quad-ownership rule of Surface Nets reads from the +x
 * neighbour chunk (object/sdf/surface_nets.rs:252-261). Buffers are DEVICE
 * pointers (send them with NCCL or a peer copy); their layout is private to
 * this library version. side: 0 = lower chunk-i, 1 = higher.
 * ASYNCHRONOUS: halo_export, halo_import, halo_kinds_export, halo_kinds_import,
 * slab_classify and ivx_mesh_push only enqueue work on the context's stream and
 * return; order them against the transport yourself (run the transport on the
 * same stream — ivx_config.stream — or call ivx_synchronize before sending and
 * synchronise the transport's stream before importing). slab_finalize
 * synchronises. The communicator below (ivx_comm_*) needs none of this. */
int ivx_object_generate_slab(ivx_ctx* ctx, const ivx_program* program, float voxel_extent,
                             const ivx_type_generator* type_generator, uint32_t chunk_i_begin,
                             uint32_t chunk_i_end, ivx_object** out_object);
/* Work per chunk plane of the x-major chunk grid (arbitrary integer units), for partitioning the planes into slabs of
 * equal work instead of equal thickness. Planning by doing: the first request for a (program content, voxel extent, type
 * generator) generates the whole object once on the calling device and counts, per plane, the chunks the SDF program ran
 * on and the chunks that needed voxel types; the result is kept with the context, so asking again — also for a program
 * rebuilt from the same nodes — costs nothing. When the object's storage bound does not fit the device, or with
 * IVX_PLANE_WORK=estimate in the environment, the conservative program-specialisation levels alone classify 2^3-chunk
 * blocks as void / inside / undecided and weigh them. Deterministic either way, so every rank that calls it derives the
 * same partition. out_work may be NULL to query *out_planes only. */
int ivx_program_plane_work(ivx_ctx* ctx, const ivx_program* program, float voxel_extent,
                           const ivx_type_generator* type_generator, uint32_t* out_work, uint32_t capacity,
                           uint32_t* out_planes);
/* exact size of every halo message of this object in bytes: chunk descriptors of one plane + the one voxel layer
 * per chunk that touches the neighbouring slab (768 bytes); fixed, so sender and receiver need no size handshake.
 * halo_export writes exactly this many bytes (*out_bytes), halo_import expects exactly this many. */
int ivx_object_halo_capacity(ivx_ctx* ctx, const ivx_object* object, size_t* out_bytes);
int ivx_object_halo_export(ivx_ctx* ctx, const ivx_object* object, int side, void* device_buffer,
                           size_t capacity, size_t* out_bytes);
int ivx_object_halo_import(ivx_ctx* ctx, ivx_object* object, int side, const void* device_buffer,
                           size_t bytes);
int ivx_object_slab_classify(ivx_ctx* ctx, ivx_object* object);
/* one byte per chunk of a plane (chunk_counts[1] * chunk_counts[2] bytes) */
int ivx_object_halo_kinds_export(ivx_ctx* ctx, const ivx_object* object, int side, void* device_buffer,
                                 size_t capacity);
int ivx_object_halo_kinds_import(ivx_ctx* ctx, ivx_object* object, int side, const void* device_buffer,
                                 size_t bytes);
int ivx_object_slab_finalize(ivx_ctx* ctx, ivx_object* object);
int ivx_object_info_get(ivx_ctx* ctx, const ivx_object* object, ivx_object_info* out);
/* → the Rust-side VoxelObject: chunks[C] in x-major linear order and the voxels
 * of the NonUniform chunks, data_offset = ordinal in that order. */
int ivx_object_download(ivx_ctx* ctx, const ivx_object* object, ivx_chunk_desc* chunks,
                        size_t chunk_capacity, ivx_voxel* voxels, size_t voxel_capacity);
void ivx_object_free(ivx_ctx* ctx, ivx_object* object);

/* ---- meshing ------------------------------------------------------------
 * ivx_object_mesh replaces VoxelObjectMesh::create / recreate (mesh.rs:280-354):
 * for_each_exposed_chunk_with_sdf + compute_surface_nets_mesh + append. */
int ivx_object_mesh(ivx_ctx* ctx, ivx_object* object, ivx_mesh_info* out);
/* ivx_object_download with the transfer left running on the context's copy stream, so that meshing (or anything
 * else queued afterwards) overlaps it; the host buffers — pinned memory, or the copies serialise — are complete after
 * ivx_synchronize. *out_non_uniform_chunks = number of 4096-voxel blocks written. */
int ivx_object_download_async(ivx_ctx* ctx, ivx_object* object, ivx_chunk_desc* chunks, size_t chunk_capacity,
                              ivx_voxel* voxels, size_t voxel_capacity, uint64_t* out_non_uniform_chunks);
int ivx_mesh_download(ivx_ctx* ctx, const ivx_object* object, float* positions, float* normals,
                      ivx_index_materials* index_materials, uint32_t* indices,
                      ivx_chunk_submesh* submeshes, uint32_t* vertex_ranges);
/* The same for a caller that sized its buffers from an earlier ivx_mesh_info: IVX_ERR_INVALID_ARGUMENT (nothing copied)
 * when the object's mesh no longer has those sizes — it was re-created, patched (ivx_object_remesh_dirty) or synced since. */
int ivx_mesh_download_checked(ivx_ctx* ctx, const ivx_object* object, uint32_t n_vertices, uint32_t n_indices,
                              uint32_t n_submeshes, float* positions, float* normals, ivx_index_materials* index_materials,
                              uint32_t* indices, ivx_chunk_submesh* submeshes, uint32_t* vertex_ranges);

/* ---- multi-GPU mesh gather over peer memory ------------------------------
 * VoxelObjectMesh::recreate appends the chunk meshes in linear chunk order (mesh.rs:286-354); with x-slabs on several
 * GPUs that is the concatenation of the slabs' meshes with vertex / index offsets rebased. Instead of sending the six
 * arrays through NCCL, every rank writes its part directly into the merged mesh in the gathering GPU's memory
 * (NVLink / NVSwitch peer stores), rebasing on the fly:
 *   gathering rank: ivx_peer_alloc → 64-byte IPC handle, sent to the other ranks once (any transport)
 *   other ranks:    ivx_peer_open(handle) → pointer valid on their device
 *   every step:     ivx_object_mesh, exchange the three counts, then
 *                   ivx_mesh_push(base, field_offsets, vertex_base, index_base, submesh_base)   [asynchronous]
 *                   and a barrier on the same stream before the gathering rank reads the result.
 * field_offsets[6]: byte offsets of positions, normals, indices, index_materials, submeshes, vertex_ranges inside
 * the merged block (16-byte aligned). Fails with IVX_ERR_UNSUPPORTED where CUDA IPC is not available. */
int ivx_peer_alloc(ivx_ctx* ctx, size_t bytes, void** out_device_ptr, unsigned char out_handle[64]);
int ivx_peer_free(ivx_ctx* ctx, void* device_ptr);
int ivx_peer_open(ivx_ctx* ctx, const unsigned char handle[64], void** out_device_ptr);
int ivx_peer_close(ivx_ctx* ctx, void* device_ptr);
int ivx_mesh_push(ivx_ctx* ctx, const ivx_object* object, void* merged_base, const uint64_t field_offsets[6],
                  uint32_t vertex_base, uint32_t index_base, uint32_t submesh_base);

/* ---- the mesh kept in sync with a modified object --------------------------
 * ivx_object_mesh_sync replaces VoxelObjectMesh::sync_with_voxel_object (mesh.rs:360-456) including its
 * ChunkSubmeshManager / RangeAllocator (mesh.rs:703-848): the mesh ivx_object_mesh created stays on the device; the
 * invalidated chunks are re-meshed and each one is written into the smallest free vertex / index range that fits, else
 * appended; chunks that are no longer exposed (or mesh to nothing) lose their submesh (swap-remove, like the reference's
 * table). `out` describes the mesh afterwards: n_vertices / n_indices are the buffer LENGTHS (freed ranges included),
 * the submesh table and the vertex ranges are the manager's, in its order. Invalidated chunks are visited in ascending
 * linear chunk index (the reference iterates a HashSet: any order is the reference's). Clears the invalidation marks.
 * ivx_mesh_modifications = VoxelObjectMesh::mesh_modifications (mesh.rs:105-118, 833-838): the vertex / index ranges
 * written since the last report (4 words per record: vertex start, end, index start, end) — what a renderer uploads —
 * and whether submeshes were removed; ivx_mesh_report_synchronized = report_gpu_resources_synchronized. */
int ivx_object_mesh_sync(ivx_ctx* ctx, ivx_object* object, ivx_mesh_info* out);
int ivx_mesh_modifications(ivx_ctx* ctx, const ivx_object* object, uint32_t* out_ranges, size_t capacity_records,
                           uint64_t* out_count, int* out_chunks_were_removed);
int ivx_mesh_report_synchronized(ivx_ctx* ctx, ivx_object* object);

/* ---- the mesh buffers a renderer draws from -------------------------------
 * VoxelMeshGPUBuffers (gpu_resource.rs:460-900): the five buffers of a meshed voxel object — vertex positions, normal
 * vectors, index materials, indices, chunk submeshes — created from the object's mesh (for_voxel_object, :484-600) and
 * brought up to date after a mesh sync by writing only the updated ranges (sync_with_voxel_object, :714-900). The
 * reference stages those ranges from host memory into wgpu buffers; here the mesh is on the device already, so each
 * buffer is an exportable device allocation: `fd` is a POSIX file descriptor a graphics API imports ONCE as external
 * memory (VK_KHR_external_memory_fd, handle type OPAQUE_FD, allocation_bytes; cuMemImportFromShareableHandle for a CUDA
 * consumer), and a sync is device-to-device copies on the context's stream, no host round trip.
 *   ivx_mesh_gpu_buffers_create   for_voxel_object: the object's current mesh (ivx_object_mesh or the synced mesh), whole.
 *   ivx_mesh_gpu_buffers_sync     sync_with_voxel_object, call it after ivx_object_mesh_sync: nothing when there are no
 *                                 modifications; the updated vertex / index ranges (ivx_mesh_modifications) while the
 *                                 slices fit their buffers; a buffer pair whose slice outgrew it is re-created with the
 *                                 whole slice (`recreated` = 1 and a NEW `fd` to import); the chunk submesh table is
 *                                 rewritten whenever anything changed; ends with ivx_mesh_report_synchronized, like the
 *                                 reference. A mesh that was re-created by ivx_object_mesh since is copied whole.
 * Descriptors are handed over: the caller closes every `fd` >= 0 it receives (importing does not consume it). The
 * copies are asynchronous on the context's stream; a consumer on another queue orders itself after ivx_synchronize or
 * an event on that stream. Element layouts: positions / normals 3 x f32, ivx_index_materials, u32,
 * ivx_chunk_submesh — the reference's buffer contents byte for byte. */
enum {
    IVX_MESH_BUFFER_POSITIONS = 0,
    IVX_MESH_BUFFER_NORMALS = 1,
    IVX_MESH_BUFFER_INDEX_MATERIALS = 2,
    IVX_MESH_BUFFER_INDICES = 3,
    IVX_MESH_BUFFER_CHUNK_SUBMESHES = 4,
    IVX_MESH_BUFFER_COUNT = 5
};
typedef struct ivx_mesh_gpu_buffers ivx_mesh_gpu_buffers;
typedef struct ivx_mesh_gpu_buffer_info {
    int32_t fd;                /* >= 0: a descriptor of the (new) allocation, owned by the caller; -1: unchanged */
    uint32_t recreated;        /* 1 when this call made the allocation */
    uint64_t allocation_bytes; /* size of the allocation (what an importer maps) */
    uint64_t valid_bytes;      /* bytes of mesh data in it (n_valid_bytes) */
    void* device_ptr;          /* the buffer in this process, for CUDA consumers */
} ivx_mesh_gpu_buffer_info;
typedef struct ivx_mesh_gpu_buffers_info {
    ivx_mesh_gpu_buffer_info buffer[IVX_MESH_BUFFER_COUNT];
    uint64_t n_vertices, n_indices, n_chunks; /* lengths of the mesh slices (freed ranges included), submesh rows */
    uint64_t bytes_copied;                    /* device-to-device bytes this call moved */
    uint32_t n_updated_ranges;                /* records of ivx_mesh_modifications this call consumed */
    uint32_t reserved;
} ivx_mesh_gpu_buffers_info;
int ivx_mesh_gpu_buffers_create(ivx_ctx* ctx, ivx_object* object, ivx_mesh_gpu_buffers** out, ivx_mesh_gpu_buffers_info* info);
int ivx_mesh_gpu_buffers_sync(ivx_ctx* ctx, ivx_object* object, ivx_mesh_gpu_buffers* buffers, ivx_mesh_gpu_buffers_info* info);
void ivx_mesh_gpu_buffers_destroy(ivx_ctx* ctx, ivx_mesh_gpu_buffers* buffers);

/* ---- collision probes ------------------------------------------------------
 * VoxelObjectCollisionProbes (collidable.rs:97-101, 346-780): per meshed chunk and per block of 1^3 .. 8^3 voxels
 * (determine_log2_block_size_for_object, :451-471) the mesh vertex of lowest mean curvature — the points the physics
 * probes other objects with. They live beside the mesh (MeshedVoxelObject, mesh.rs:36-44):
 *   ivx_object_collision_probes       replaces compute_for_all_chunks / recompute_for_all_chunks (:355-392, MeshedVoxelObject::
 *                                     create, mesh.rs:156-168) on the object's current mesh (ivx_object_mesh or the synced mesh)
 *   ivx_object_collision_probes_sync  replaces sync_with_voxel_object_and_mesh (:394-433, 524-612) — call it right after
 *                                     ivx_object_mesh_sync (MeshedVoxelObject::sync_mesh_with_object, mesh.rs:193-205): the
 *                                     chunks that sync visited are re-probed and their points go into the smallest free
 *                                     range that fits, else to the end (the reference's RangeAllocator), in the same order
 *   ivx_collision_probes_download     probe_points() (buffer length n_points, freed ranges keep obsolete points) and
 *                                     chunk_point_ranges(), here in ascending linear chunk index
 * d_points stays valid until the next probes call on the object. */
typedef struct ivx_probes_info {
    uint32_t log2_block_size;
    uint32_t _pad;
    uint64_t n_points;   /* length of the point buffer */
    uint64_t n_chunks;   /* chunks that have points */
    const float* d_points;  /* device, 3 floats per point */
} ivx_probes_info;
typedef struct ivx_probe_range {
    uint32_t chunk_indices[3];
    uint32_t point_start, point_end;
} ivx_probe_range;
int ivx_object_collision_probes(ivx_ctx* ctx, ivx_object* object, ivx_probes_info* out);
int ivx_object_collision_probes_sync(ivx_ctx* ctx, ivx_object* object, ivx_probes_info* out);
int ivx_collision_probes_download(ivx_ctx* ctx, const ivx_object* object, float* points, size_t capacity_points,
                                  ivx_probe_range* ranges, size_t capacity_ranges);

/* ---- multi-GPU communicator over peer memory ---------------------------------
 * The whole multi-GPU plane of the path behind the C ABI — a host needs no NCCL and no torch for it:
 *
 *   once:        ivx_comm_create on every rank → 64-byte handle; exchange the handles by any transport (MPI, a socket,
 *                torch.distributed ...); ivx_comm_connect(all handles in rank order). Ranks living in one process use
 *                ivx_comm_connect_local instead.
 *   every step:  ivx_object_generate_slab(planes of this rank)
 *                ivx_object_exchange_halos(comm, object, lower rank, upper rank)   [asynchronous on the ctx stream]
 *                ivx_object_mesh_gather(comm, object, &local, &merged)            [synchronises the ctx stream]
 *
 * exchange_halos replaces the explicit slab protocol above (halo_export ... slab_finalize): boundary planes and the
 * quad-ownership bits are stored straight into the neighbour's window over NVLink and published with a flag the
 * neighbour's stream waits for on the device (V/object.rs:423-427 for the split, :1682-1704 and V/object/sdf.rs:410-428
 * for what the neighbour plane is needed for, V/object/sdf/surface_nets.rs:252-261 for the bits). mesh_gather meshes
 * the slab (ivx_object_mesh), publishes its sizes to all ranks, takes its place in the merged mesh from the sizes
 * of the lower ranks (slab order = the reference's linear chunk order, V/mesh.rs:286-354) and stores its part, rebased,
 * into the gather rank's window. On the gather rank `merged` describes the result: device pointers into the window,
 * valid until the gather of the step after the next one. Every rank must make both calls once per step, in the same
 * order. lower_rank / upper_rank: the ranks owning the planes just below / above this rank's (empty slabs skipped),
 * -1 for none. The mesh capacities are those of the merged mesh of the whole job; IVX_ERR_CAPACITY (on the rank whose
 * part does not fit, and on the gather rank) if a step exceeds them — create a larger communicator then. */
typedef struct ivx_comm ivx_comm;
typedef struct ivx_comm_config {
    uint32_t rank, world, gather_rank;
    uint32_t plane_chunks;  /* chunk_counts[1] * chunk_counts[2] of the objects exchanged */
    uint64_t mesh_vertices, mesh_indices, mesh_submeshes;
} ivx_comm_config;
typedef struct ivx_gathered_mesh {
    uint64_t n_vertices, n_indices, n_submeshes;
    void* d_positions;        /* 3 x f32 per vertex */
    void* d_normals;
    void* d_indices;          /* u32 per index */
    void* d_index_materials;  /* ivx_index_materials per index */
    void* d_submeshes;        /* ivx_chunk_submesh */
    void* d_vertex_ranges;    /* 2 x u32 per submesh */
} ivx_gathered_mesh;
int ivx_comm_create(ivx_ctx* ctx, const ivx_comm_config* config, ivx_comm** out_comm, unsigned char out_handle[64]);
int ivx_comm_connect(ivx_ctx* ctx, ivx_comm* comm, const unsigned char* all_handles /* world x 64 bytes, rank order */);
int ivx_comm_connect_local(ivx_ctx* ctx, ivx_comm* comm, ivx_comm* const* all_comms /* world pointers, rank order */);
void ivx_comm_destroy(ivx_ctx* ctx, ivx_comm* comm);
int ivx_object_exchange_halos(ivx_ctx* ctx, ivx_comm* comm, ivx_object* object, int lower_rank, int upper_rank);
int ivx_object_mesh_gather(ivx_ctx* ctx, ivx_comm* comm, ivx_object* object, ivx_mesh_info* out_local,
                           ivx_gathered_mesh* out_merged);
/* The same step with the mesh left where it was made ("or leave the mesh distributed and return per-rank
 * views"): only the mesh sizes travel; every rank's part is rebased in place to the numbering it has in the mesh of the
 * whole job (vertex indices and vertex ranges + out_bases[0], ChunkSubmesh::index_offset + out_bases[1]; out_bases[2] =
 * submeshes of the lower ranks) and described by out_local. For hosts that consume the parts rank by rank — e.g. each
 * rank downloading its part over its own PCIe link (ivx_mesh_download) instead of one rank downloading everything.
 * Takes the place of ivx_object_mesh_gather in a step; call it once per object. */
int ivx_object_mesh_distributed(ivx_ctx* ctx, ivx_comm* comm, ivx_object* object, ivx_mesh_info* out_local, uint64_t out_bases[3]);

/* ---- modification -------------------------------------------------------
 * ivx_object_absorb_sphere replaces apply_sphere_absorption
 * (interaction/absorption.rs:801-844) → modify_voxels_within_sphere
 * (object/intersection.rs:283-394): centre / radii in normalized voxel space
 * (voxel extent 1, grid lower corner at the origin).
 * ivx_object_remesh_dirty replaces VoxelObjectMesh::sync_with_voxel_object
 * (mesh.rs:360-456) for the invalidated chunk set and clears it. */
int ivx_object_absorb_sphere(ivx_ctx* ctx, ivx_object* object, const float center[3], float radius,
                             float influence_radius, ivx_absorb_stats* out_stats);
/* ivx_object_absorb_capsule replaces apply_capsule_absorption (interaction/absorption.rs:846-889) →
 * modify_voxels_within_capsule (object/intersection.rs:417-537): the influence capsule is given by its segment start
 * and segment vector (Capsule::new, impact_geometry/src/capsule.rs:61) and `influence_radius`; voxels whose centre
 * lies within or on it get max(sd, -(distance_to_segment - radius)). Negative radii return
 * IVX_ERR_INVALID_ARGUMENT (the reference asserts). */
int ivx_object_absorb_capsule(ivx_ctx* ctx, ivx_object* object, const float segment_start[3],
                              const float segment_vector[3], float radius, float influence_radius,
                              ivx_absorb_stats* out_stats);
/* ivx_objects_absorb_mutually replaces apply_mutual_absorption (interaction/absorption.rs:891-1080): two overlapping
 * voxel objects subtract each other's volume. Object A's voxels in the intersection ranges (padded by
 * ceil(extent_b / extent_a) voxels, the reference's snapshot ranges) sample B's signed distance field trilinearly
 * (sample_voxel_object_sdf, object/sdf.rs:636-675) at their centre carried into B's frame; B's voxels sample a snapshot
 * of A's distances taken before A was modified; both get sdf_subtraction(sd, max(sd, other), smoothness)
 * (compute_subtracted_signed_distance, absorption.rs:1082-1094) through Voxel::set_signed_distance, then
 * modify_voxels_within_ranges' bookkeeping (object/intersection.rs:167-261: internal state, removed chunks, invalidated
 * meshes, boundary refresh).
 *   transform_from_b_to_a   Isometry3 = transform_from_world_to_a * transform_from_world_to_b.inverted()
 *   ranges_in_a / _in_b     [dim * 2 + {start, end}]: VoxelObject::determine_voxel_ranges_encompassing_intersection
 *                           (object/intersection.rs:707-745), a function of the two occupied voxel ranges
 *                           (ivx_object_info) and the transform that stays with the host's geometry code; when it
 *                           returns None the host does not call
 *   inout_a / inout_b       NULL, or both objects' inertial moments: updated like ivx_object_absorb_*_inertial
 * The rotation follows glam's Quat::mul_vec3a operation order (third party, restated). */
typedef struct ivx_isometry {
    float rotation[4];    /* unit quaternion x, y, z, w */
    float translation[3];
} ivx_isometry;
int ivx_objects_absorb_mutually(ivx_ctx* ctx, ivx_object* object_a, ivx_object* object_b,
                                const ivx_isometry* transform_from_b_to_a, float smoothness,
                                const uint32_t ranges_in_a[6], const uint32_t ranges_in_b[6],
                                const float* voxel_type_densities, uint32_t n_densities,
                                ivx_inertial_moments* inout_a, ivx_inertial_moments* inout_b,
                                ivx_absorb_stats* stats_a, ivx_absorb_stats* stats_b);
/* Host-only helpers for the call above (no device, no ctx):
 * ivx_intersection_voxel_ranges replaces VoxelObject::determine_voxel_ranges_encompassing_intersection
 * (object/intersection.rs:707-745) from the two objects' occupied voxel ranges ([dim * 2 + {start, end}], as in
 * ivx_object_info) and voxel extents; *out_intersect = 0 is the reference's `None`.
 * ivx_box_intersection_bounds is its core, compute_box_intersection_bounds (impact_geometry/src/oriented_box.rs:315-431):
 * the bounds of the overlap of an axis-aligned box A and an oriented box B (centre, unit quaternion x, y, z, w, half
 * extents), in A's frame and in B's own frame relative to B's centre, each as lower xyz, upper xyz. */
int ivx_intersection_voxel_ranges(const uint32_t occupied_a[6], float voxel_extent_a, const uint32_t occupied_b[6],
                                  float voxel_extent_b, const ivx_isometry* transform_from_b_to_a,
                                  uint32_t out_ranges_in_a[6], uint32_t out_ranges_in_b[6], int* out_intersect);
int ivx_box_intersection_bounds(const float a_lower[3], const float a_upper[3], const float b_center[3],
                                const float b_orientation[4], const float b_half_extents[3], float out_in_a[6],
                                float out_in_b[6], int* out_intersect);
int ivx_object_dirty_chunks(ivx_ctx* ctx, const ivx_object* object, uint32_t* out_linear_indices,
                            uint32_t capacity, uint32_t* out_count);
int ivx_object_remesh_dirty(ivx_ctx* ctx, ivx_object* object, ivx_mesh_info* out);

/* ---- connected regions ("split detection") -------------------------------
 * ivx_object_resolve_connected_regions replaces, on the object's current state,
 *   update_local_connected_regions_for_all_chunks  (object/split_detection.rs:305-317, 662-893)
 *   the connection updates of update_mutual_face_adjacencies (split_detection.rs:1046-1326, 1424-1463)
 *   resolve_connected_regions_between_all_chunks   (split_detection.rs:323-488)
 *   count_regions / find_two_disconnected_regions  (split_detection.rs:193-301)
 *   and the choice made by extract_smallest_region_with_property_transferrer
 *   (object/extraction.rs:121-281: fewest non-uniform chunks, then fewest chunks).
 * Local labels follow the reference's numbering (boundary regions first, in its face traversal order), and
 * the global pass visits chunks and regions in its order, so the representative (root) region of every
 * global region — hence `candidates[].label` — is the reference's. Whole objects only (a multi-GPU slab
 * object gathers on one rank first). Returns IVX_ERR_UNSUPPORTED for chunks with more local
 * regions or adjacent-region connections than the reference's fixed capacities (its asserts / overwrites,
 * split_detection.rs:798, 835, 1519-1546).
 * GlobalRegionLabel = linear chunk index << 8 | local region index (split_detection.rs:1550-1571). */
typedef struct ivx_region_candidate {
    uint32_t label;                   /* root GlobalRegionLabel */
    uint32_t chunk_count;             /* non-void chunks containing voxels of the region */
    uint32_t non_uniform_chunk_count;
    uint32_t chunk_min[3], chunk_max[3]; /* inclusive chunk index range */
} ivx_region_candidate;
typedef struct ivx_split_info {
    uint32_t n_regions;               /* count_regions */
    uint32_t has_two;                 /* find_two_disconnected_regions().is_some() */
    ivx_region_candidate candidates[2];
    uint32_t smallest;                /* index into candidates of the region an extraction would split off */
    uint32_t n_local_regions;         /* total local regions (entries of region_roots) */
    uint32_t n_connections;           /* distinct cross-chunk region connections found */
    uint32_t n_relabelled_chunks;     /* chunks whose local labels were recomputed (modified since the last resolve) */
    float device_ms;                  /* device time of the labelling + connection kernels */
    float host_ms;                    /* host time of the chunk-level union-find */
} ivx_split_info;
typedef struct ivx_chunk_regions {
    uint16_t region_count;            /* NonUniformChunkSplitDetectionData (split_detection.rs:77-84); uniform: 1 */
    uint16_t boundary_region_count;
    uint32_t first_region;            /* index of the chunk's region 0 in region_roots */
} ivx_chunk_regions;
int ivx_object_resolve_connected_regions(ivx_ctx* ctx, ivx_object* object, ivx_split_info* out);
/* Results of the last resolve: per-voxel LocalRegionLabels (4096 per NonUniform chunk in linear chunk
 * order like ivx_object_download; 255 = empty), per-chunk region counts, and the resolved root
 * GlobalRegionLabel of every local region. Any pointer may be NULL. */
int ivx_object_split_detection_download(ivx_ctx* ctx, const ivx_object* object, uint8_t* voxel_labels,
                                        size_t label_capacity, ivx_chunk_regions* per_chunk,
                                        size_t chunk_capacity, uint32_t* region_roots,
                                        size_t region_capacity);

/* ---- disconnected-region extraction --------------------------------------
 * ivx_object_extract_disconnected_region replaces VoxelObject::extract_any_disconnected_region
 * (object/extraction.rs:78-113 → extract_smallest_region… :121-281 → extract_disconnected_region :297-600 →
 * complete_extracted_voxel_object :1902-2187): connected regions are resolved, and if there are at least two, the
 * smaller of the first two found (fewest NonUniform chunks, then fewest chunks) leaves `object` as a new object whose
 * chunk grid is the bounding box of the region's chunks. Chunks holding only that region move whole, chunks shared with
 * other regions are split voxel by voxel (empty voxels are copied to both sides). A fragment with fewer than 8
 * non-empty voxels and no uniform chunk is dropped (`discarded`); one of at most 2 x 2 x 2 chunks whose voxels span at
 * most 14 per axis is re-packed into a single chunk. Both objects leave with their derived state (adjacencies,
 * obscuredness, occupied ranges) up to date; `object`'s chunks that lost voxels are marked for re-meshing
 * (ivx_object_remesh_dirty). The PropertyTransferrer's bookkeeping is replaced by ivx_object_inertial_moments on both
 * objects after the call. */
typedef struct ivx_extraction_info {
    uint32_t n_regions_before;          /* count_regions of `object` */
    uint32_t found_two;                 /* find_two_disconnected_regions().is_some() */
    uint32_t extracted;                 /* ExtractionResult::Extracted */
    uint32_t discarded;                 /* removed from `object` but too small to keep (NotExtracted) */
    uint32_t single_chunk;              /* re-packed into one chunk */
    uint32_t region_label;              /* GlobalRegionLabel of the extracted region */
    uint32_t region_chunks;             /* chunks of `object` the region touched */
    uint32_t moved_non_empty_voxels;    /* non-empty voxels that left through NonUniform chunks */
    uint32_t origin_offset_in_parent[3];/* ExtractedVoxelObject::origin_offset_in_parent, voxels */
} ivx_extraction_info;
int ivx_object_extract_disconnected_region(ivx_ctx* ctx, ivx_object* object, ivx_extraction_info* out_info,
                                           ivx_object** out_extracted);

/* ---- surface voxel queries ---------------------------------------------------
 * What collision detection and the interaction code ask an object for (collidable.rs, interaction/): the non-empty voxels
 * with at least one exposed face inside some voxel ranges, in the order the reference's closure would see them (chunks
 * i → j → k, voxels i → j → k inside the chunk's part of the ranges).
 *   ivx_object_surface_voxels_in_ranges       VoxelObject::for_each_surface_voxel_in_voxel_ranges (object/intersection.rs:
 *                                             97-151); with the occupied voxel ranges: for_each_surface_voxel (:87-95)
 *   ..._touching_sphere / _touching_capsule   for_each_surface_voxel_maybe_intersecting_sphere / _capsule (:51-85): the
 *                                             ranges of the shape's box clipped to the occupied ranges; shape given in
 *                                             normalized voxel space (voxel extent 1), like the absorption calls
 * placement: VoxelSurfacePlacement (lib.rs:109-114) 0 Face (5 blocked faces), 1 Edge (4), 2 Corner (<= 3).
 * *out_count is always the number found; IVX_ERR_CAPACITY when `capacity` is smaller (nothing is written). */
typedef struct ivx_surface_voxel {
    uint32_t indices[3];  /* object voxel indices */
    ivx_voxel voxel;
    uint8_t placement;
} ivx_surface_voxel;
int ivx_object_surface_voxels_in_ranges(ivx_ctx* ctx, const ivx_object* object, const uint32_t ranges[6],
                                        ivx_surface_voxel* out, size_t capacity, uint64_t* out_count);
/* ivx_object_sphere_contacts replaces for_each_sphere_voxel_object_contact (collidable.rs:1097-1127): the sphere query with
 * its closure fused in — every surface voxel is a sphere of radius -signed_distance * voxel_extent (compute_voxel_radius,
 * collidable.rs:1453-1455) around its centre, carried into the sphere's space by the inverse of
 * `transform_to_object_space`, and determine_sphere_sphere_contact_geometry (impact_physics/src/collision/collidable/
 * sphere.rs:105-136) decides whether there is a contact and where. `center` / `radius`: the sphere in the space the
 * transform starts from (world); contacts come in the order the reference's closure `f` would have been called. */
typedef struct ivx_voxel_contact {
    uint32_t indices[3];       /* object voxel indices */
    float position[3];         /* ContactGeometry::position */
    float surface_normal[3];   /* ContactGeometry::surface_normal */
    float penetration_depth;
} ivx_voxel_contact;
int ivx_object_sphere_contacts(ivx_ctx* ctx, const ivx_object* object, const ivx_isometry* transform_to_object_space,
                               const float center[3], float radius, ivx_voxel_contact* out, size_t capacity,
                               uint64_t* out_count);
/* for_each_voxel_object_plane_contact (collidable.rs:1176-1209): the plane { x : unit_normal . x = displacement } in the
 * space transform_to_object_space starts from; only corner voxels make contacts (determine_sphere_plane_contact_geometry,
 * impact_physics sphere.rs:138-156). Same records, same order as the closure's calls. */
int ivx_object_plane_contacts(ivx_ctx* ctx, const ivx_object* object, const ivx_isometry* transform_to_object_space,
                              const float unit_normal[3], float displacement, ivx_voxel_contact* out, size_t capacity,
                              uint64_t* out_count);
/* for_each_capsule_voxel_object_contact (collidable.rs:1257-1288, determine_capsule_sphere_contact_geometry, impact_physics
 * capsule.rs:212-270): the capsule (segment start, segment vector, radius) in the space the transform starts from. */
int ivx_object_capsule_contacts(ivx_ctx* ctx, const ivx_object* object, const ivx_isometry* transform_to_object_space,
                                const float segment_start[3], const float segment_vector[3], float radius,
                                ivx_voxel_contact* out, size_t capacity, uint64_t* out_count);
/* for_each_mutual_voxel_object_contact (collidable.rs:859-1050) for two objects that both have collision probes
 * (ivx_object_collision_probes): the probes of A inside the box of the intersection are sampled in B's distance field
 * (determine_sdf_value_and_normal_at_point_if_intersecting, :1288-1440: trilinear value → penetration depth, gradient →
 * normal; centre of mass direction where the field is clamped), then B's probes in A with the normal flipped. The shim
 * keeps `transform_from_b_to_a = world_to_a * world_to_b.inverted()` and the ranges
 * (ivx_intersection_voxel_ranges, like for ivx_objects_absorb_mutually) and returns early when there is no intersection.
 * Only mass and moments of the inertial managers are read (derive_center_of_mass). Records: indices = voxel of the
 * probing object the probe lies in (the reference's contact id is [0, i, j, k]); first the out_count_a_against_b
 * contacts of A's probes, then B's; within each, ascending chunk and point order (the reference walks a HashMap).
 * IVX_ERR_CAPACITY with both counts set when `capacity` is too small. */
int ivx_objects_mutual_contacts(ivx_ctx* ctx, const ivx_object* object_a, const ivx_object* object_b,
                                const ivx_isometry* world_to_a, const ivx_isometry* world_to_b, const uint32_t ranges_in_a[6],
                                const uint32_t ranges_in_b[6], const ivx_inertial_moments* inertial_a,
                                const ivx_inertial_moments* inertial_b, ivx_voxel_contact* out, size_t capacity,
                                uint64_t* out_count_a_against_b, uint64_t* out_count_b_against_a);
int ivx_object_surface_voxels_touching_sphere(ivx_ctx* ctx, const ivx_object* object, const float center[3], float radius,
                                              ivx_surface_voxel* out, size_t capacity, uint64_t* out_count);
int ivx_object_surface_voxels_touching_capsule(ivx_ctx* ctx, const ivx_object* object, const float segment_start[3],
                                               const float segment_vector[3], float radius, ivx_surface_voxel* out,
                                               size_t capacity, uint64_t* out_count);
/* for_each_surface_voxel_maybe_intersecting_negative_halfspace_of_plane (object/intersection.rs:30-40): the plane
 * { x : unit_normal . x = displacement } in normalized voxel space; the ranges are the occupied box fitted to the negative
 * halfspace (AxisAlignedBox::projected_onto_negative_halfspace, impact_geometry/src/axis_aligned_box.rs:460-488), which
 * ivx_voxel_ranges_within_plane (host only, voxel_ranges_within_plane :751-761) also returns on its own. */
int ivx_object_surface_voxels_within_plane(ivx_ctx* ctx, const ivx_object* object, const float unit_normal[3],
                                           float displacement, ivx_surface_voxel* out, size_t capacity, uint64_t* out_count);
int ivx_voxel_ranges_within_plane(const uint32_t occupied[6], const float unit_normal[3], float displacement,
                                  uint32_t out_ranges[6]);

/* ---- inertial properties ----------------------------------------------------
 * `VoxelObjectInertialPropertyManager` (object/inertia.rs:19-25): mass, moments (m x, m y, m z), moments of inertia
 * (diagonal of the inertia tensor) and products of inertia (m x y, m y z, m z x) integrated over the non-empty voxels
 * with respect to the origin of the voxel grid.
 * ivx_object_inertial_moments replaces VoxelObjectInertialPropertyManager::initialized_from (inertia.rs:125-137 →
 * compute_inertial_property_moments_for_object :754-789, compute_moments_for_non_uniform_chunk :629-706,
 * compute_moments_for_uniform_chunk :710-752) and returns the same f32 bits: the additions are made in the reference's
 * order (voxels i → j → k inside a chunk, chunk terms in linear chunk order).
 *   voxel_type_densities[n_densities]  mass density per voxel type (a non-empty voxel whose type has none is
 *                                      IVX_ERR_INVALID_ARGUMENT; the reference panics on the slice index)
 *   initial            NULL, or the sums to continue from: a slab object (multi-GPU) sums only its own chunk planes, so
 *                      passing rank r-1's result to rank r reproduces the whole object's chain exactly
 *   per_chunk_terms    NULL, or room for ten floats per owned chunk in linear chunk order (zero for void chunks)
 * The derived quantities (centre of mass, inertia tensor about it: derive_inertial_properties, inertia.rs:160-167)
 * are a handful of host flops on these ten numbers and stay with the host. */
int ivx_object_inertial_moments(ivx_ctx* ctx, const ivx_object* object, const float* voxel_type_densities,
                                uint32_t n_densities, const ivx_inertial_moments* initial, ivx_inertial_moments* out,
                                float* per_chunk_terms, size_t per_chunk_capacity);
/* ivx_object_absorb_sphere / _capsule with the reference's VoxelObjectInertialPropertyUpdater attached
 * (apply_sphere_absorption / apply_capsule_absorption, interaction/absorption.rs:801-889: the closure calls
 * remove_voxel(object_voxel_indices, voxel_type) for every voxel that goes from non-empty to empty, object/inertia.rs:
 * 374-395 → compute_moments_for_voxel :591-625). `inout_moments` leaves with the same f32 bits as the reference's
 * manager: the emptied voxels' terms are computed in parallel, then subtracted one after the other in the order the
 * reference visits them (chunks of the touched range i → j → k, voxels i → j → k inside a chunk). If an emptied voxel's
 * type has no density the call fails with IVX_ERR_INVALID_ARGUMENT after the voxels were modified (the reference
 * panics at the same point); `inout_moments` is then left untouched. */
int ivx_object_absorb_sphere_inertial(ivx_ctx* ctx, ivx_object* object, const float center[3], float radius,
                                      float influence_radius, const float* voxel_type_densities, uint32_t n_densities,
                                      ivx_inertial_moments* inout_moments, ivx_absorb_stats* out_stats);
int ivx_object_absorb_capsule_inertial(ivx_ctx* ctx, ivx_object* object, const float segment_start[3],
                                       const float segment_vector[3], float radius, float influence_radius,
                                       const float* voxel_type_densities, uint32_t n_densities,
                                       ivx_inertial_moments* inout_moments, ivx_absorb_stats* out_stats);

#ifdef __cplusplus
}
#endif
#endif /* IMPACT_VOXEL_CUDA_H */
