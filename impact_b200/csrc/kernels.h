// Kernel argument blocks and host launch wrappers (internal).
#pragma once
#include <cuda_runtime.h>

#include "ivx_internal.h"

namespace ivx {

// ---- generate.cu -----------------------------------------------------------
struct FoldArgs {
    const ivx_node* nodes;
    GenParams gp;
    // this level's blocks
    uint32_t n_blocks;
    uint32_t nb[3];           // blocks per axis at this level
    uint32_t block_chunks;    // block edge in chunks
    uint32_t first_chunk[3];  // chunk coordinates (full grid) of block (0,0,0)'s first chunk
    const float* explicit_origins;  // non-null: n_blocks x 3 chunk origins in root space, parent = 0
    // parent level
    uint32_t ratio;           // parent block edge / this block edge
    uint32_t parent_nb[3];    // 0,0,0 → single parent (the root list)
    const Instr* parent_instrs;
    const uint32_t* parent_off;
    const uint32_t* parent_len;
    // outputs
    Instr* out_instrs;
    const uint32_t* out_off;
    uint32_t* out_len;
    DevChunk* chunks;         // exact level only (may be null for explicit origins)
    uint32_t* max_depth;      // exact level only
    uint32_t* occ;            // exact level only: min xyz, max xyz of non-empty voxels
    uint32_t* error_flag;
    uint32_t prune;           // 1: drop operands that provably cannot influence any voxel of the block
    uint32_t saturate;        // 1: replace a program whose value range quantises to one code by a constant
    uint32_t own_lo, own_hi;  // exact level: local chunk planes [own_lo, own_hi) are generated, the rest are halo planes
};
cudaError_t launch_fold(bool exact, const FoldArgs& a, cudaStream_t st);

struct EvalArgs {
    const ivx_node* nodes;
    GenParams gp;
    uint32_t n_active;
    const uint32_t* active;   // compacted chunk indices (null → identity)
    uint32_t nb[3];
    uint32_t first_i;
    const float* explicit_origins;
    const Instr* instrs;
    const uint32_t* off;
    const uint32_t* len;
    const uint32_t* slot_of;  // per chunk
    unsigned char* voxels;
    DevChunk* chunks;
    uint32_t* occ;
    float* raw_out;           // non-null: write f32 distances instead of voxels
    int saturate_final_noise; // 1: skip a final noise term where it cannot change the stored code
    int smem_levels;
    float* spill;
    int spill_levels;
    float neg_zero;           // -0.0f, deliberately a run-time value (packed.cuh)
};
cudaError_t launch_eval(const EvalArgs& a, uint32_t grid, cudaStream_t st);
int eval_max_blocks_per_sm(int smem_levels);
cudaError_t launch_eval_blocks(const ivx_node* nodes, uint32_t n_nodes, const float* origins, uint32_t n_blocks, int size,
                               float* out, cudaStream_t st);

// ---- types.cu ----------------------------------------------------------------
struct TypesArgs {
    GenParams gp;
    uint32_t n_active;
    const uint32_t* active;   // compacted chunk indices (null → identity)
    uint32_t nb[3];
    uint32_t first_i;
    const uint32_t* slot_of;  // per chunk
    unsigned char* voxels;
    DevChunk* chunks;
    uint32_t* occ;
    float neg_zero;           // -0.0f, deliberately a run-time value (see simplex4_tab2 in types.cu)
    unsigned long long* noise_evaluations;  // profiling: 4-D simplex evaluations performed (null: not counted)
};
cudaError_t launch_types(const TypesArgs& a, uint32_t grid, cudaStream_t st);
int types_max_blocks_per_sm();

// ---- scan.cu ----------------------------------------------------------------
// Exclusive prefix sums over small arrays (chunk-count sized), single CTA.
cudaError_t launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* total, cudaStream_t st);

cudaError_t launch_child_caps(const uint32_t* parent_len, uint32_t n_blocks, const uint32_t nb[3],
                              const uint32_t parent_nb[3], uint32_t ratio, uint32_t* caps, cudaStream_t st);
cudaError_t launch_plan_slots(const DevChunk* chunks, uint32_t n, const uint32_t nb[3], uint32_t* active_flag,
                              uint32_t* slot_flag, cudaStream_t st);
cudaError_t launch_scatter_active(const uint32_t* active_flag, const uint32_t* active_scan, uint32_t n,
                                  uint32_t* active_list, cudaStream_t st);
cudaError_t launch_fill_u32(uint32_t* p, uint32_t n, uint32_t v, cudaStream_t st);

// ---- derive.cu ---------------------------------------------------------------
struct AbsorbRange;
// a box of chunks: first chunk and extent per axis
struct ChunkBox {
    uint32_t c0[3], d[3];
};
// The boundary refresh after a modification, over the chunks of `box` only (every chunk with a face in a refreshed
// pair lies in it): `prep` decides conversions and their slots (one CTA; *total = slots handed out, numbered from
// first_slot + *first_extra in chunk order), `apply` rewrites adjacency bits and obscuredness. face_mask /
// convert_flag / slot_of hold one entry per chunk of the box; a box of more than BOUNDARY_BOX_ONE_CTA chunks is prepared
// by a grid and the library's prefix sum instead and needs `need` / `ord` (one word per chunk of the box each).
constexpr uint32_t BOUNDARY_BOX_ONE_CTA = 8192;
cudaError_t launch_boundary_refresh_box(DevChunk* chunks, uint32_t n, const uint32_t nb[3], const ChunkBox& box,
                                        const AbsorbRange& range, uint32_t first_slot, const uint32_t* first_extra,
                                        uint8_t* face_mask, uint32_t* convert_flag, uint32_t* slot_of, uint8_t* label_stale,
                                        uint32_t* total, uint32_t* need, uint32_t* ord, bool prep, bool apply,
                                        unsigned char* voxels, uint32_t grid, cudaStream_t st);
// own_lo / own_hi: only chunks of local planes [own_lo, own_hi) are classified / updated (halo planes belong to
// the neighbour rank)
cudaError_t launch_boundary_classify(const DevChunk* chunks, uint32_t n, const uint32_t nb[3], const uint8_t* face_mask,
                                     uint32_t* convert_flag, uint32_t own_lo, uint32_t own_hi, cudaStream_t st);
cudaError_t launch_boundary_apply(DevChunk* chunks, uint32_t n, const uint32_t nb[3], const uint8_t* face_mask,
                                  const uint32_t* convert_flag, const uint32_t* slot_of, unsigned char* voxels,
                                  const uint32_t* work_list, uint32_t n_work, uint32_t own_lo, uint32_t own_hi,
                                  uint32_t grid, cudaStream_t st);

// ---- extract.cu --------------------------------------------------------------
struct ExtractArgs {
    // the object the region leaves
    DevChunk* src_chunks;
    unsigned char* src_voxels;
    const uint8_t* src_labels;   // LocalRegionLabel per voxel, per slot (split.cu)
    uint8_t* src_dirty;
    uint8_t* src_label_stale;    // may be null
    // per chunk of the extracted object's chunk grid
    uint32_t n_ext;
    const uint8_t* mode;         // 0 padding, 1 Uniform, 2 NonUniform moved whole, 3 NonUniform shared with other regions
    const uint32_t* src_index;   // linear chunk index in the source object
    const uint32_t* first_region;  // index of the source chunk's region 0 in region_is_r
    const uint32_t* dst_slot;    // slot of NonUniform chunks in the extracted object
    const uint8_t* region_is_r;  // per local region of the source: 1 = part of the extracted global region
    DevChunk* dst_chunks;
    unsigned char* dst_voxels;
    uint32_t* non_empty_count;   // non-empty voxels that moved
};
cudaError_t launch_extract_chunks(const ExtractArgs& a, uint32_t grid, cudaStream_t st);
cudaError_t launch_ingest_chunks(const unsigned char* src, const uint8_t* sparseness, uint32_t n, DevChunk* chunks,
                                 unsigned char* voxels, uint32_t* slot_counter, uint32_t grid, cudaStream_t st);
cudaError_t launch_repack_single(const DevChunk* chunks, const uint32_t nb[3], const unsigned char* voxels, const uint32_t org[3],
                                 const uint32_t occ_lo[3], const uint32_t occ_hi[3], DevChunk* out_chunk, unsigned char* out_slot,
                                 cudaStream_t st);

// ---- halo.cu -----------------------------------------------------------------
size_t halo_message_bytes(uint32_t plane_chunks);
cudaError_t launch_halo_pack(const DevChunk* chunks, uint32_t plane_first, uint32_t plane_chunks, uint32_t layer_i,
                             const unsigned char* voxels, unsigned char* dst, cudaStream_t st);
cudaError_t launch_halo_unpack(DevChunk* chunks, uint32_t plane_first, uint32_t plane_chunks, uint32_t first_slot,
                               uint32_t layer_i, unsigned char* voxels, const unsigned char* src, cudaStream_t st);
cudaError_t launch_halo_kinds_pack(const DevChunk* chunks, const uint32_t* convert_flag, uint32_t plane_first,
                                   uint32_t plane_chunks, uint8_t* dst, cudaStream_t st);
cudaError_t launch_halo_kinds_unpack(DevChunk* chunks, uint32_t plane_first, uint32_t plane_chunks, const uint8_t* src,
                                     cudaStream_t st);
cudaError_t launch_push_words(const void* src, void* dst, size_t n_words, uint32_t add, uint32_t period, uint32_t col,
                              uint32_t grid, cudaStream_t st);

// ---- mesh.cu -----------------------------------------------------------------
struct MeshArgs {
    const DevChunk* chunks;
    const unsigned char* voxels;
    uint32_t nb[3];
    uint32_t first_i;
    float voxel_extent;
    const uint32_t* work;  // chunk indices to mesh, ascending linear order
    uint32_t n_work;
    // counting pass outputs / emit pass inputs (per work item)
    uint32_t* vertex_count;
    uint32_t* index_count;
    uint32_t* has_submesh;
    const uint32_t* vertex_offset;
    const uint32_t* index_offset;
    const uint32_t* submesh_ord;
    // emit outputs
    float* positions;
    float* normals;
    uint32_t* indices;
    ivx_index_materials* index_materials;
    ivx_chunk_submesh* submeshes;
    uint32_t* vertex_ranges;
    // quads whose four vertices do not share one material are recorded here by the emit pass and finished by
    // k_mesh_materials (one thread per quad) instead of by a few lanes of the emit kernel
    uint4* mq_entries;      // 3 x uint4 per quad
    uint32_t* mq_count;
    uint32_t mq_capacity;   // quads; beyond it the emit kernel finishes the quad in place
    // capacity of the output buffers when they were sized from a plan instead of from this call's own counts
    // (0 = sized exactly): a chunk that would not fit is skipped, the plan check reports the mismatch
    uint32_t cap_vertices, cap_indices, cap_submeshes;
};
cudaError_t launch_mesh_materials(const uint4* entries, const uint32_t* count, uint32_t capacity,
                                  ivx_index_materials* index_materials, cudaStream_t st);
cudaError_t launch_exposed_flags(const DevChunk* chunks, uint32_t n, const uint32_t nb[3], uint32_t own_lo,
                                 uint32_t own_hi, uint32_t* flag, cudaStream_t st);
cudaError_t launch_mesh(bool emit, const MeshArgs& a, uint32_t grid, cudaStream_t st);

// ---- modify.cu ---------------------------------------------------------------
struct AbsorbRange {
    uint32_t c0[3], c1[3];  // chunk range
    uint32_t v0[3], v1[3];  // voxel range
};
// the absorbing shape in normalized voxel space: a sphere, or a capsule (segment start, segment vector)
// mutual absorption between two objects (interaction/absorption.rs:891-1080): the object being modified samples the
// other one's signed distances — B's voxels directly while A is modified (mode 2), A's pre-modification snapshot while B
// is modified (mode 3)
struct MutualArgs {
    float q[4], t[3];          // transform_from_b_to_a: unit quaternion (x, y, z, w), translation
    float extent;              // voxel extent of the object being modified
    float inv_extent_other;    // inverse voxel extent of the other object
    float dist_scale;          // mode 2: b_dist_to_a = extent_b / extent_a; mode 3: a_dist_to_b
    float smoothness, qik;     // Smoothness { smoothness, 0.25 / smoothness }
    const DevChunk* o_chunks;  // mode 2: object B
    const unsigned char* o_voxels;
    uint32_t o_nb[3];
    float* snapshot;           // A's signed distances over the padded intersection ranges [s0, s1)
    uint32_t s0[3], s1[3];
};
struct AbsorbShape {
    int capsule;               // 0 sphere, 1 capsule, 2 / 3 voxel ranges of a mutual absorption (object A / object B)
    float center[3];           // sphere centre / capsule segment start
    float seg[3];              // capsule segment vector
    float seg_over_len2[3];    // CapsulePointContainmentTester (capsule.rs:168-181)
    float radius;              // absorbing radius
    float influence_radius;    // radius of the shape whose voxels are visited
    float influence_radius_sq;
};
struct AbsorbArgs {
    DevChunk* chunks;
    uint3 nb;
    AbsorbRange range;
    uint32_t n_range;
    unsigned char* voxels;
    AbsorbShape shape;
    uint32_t first_new_slot;
    const uint32_t* new_slot_ord;
    uint8_t* dirty;
    uint8_t* label_stale;  // may be null: set for chunks whose region labels no longer hold (a voxel was emptied, or the
                           // chunk was Uniform and has none yet)
    uint32_t* stats;  // touched chunks, touched voxels, emptied voxels, removed chunks
    // inertial-property update (null unless asked for): which voxels went from non-empty to empty, per chunk of the
    // range in its visiting order — 256 columns x 16 bits — and {count, voxel slot} per chunk
    uint16_t* removed_cols;
    uint32_t* removed_info;
    MutualArgs mutual;  // shape.capsule >= 2
};
cudaError_t launch_absorb_plan(const DevChunk* chunks, const uint32_t nb[3], const AbsorbRange& r, const AbsorbShape& shape,
                               uint32_t* need_slot, uint32_t n_range, cudaStream_t st);
cudaError_t launch_absorb_apply(const AbsorbArgs& a, uint32_t grid, cudaStream_t st);
cudaError_t launch_absorb_face_mask(const uint32_t nb[3], const AbsorbRange& b, uint8_t* face_mask, uint32_t n,
                                    cudaStream_t st);
// also marks the converted chunks' region labels stale (label_stale may be null)
cudaError_t launch_need_slot_for_convert(const DevChunk* chunks, const uint32_t* convert_flag, uint32_t n, uint32_t* need,
                                         uint8_t* label_stale, cudaStream_t st);
cudaError_t launch_count_nonzero_u8(const uint8_t* a, uint32_t n, uint32_t* out, cudaStream_t st);
cudaError_t launch_assign_slots(const DevChunk* chunks, const uint32_t* need, const uint32_t* ord, uint32_t first,
                                const uint32_t* first_extra,
                                uint32_t n, uint32_t* slot_of, cudaStream_t st);
// `gate` (may be null): a device word; when it is zero the kernels return at once and `occ` keeps its initial value
cudaError_t launch_occupied_ranges(const DevChunk* chunks, uint32_t n, const uint32_t nb[3], uint32_t first_i,
                                   const unsigned char* voxels, uint32_t* occ, uint32_t* chunk_minmax_scratch,
                                   const uint32_t* gate, cudaStream_t st);

// ---- split.cu ------------------------------------------------------------------
// sets arr[chunk] = value for the chunks of the box [lo, hi) (clamped to the grid); no-op when arr is null
cudaError_t launch_mark_box(uint8_t* arr, const uint32_t nb[3], const uint32_t lo[3], const uint32_t hi[3], uint8_t value,
                            cudaStream_t st);

// ---- regions.cu: the global pass of connected-region detection ------------------
// words of the result block (device counters, read back once per resolve)
enum RegionWord : uint32_t {
    RW_WORK = 0,          // chunks re-labelled
    RW_RECORDS = 1,       // connection records found (may exceed the capacity: retry)
    RW_LABEL_ERROR = 2,   // k_local_regions / k_region_connections limits
    RW_TOTAL = 3,         // local regions of the object
    RW_TREES = 4,
    RW_EVENTS = 5,
    RW_ERROR = 6,         // RegionError
    RW_ROOTS = 7,         // connected regions
    RW_FIRST_ROOT = 8,    // region index of the first / second root in linear order
    RW_SECOND_ROOT = 9,
    RW_FIRST_LABEL = 10,  // the same as GlobalRegionLabel (chunk << 8 | region)
    RW_SECOND_LABEL = 11,
    RW_ERROR_INFO = 12,   // 4 words
    RW_CANDIDATES = 16,   // 2 x { chunks, NonUniform chunks, chunk min[3], chunk max[3] }
    RW_COUNT = 32
};
enum RegionError : uint32_t { RERR_REGION_CAPACITY = 1, RERR_TOO_MANY_CONNECTIONS = 2, RERR_PAIR_TABLE_FULL = 3 };

struct RegionPass {
    uint32_t n;                  // chunks
    uint32_t stride0, stride1;   // linear chunk index strides of dimension 0 and 1
    const uint32_t* regions;     // per chunk: kind << 16 | boundary_region_count << 8 | region_count
    uint32_t* words;             // RegionWord
    uint32_t cap;                // capacity of the per-region arrays
    uint32_t* counts;            // per chunk (scratch)
    uint32_t* first;             // per chunk: index of its region 0
    uint32_t* label;             // per region: chunk << 8 | region
    uint32_t* root;              // per region: region index of its root
    uint32_t* lowest;            // per region: lowest adjacent region below it, else itself
    uint32_t* tree;              // per region: fixed point of `lowest`
    uint32_t* degree;            // zeroed by the caller
    uint32_t* fresh_flag;        // zeroed by the caller
    uint32_t* tree_number;
    uint32_t* tree_vertex;       // per tree: its fresh region
    uint32_t* tree_root;         // per tree: region index of the root of its set
    uint32_t* tree_parent;       // replay forest when the trees do not fit shared memory
    uint32_t* event_count;       // per region (as visiting time); zeroed by the caller
    uint32_t* event_offset;
    uint32_t* event_cursor;      // zeroed by the caller
    const uint2* records;        // k_region_connections output
    uint32_t record_cap;
    uint2* edges;                // per record: (lower region, upper region)
    uint2* events;               // per tree pair, in time order: (absorbing tree, absorbed tree)
    unsigned long long* slot_keys;  // pair table, all ones = empty; slot_values all ones
    uint32_t* slot_values;
    uint32_t slot_mask;
};
cudaError_t launch_region_result_init(uint32_t* words, cudaStream_t st);
cudaError_t launch_region_global_pass(const RegionPass& p, uint32_t n_scan, uint32_t* launches, int max_shared_bytes,
                                      cudaStream_t st);
cudaError_t launch_region_root_labels(const uint32_t* root, const uint32_t* label, uint32_t total, uint32_t* out, cudaStream_t st);

struct ExtractPlanArgs {
    const uint32_t* regions;
    const uint32_t* first;
    const uint32_t* root;
    uint32_t total;          // local regions of the source object
    uint32_t region_root;    // region index of the extracted region's root
    uint32_t n_ext, ext[3], lo[3], nb1, nb2;
    uint8_t* mode;
    uint32_t* src_index;
    uint32_t* first_region;
    uint32_t* non_uniform_flag;
    uint32_t* dst_slot;
    uint32_t* uniform_count;
    uint8_t* is_member;      // per local region of the source
};
cudaError_t launch_extract_plan(const ExtractPlanArgs& a, uint32_t* slot_total, cudaStream_t st);

// ---- api.cu helpers ------------------------------------------------------------
cudaError_t launch_flag_dirty_exposed(const DevChunk* chunks, const uint8_t* dirty, uint32_t n, uint32_t* exposed_flag,
                                      uint32_t* dirty_flag, cudaStream_t st);
cudaError_t launch_pack_voxels(const DevChunk* chunks, uint32_t n, const uint32_t* ordinal, const uint32_t* part_counts,
                               uint32_t part, const unsigned char* voxels, ivx_voxel* out, ivx_chunk_desc* out_chunks,
                               uint32_t grid, cudaStream_t st);
cudaError_t launch_nonuniform_flags(const DevChunk* chunks, uint32_t n, uint32_t* flag, cudaStream_t st);
cudaError_t launch_set_reserved_slots(DevChunk* chunks, uint32_t n, const uint32_t* slot_flag, const uint32_t* slot_scan,
                                      uint32_t* slot_of, cudaStream_t st);

}  // namespace ivx
