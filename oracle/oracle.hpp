// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// CPU restatement of the reference's `impact_voxel` hot path:
//   atomic SDF graph → generator (atomic.rs:228-596)
//   block evaluator with culling (atomic.rs:633-875, 1601-1848)
//   voxel generation / classification (generation.rs:207-371, voxel_type.rs)
//   chunked object + derived state (object.rs)
//   18³ brick fill + Surface Nets + mesh assembly (object/sdf.rs,
//   object/sdf/surface_nets.rs, mesh.rs)
//   sphere absorption + dirty-chunk bookkeeping (object/intersection.rs,
//   interaction/absorption.rs)
//   inertial moments of an object and their incremental update (object/inertia.rs)
// Paths are relative to /root/reference/engine/crates/impact_voxel/src/.
//
// PARITY PINNING: the reference is Rust and cannot be built or run here (no
// cargo/rustc, un-vendored deps). Pinned against the reference's own golden
// vectors: surface_nets.rs:676-877 (vertex/index materials) and the structural
// tests in object.rs:3563-4071 / intersection.rs:1093-1348 restated in
// tests/. The NOISE functions restate simdnoise 3.1.7 (fork e59c958e) from its
// published algorithm; no reference test pins them → noise parity UNPINNED.
#pragma once
#include <cstdint>
#include <string>
#include <map>
#include <unordered_map>
#include <vector>

#include "oracle_math.hpp"

namespace orc {

constexpr int CHUNK_SIZE = 16;
constexpr int CHUNK_VOXELS = 4096;
constexpr int BRICK_SIZE = 18;
constexpr int BRICK_CELLS = 5832;

// --- voxel encoding (lib.rs:60-101, 154-269) -------------------------------
constexpr float QUANT_STEP = 0.02f;
constexpr float INV_QUANT_STEP = 1.0f / 0.02f;  // == 50.0f in f32
constexpr float SD_MAX_F32 = 0.02f * 127.0f;
constexpr float SD_MIN_F32 = 0.02f * -128.0f;
constexpr int8_t VOID_LIMIT = 100;

constexpr uint8_t FLAG_EMPTY = 1 << 0;
constexpr uint8_t FLAG_ADJ_X_DN = 1 << 2;
constexpr uint8_t FLAG_ADJ_Y_DN = 1 << 3;
constexpr uint8_t FLAG_ADJ_Z_DN = 1 << 4;
constexpr uint8_t FLAG_ADJ_X_UP = 1 << 5;
constexpr uint8_t FLAG_ADJ_Y_UP = 1 << 6;
constexpr uint8_t FLAG_ADJ_Z_UP = 1 << 7;
constexpr uint8_t FLAG_FULL_ADJ = 0xFC;
constexpr uint8_t TYPE_DUMMY = 255;

struct Voxel {
    uint8_t type;
    int8_t sd;
    uint8_t flags;
};
static_assert(sizeof(Voxel) == 3, "Voxel is 3 bytes repr(C)");

// `(value * 50.0) as i8`: saturating, truncating, NaN → 0 (lib.rs:195-201)
static inline int8_t sd_encode(float v) {
    float s = v * INV_QUANT_STEP;
    if (s != s) return 0;
    if (s >= 127.0f) return 127;
    if (s <= -128.0f) return -128;
    return (int8_t)(int)s;
}
static inline float sd_decode(int8_t e) { return (float)e * QUANT_STEP; }

// --- atomic SDF graph (atomic.rs:55-181) ------------------------------------
enum NodeKind : uint32_t {
    K_SPHERE = 0,
    K_CAPSULE = 1,
    K_BOX = 2,
    K_TRANSLATION = 3,
    K_ROTATION = 4,
    K_SCALING = 5,
    K_NOISE = 6,
    K_UNION = 7,
    K_SUBTRACTION = 8,
    K_INTERSECTION = 9,
};

// Input node, POD (same layout as include/impact_voxel_cuda.h ivx_sdf_node).
//   sphere:  p[0]=radius
//   capsule: p[0]=segment_length p[1]=radius
//   box:     p[0..3]=extents
//   translation: child[0], p[0..3]
//   rotation:    child[0], p[0..4] = quaternion x,y,z,w
//   scaling:     child[0], p[0]
//   noise:       child[0], octaves, seed, p[0]=frequency p[1]=lacunarity
//                p[2]=persistence p[3]=amplitude
//   union/sub/inter: child[0], child[1], p[0]=smoothness
struct SdfNode {
    uint32_t kind;
    uint32_t child[2];
    uint32_t octaves;
    uint32_t seed;
    float p[8];
};

// ProcessedSDFNode mirror (atomic.rs:83-102).
//   leaf params: sphere p[0]=radius; capsule p[0]=half_segment_length p[1]=radius;
//   box p[0..3]=half_extents; scaling p[0]; noise p[0]=frequency p[1]=lacunarity
//   p[2]=persistence p[3]=amplitude p[4]=noise_scale; combine p[0]=smoothness
//   p[1]=0.25/smoothness
struct ProgNode {
    uint32_t kind;
    uint32_t octaves;
    uint32_t seed;
    uint32_t leaf_count;
    float p[8];
    float transform[16];  // column-major root→node
    float dom_lo[3];
    float dom_hi[3];
    float margin;
    uint32_t _pad;
};

struct Generator {
    std::vector<ProgNode> nodes;  // post-order, DAG unrolled
    uint32_t stack_size = 0;      // required_forward_stack_size
    Aabb domain{{0, 0, 0}, {0, 0, 0}};
    bool empty() const { return nodes.empty(); }
};

// Returns "" on success, else the error string (cycle / missing node).
std::string build_generator(const SdfNode* nodes, uint32_t n, uint32_t root, Generator& out);

// compute_signed_distances_for_block<16,4096> (atomic.rs:633-875)
// `stack` must hold (stack_size+1)*4096 floats; result in stack[0..4096].
// If `decisions` != nullptr it receives one byte per program node:
//   leaf: 0 = evaluated, 1 = filled +margin, 2 = filled -margin
//   noise/combine: 0 = applied, 1 = skipped; others 0.
void eval_chunk(const Generator& g, V3 chunk_lo_root, float* stack, uint8_t* decisions);
// compute_signed_distances_for_block_preserving_gradients<SIZE,COUNT> (atomic.rs:877-998)
void eval_block_preserving_gradients(const Generator& g, V3 block_origin_root, int size,
                                     float* stack);

// --- noise (simdnoise 3.1.7 restatement; UNPINNED) ---------------------------
float simplex3(float x, float y, float z, int32_t seed);
float fbm3(float x, float y, float z, float lacunarity, float gain, uint32_t octaves,
           int32_t seed);
float simplex4(float x, float y, float z, float w, int32_t seed);

// --- voxel generator (generation.rs:70-371, voxel_type.rs) -------------------
struct TypeGen {
    uint32_t kind = 0;  // 0 = Same, 1 = GradientNoise
    uint8_t same_type = 0;
    uint32_t n_types = 1;
    float noise_frequency = 0;
    float voxel_type_frequency = 0;
    uint32_t seed = 0;
};

struct VoxelGenerator {
    float voxel_extent = 1.0f;
    uint32_t grid_shape[3] = {0, 0, 0};
    V3 shifted_center{-0.5f, -0.5f, -0.5f};
    Generator sdf;
    TypeGen types;
};
void make_voxel_generator(VoxelGenerator& vg, float voxel_extent);  // SDFVoxelGenerator::new

struct Sparseness {
    bool only_empty, is_void;
};
// generate_chunk (generation.rs:293-371); scratch = (stack_size+1)*4096 floats,
// type_scratch = 4096*n_types floats.
Sparseness generate_chunk(const VoxelGenerator& vg, const uint32_t origin[3], Voxel* voxels,
                          float* scratch, float* type_scratch);

// --- chunked object (object.rs) ----------------------------------------------
enum ChunkKind : uint8_t { CK_VOID = 0, CK_UNIFORM = 1, CK_NONUNIFORM = 2 };
enum FaceDist : uint8_t { FD_EMPTY = 0, FD_FULL = 1, FD_MIXED = 2 };
constexpr uint8_t CF_OBSCURED_ALL = 0x3F;
constexpr uint8_t CF_ONLY_EMPTY = 1 << 6;

struct Chunk {
    uint8_t kind = CK_VOID;
    Voxel uniform_voxel{0, 0, 0};
    uint32_t data_offset = 0;
    uint8_t face[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    uint8_t flags = 0;
};

struct Object {
    float voxel_extent = 1.0f;
    uint32_t chunk_counts[3] = {0, 0, 0};
    uint32_t occ_chunks[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    uint32_t occ_voxels[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    std::vector<Chunk> chunks;
    std::vector<Voxel> voxels;
    std::vector<uint32_t> dirty;  // linear chunk indices, insertion order, unique

    uint32_t lin(uint32_t i, uint32_t j, uint32_t k) const {
        return (i * chunk_counts[1] + j) * chunk_counts[2] + k;
    }
    Chunk get_chunk(int64_t i, int64_t j, int64_t k) const {
        if (i < 0 || j < 0 || k < 0 || i >= chunk_counts[0] || j >= chunk_counts[1] ||
            k >= chunk_counts[2])
            return Chunk{};
        return chunks[lin((uint32_t)i, (uint32_t)j, (uint32_t)k)];
    }
    Voxel* chunk_voxels(uint32_t data_offset) { return voxels.data() + (size_t)data_offset * 4096; }
    const Voxel* chunk_voxels(uint32_t data_offset) const {
        return voxels.data() + (size_t)data_offset * 4096;
    }
    void mark_dirty(uint32_t idx);
};

// VoxelObject::generate (object.rs:239-263): generate_without_derived_state,
// update_occupied_voxel_ranges, compute_all_derived_state (minus split detection).
// n_threads > 1 splits the linear chunk index into contiguous ranges
// (object.rs:423-427); result is identical.
void generate_object(const VoxelGenerator& vg, Object& obj, int n_threads,
                     double* t_generate_s = nullptr, double* t_derive_s = nullptr);
void generate_without_derived_state(const VoxelGenerator& vg, Object& obj, int n_threads);
// Same, restricted to chunk planes [plane_begin, plane_end) of the x-major chunk grid (a bounded
// sample of a large object for bench.py's cpu_baseline; the slab is treated as its own object).
void generate_slab_without_derived_state(const VoxelGenerator& vg, Object& obj, int n_threads,
                                         uint32_t plane_begin, uint32_t plane_end);
void update_occupied_voxel_ranges(Object& obj);
// generate_without_derived_state for a generator whose chunks are given as data (sparseness: bit 0 has_only_empty_voxels,
// bit 1 is_void per chunk)
void object_from_generated_chunks(const Voxel* voxels, const uint8_t* sparseness, const uint32_t grid_shape[3],
                                  float voxel_extent, Object& obj);
void update_occupied_chunk_ranges(Object& obj);
void compute_all_derived_state(Object& obj);
void update_internal_adjacencies(Voxel* chunk_voxels);
void update_upper_boundary_adjacencies_in_ranges(Object& obj, const uint32_t r[3][2]);
Sparseness update_all_internal_state(Chunk& c, Voxel* chunk_voxels);

// --- brick + surface nets + mesh (object/sdf.rs, surface_nets.rs, mesh.rs) ---
struct Brick {
    float values[BRICK_CELLS];
    uint8_t types[BRICK_CELLS];
    bool adj_non_uniform[3][2];
};
// fill_sdf_for_chunk_if_exposed (object/sdf.rs:181-213); returns false if not exposed.
bool fill_brick_if_exposed(const Object& obj, uint32_t ci, uint32_t cj, uint32_t ck, Brick& b,
                           uint8_t* chunk_flags);

struct VertexMaterials {
    uint8_t indices[8];
    uint8_t weights[8];
};
struct IndexMaterials {
    uint8_t indices[4];
    uint8_t weights[4];
};
VertexMaterials vertex_materials_compute(const bool has_voxel[8], const uint8_t mat[8]);
void index_materials_for_triangle(const VertexMaterials* vm[3], IndexMaterials out[3]);

struct ChunkMesh {
    std::vector<float> positions;  // 3 per vertex
    std::vector<float> normals;    // 3 per vertex
    std::vector<VertexMaterials> vertex_materials;
    std::vector<IndexMaterials> index_materials;
    std::vector<uint16_t> indices;
    std::vector<uint16_t> surface_lin;  // brick linear idx per vertex
    std::vector<uint16_t> lin_to_vertex;
};
// compute_surface_nets_mesh (surface_nets.rs:131-148)
void surface_nets(const Brick& b, float voxel_extent, V3 position_offset, ChunkMesh& out);

struct Submesh {
    uint32_t chunk_indices[3];
    uint32_t index_offset;
    uint32_t index_count;
    uint32_t obscured[2][2][2];
};
struct Mesh {
    std::vector<float> positions;
    std::vector<float> normals;
    std::vector<IndexMaterials> index_materials;
    std::vector<uint32_t> indices;
    std::vector<Submesh> submeshes;
    std::vector<uint32_t> vertex_ranges;  // 2 per submesh: start, end
};
// VoxelObjectMesh::recreate (mesh.rs:286-354)
void mesh_object(const Object& obj, Mesh& mesh, int n_threads = 1);
// One chunk of VoxelObjectMesh::sync_with_voxel_object (mesh.rs:360-456): returns
// false when the chunk is not exposed or yields an empty mesh (→ submesh removed).
bool mesh_chunk(const Object& obj, uint32_t ci, uint32_t cj, uint32_t ck, ChunkMesh& cm,
                uint8_t* chunk_flags);

// --- the mesh kept in sync with a modified object (mesh.rs:360-456, 749-848) ---
// RangeAllocator (impact_containers/src/range_allocator.rs): free ranges ordered by start
struct RangeAllocator {
    std::map<size_t, size_t> free_ranges;  // start → end
    void free_range(size_t start, size_t end);
    bool allocate_range(size_t required_len, size_t& start);  // smallest fitting range, the first of equals
    void merge_consecutive_ranges();
};
// VoxelObjectMesh + ChunkSubmeshManager: `mesh.submeshes` / `mesh.vertex_ranges` are the manager's tables in its own
// order (push order, swap-remove on removal); the buffers keep obsolete data in freed ranges.
struct SyncedMesh {
    Mesh mesh;
    std::unordered_map<uint32_t, uint32_t> index_of_chunk;  // KeyIndexMapper: linear chunk index → table index
    std::vector<uint32_t> chunk_at_index;
    RangeAllocator free_vertices, free_indices;
    std::vector<uint32_t> updated;  // ChunkSubmeshDataRanges since the last report: vertex start, end, index start, end
    bool chunks_were_removed = false;
};
void synced_mesh_create(const Object& obj, SyncedMesh& sm, int n_threads = 1);
// sync_with_voxel_object over `dirty` (linear chunk indices) in the given order — the reference iterates a HashSet
void synced_mesh_sync(const Object& obj, SyncedMesh& sm, const uint32_t* dirty, size_t n_dirty);

// --- collision probes (collidable.rs:97-101, 346-780) --------------------------
struct CollisionProbes {
    std::vector<float> points;  // xyz; freed ranges keep obsolete points
    std::unordered_map<uint32_t, std::pair<size_t, size_t>> range_of_chunk;  // linear chunk index → point range
    RangeAllocator free_points;
};
uint32_t probes_log2_block_size(const Object& obj);
void probes_points_for_chunk(uint32_t log2_block_size, const uint32_t chunk_indices[3], const float* positions, const float* normals,
                             size_t n_vertices, const uint32_t* indices, size_t n_indices, uint32_t start_index,
                             float inverse_voxel_extent, std::vector<float>& points);
void probes_compute_for_all_chunks(const Object& obj, const Mesh& mesh, CollisionProbes& pr);
void probes_sync(const Object& obj, const SyncedMesh& sm, const uint32_t* dirty, size_t n_dirty, CollisionProbes& pr);
struct VoxelContact;
struct Isometry;
struct InertialMoments;
void mutual_voxel_object_contacts(const Object& a, const CollisionProbes& probes_a, const InertialMoments& inertial_a,
                                  const Isometry& world_to_a, const Object& b, const CollisionProbes& probes_b,
                                  const InertialMoments& inertial_b, const Isometry& world_to_b, const uint32_t ranges_in_a[3][2],
                                  const uint32_t ranges_in_b[3][2], std::vector<VoxelContact>& a_against_b,
                                  std::vector<VoxelContact>& b_against_a);

// --- inertial properties (object/inertia.rs) ----------------------------------
// VoxelObjectInertialPropertyManager (inertia.rs:19-25): mass, moments (m x), moments of inertia, products of inertia,
// all with respect to the origin of the voxel grid.
struct InertialMoments {
    float mass = 0.0f;
    float moments[3] = {0, 0, 0};
    float moi[3] = {0, 0, 0};
    float poi[3] = {0, 0, 0};
};
void moments_for_voxel(float e, float e2, float e3, const float* densities, const uint32_t ijk[3], uint8_t type,
                       InertialMoments& out);
void moments_for_non_uniform_chunk(float e, const Voxel* voxels, const float* densities, const uint32_t cc[3],
                                   InertialMoments& out);
void moments_for_uniform_chunk(float e, const float* densities, uint8_t type, const uint32_t cc[3],
                               InertialMoments& out);
// initialized_from (inertia.rs:125-137); per_chunk (optional, one per chunk of the grid) receives the chunk terms
void inertial_moments_for_object(const Object& obj, const float* densities, InertialMoments& out,
                                 InertialMoments* per_chunk);
// VoxelObjectInertialPropertyUpdater (inertia.rs:27-35, 174-189, 374-395)
struct InertialUpdater {
    InertialMoments* parent;
    const float* densities;
    float e, e2, e3;
    uint64_t removed = 0;
    InertialUpdater(InertialMoments* p, float voxel_extent, const float* dens)
        : parent(p), densities(dens), e(voxel_extent), e2(voxel_extent * voxel_extent), e3(e2 * voxel_extent) {}
    void remove_voxel(const uint32_t ijk[3], uint8_t type);
};

// --- modification (object/intersection.rs, interaction/absorption.rs) --------
struct AbsorbStats {
    uint32_t touched_chunks;
    uint32_t touched_voxels;
    uint32_t emptied_voxels;
    uint32_t removed_chunks;
};
// apply_sphere_absorption restricted to the voxel-object side: influence sphere
// (centre, influence_radius) and absorbing radius, all in normalized voxel space.
void absorb_sphere(Object& obj, V3 center, float radius, float influence_radius,
                   AbsorbStats* stats, InertialUpdater* updater = nullptr);
// apply_capsule_absorption likewise: influence capsule (segment start, segment vector, influence_radius)
void absorb_capsule(Object& obj, V3 segment_start, V3 segment_vector, float radius, float influence_radius,
                    AbsorbStats* stats, InertialUpdater* updater = nullptr);

// for_each_surface_voxel_in_voxel_ranges (object/intersection.rs:97-151) as a list in the closure's call order;
// placement: VoxelSurfacePlacement 0 Face, 1 Edge, 2 Corner (lib.rs:109-114, 330-343)
struct SurfaceVoxel {
    uint32_t ijk[3];
    Voxel voxel;
    uint8_t placement;
};
static_assert(sizeof(SurfaceVoxel) == 16, "16-byte records");
void surface_voxels_in_ranges(const Object& obj, const uint32_t ranges[3][2], std::vector<SurfaceVoxel>& out);

// for_each_sphere_voxel_object_contact (collidable.rs:1097-1127) + determine_sphere_sphere_contact_geometry
// (impact_physics/src/collision/collidable/sphere.rs:105-136) as a list in the closure's call order
struct VoxelContact {
    uint32_t ijk[3];
    float position[3], normal[3], penetration_depth;
};
static_assert(sizeof(VoxelContact) == 40, "ten words");
struct Isometry;
void sphere_voxel_object_contacts(const Object& obj, const Isometry& transform_to_object_space, V3 center, float radius,
                                  std::vector<VoxelContact>& out);

// for_each_voxel_object_plane_contact (collidable.rs:1176-1209) + determine_sphere_plane_contact_geometry (impact_physics
// sphere.rs:138-158): corner voxels only; the plane { x : normal . x = displacement } in the space the transform starts from
void plane_voxel_object_contacts(const Object& obj, const Isometry& transform_to_object_space, V3 unit_normal,
                                 float displacement, std::vector<VoxelContact>& out);
void capsule_voxel_object_contacts(const Object& obj, const Isometry& T, V3 seg_start, V3 seg_vector, float radius,
                                   std::vector<VoxelContact>& out);
// voxel_ranges_within_plane (object/intersection.rs:751-761) with AxisAlignedBox::projected_onto_negative_halfspace
// (impact_geometry/src/axis_aligned_box.rs:460-488); plane in normalized voxel space
void voxel_ranges_within_plane(const uint32_t occupied[3][2], V3 unit_normal, float displacement, uint32_t out[3][2]);

// apply_mutual_absorption (interaction/absorption.rs:891-1080) given the voxel ranges encompassing the intersection
// (determine_voxel_ranges_encompassing_intersection, a pure function of the two occupied ranges and the transform that
// stays with the caller): both objects subtract each other's volume.
struct Isometry {
    Quat q;  // unit quaternion x, y, z, w
    V3 t;
};
void absorb_mutually(Object& a, Object& b, Isometry transform_from_b_to_a, float smoothness,
                     const uint32_t ranges_in_a[3][2], const uint32_t ranges_in_b[3][2], InertialUpdater* updater_a,
                     InertialUpdater* updater_b, AbsorbStats* stats_a, AbsorbStats* stats_b);

// --- connected regions (object/split_detection.rs, object/extraction.rs:121-281) ---
struct ChunkRegions {
    uint16_t region_count;           // local regions in the chunk (uniform chunk: 1)
    uint16_t boundary_region_count;  // of which touch the chunk boundary (labelled first)
    uint32_t first_region;           // index of the chunk's region 0 in the per-region arrays
};
struct RegionStats {
    uint32_t chunk_count, non_uniform_chunk_count;
    uint32_t chunk_min[3], chunk_max[3];
};
struct SplitDetection {
    std::vector<uint8_t> voxel_labels;  // 4096 per non-uniform chunk, data_offset order; 255 = empty
    std::vector<ChunkRegions> per_chunk;
    std::vector<uint32_t> region_parent;  // GlobalRegionLabel = chunk_idx << 8 | region_idx
    std::vector<uint32_t> region_root;    // resolved representative of every local region
    uint32_t n_regions = 0;               // count_regions
    bool has_two = false;                 // find_two_disconnected_regions
    uint32_t two[2] = {0, 0};
    RegionStats stats[2];
    uint32_t smallest = 0;                // which of `two` extract_smallest_region would extract
    bool overflow = false;                // more regions / connections than the reference's fixed capacities
};
// --- region extraction (object/extraction.rs) ---
struct Extraction {
    bool found_two = false;     // find_two_disconnected_regions found something to extract
    bool extracted = false;     // ExtractionResult::Extracted
    bool discarded = false;     // fewer than NON_EMPTY_VOXEL_THRESHOLD non-empty voxels: removed from the parent, dropped
    bool single_chunk = false;  // re-packed into one chunk (create_extracted_voxel_object_in_single_chunk_if_possible)
    uint32_t region_label = 0;
    uint32_t origin_offset_in_parent[3] = {0, 0, 0};  // voxels
    Object object;
};
// extract_any_disconnected_region: both objects end with their derived state up to date (connected regions are
// resolved again on demand by resolve_connected_regions).
void extract_any_disconnected_region(Object& obj, Extraction& out);
void update_all_chunk_boundary_adjacencies(Object& obj);

bool local_regions_for_chunk(const Voxel* voxels, bool only_empty, uint8_t* labels, uint16_t* boundary_region_count,
                             uint16_t* region_count);
void resolve_connected_regions(const Object& obj, SplitDetection& sd);
uint32_t count_regions_brute_force(const Object& obj);

}  // namespace orc
