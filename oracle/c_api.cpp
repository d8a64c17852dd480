// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
// extern "C" surface so tests/ and bench.py's cpu_baseline can drive the CPU
// restatement through ctypes. Never loaded by the product library.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>

#include "oracle.hpp"

using namespace orc;

extern "C" {

struct OrcTypeGen {
    uint32_t kind;
    uint32_t same_type;
    uint32_t n_types;
    float noise_frequency;
    float voxel_type_frequency;
    uint32_t seed;
};

struct OrcChunk {
    uint8_t kind;
    uint8_t flags;
    uint8_t face[6];  // [dim*2 + side]
    uint8_t uniform_type;
    int8_t uniform_sd;
    uint8_t uniform_flags;
    uint8_t _pad;
    uint32_t data_offset;
};

int orc_build_generator(const SdfNode* nodes, uint32_t n, uint32_t root, void** out, char* err,
                        size_t errcap) {
    Generator* g = new Generator();
    std::string e = build_generator(nodes, n, root, *g);
    if (!e.empty()) {
        if (err && errcap) std::snprintf(err, errcap, "%s", e.c_str());
        delete g;
        *out = nullptr;
        return 1;
    }
    *out = g;
    return 0;
}
uint32_t orc_generator_node_count(const void* g) { return (uint32_t)((const Generator*)g)->nodes.size(); }
uint32_t orc_generator_stack_size(const void* g) { return ((const Generator*)g)->stack_size; }
void orc_generator_nodes(const void* g, ProgNode* out) {
    const Generator* gen = (const Generator*)g;
    std::memcpy(out, gen->nodes.data(), gen->nodes.size() * sizeof(ProgNode));
}
void orc_generator_domain(const void* g, float lo[3], float hi[3]) {
    const Generator* gen = (const Generator*)g;
    lo[0] = gen->domain.lo.x; lo[1] = gen->domain.lo.y; lo[2] = gen->domain.lo.z;
    hi[0] = gen->domain.hi.x; hi[1] = gen->domain.hi.y; hi[2] = gen->domain.hi.z;
}
void orc_generator_free(void* g) { delete (Generator*)g; }

void orc_eval_chunk(const void* g, const float lo[3], float* out4096, uint8_t* decisions) {
    const Generator* gen = (const Generator*)g;
    std::vector<float> stack((size_t)(gen->stack_size + 1) * CHUNK_VOXELS);
    eval_chunk(*gen, v3(lo[0], lo[1], lo[2]), stack.data(), decisions);
    std::memcpy(out4096, stack.data(), sizeof(float) * CHUNK_VOXELS);
}
void orc_eval_block_preserving_gradients(const void* g, const float origin[3], int size, float* out) {
    const Generator* gen = (const Generator*)g;
    size_t count = (size_t)size * size * size;
    std::vector<float> stack((size_t)(gen->stack_size + 1) * count);
    eval_block_preserving_gradients(*gen, v3(origin[0], origin[1], origin[2]), size, stack.data());
    std::memcpy(out, stack.data(), sizeof(float) * count);
}

void* orc_voxel_generator_create(const void* g, float voxel_extent, const OrcTypeGen* t) {
    VoxelGenerator* vg = new VoxelGenerator();
    vg->sdf = *(const Generator*)g;
    vg->types.kind = t->kind;
    vg->types.same_type = (uint8_t)t->same_type;
    vg->types.n_types = t->n_types;
    vg->types.noise_frequency = t->noise_frequency;
    vg->types.voxel_type_frequency = t->voxel_type_frequency;
    vg->types.seed = t->seed;
    make_voxel_generator(*vg, voxel_extent);
    return vg;
}
void orc_voxel_generator_info(const void* vgp, uint32_t grid_shape[3], float shifted_center[3]) {
    const VoxelGenerator* vg = (const VoxelGenerator*)vgp;
    for (int d = 0; d < 3; ++d) grid_shape[d] = vg->grid_shape[d];
    shifted_center[0] = vg->shifted_center.x;
    shifted_center[1] = vg->shifted_center.y;
    shifted_center[2] = vg->shifted_center.z;
}
void orc_voxel_generator_free(void* vg) { delete (VoxelGenerator*)vg; }

void orc_generate_chunk(const void* vgp, const uint32_t origin[3], Voxel* out, uint8_t sparse[2]) {
    const VoxelGenerator* vg = (const VoxelGenerator*)vgp;
    std::vector<float> scratch((size_t)(vg->sdf.stack_size + 1) * CHUNK_VOXELS);
    Sparseness sp = generate_chunk(*vg, origin, out, scratch.data(), nullptr);
    sparse[0] = sp.only_empty;
    sparse[1] = sp.is_void;
}

void* orc_object_generate(const void* vgp, int n_threads, double* t_gen, double* t_derive) {
    Object* obj = new Object();
    generate_object(*(const VoxelGenerator*)vgp, *obj, n_threads, t_gen, t_derive);
    return obj;
}

// Bounded sample for the CPU baseline: generate + derive + mesh chunk planes [begin, end) as a slab.
void* orc_object_generate_slab(const void* vgp, int n_threads, uint32_t plane_begin, uint32_t plane_end,
                               double* t_gen, double* t_derive) {
    Object* obj = new Object();
    auto t0 = std::chrono::steady_clock::now();
    generate_slab_without_derived_state(*(const VoxelGenerator*)vgp, *obj, n_threads, plane_begin, plane_end);
    auto t1 = std::chrono::steady_clock::now();
    update_occupied_voxel_ranges(*obj);
    compute_all_derived_state(*obj);
    auto t2 = std::chrono::steady_clock::now();
    if (t_gen) *t_gen = std::chrono::duration<double>(t1 - t0).count();
    if (t_derive) *t_derive = std::chrono::duration<double>(t2 - t1).count();
    return obj;
}

// Test fixture equivalent to the reference's ManualVoxelGenerator
// (object.rs:3387-3561): dense per-voxel sd codes + types over `shape`.
void* orc_object_from_dense(const int8_t* sd, const uint8_t* types, const uint32_t shape[3],
                            float voxel_extent) {
    Object* objp = new Object();
    Object& obj = *objp;
    obj.voxel_extent = voxel_extent;
    for (int d = 0; d < 3; ++d) obj.chunk_counts[d] = (shape[d] + 15) / 16;
    uint32_t total = obj.chunk_counts[0] * obj.chunk_counts[1] * obj.chunk_counts[2];
    // Reuse the generic path by building a tiny generator-like loop here.
    obj.chunks.assign(total, Chunk{});
    std::vector<Voxel> buf(CHUNK_VOXELS);
    uint32_t nu = 0;
    for (uint32_t ci = 0; ci < total; ++ci) {
        uint32_t ijk[3] = {ci / (obj.chunk_counts[2] * obj.chunk_counts[1]),
                           (ci / obj.chunk_counts[2]) % obj.chunk_counts[1], ci % obj.chunk_counts[2]};
        bool only_empty = true, is_void = true;
        int idx = 0;
        for (int a = 0; a < 16; ++a)
            for (int b = 0; b < 16; ++b)
                for (int c = 0; c < 16; ++c, ++idx) {
                    uint32_t i = ijk[0] * 16 + a, j = ijk[1] * 16 + b, k = ijk[2] * 16 + c;
                    if (i >= shape[0] || j >= shape[1] || k >= shape[2]) {
                        buf[idx] = Voxel{TYPE_DUMMY, 127, FLAG_EMPTY};
                    } else {
                        size_t g = ((size_t)i * shape[1] + j) * shape[2] + k;
                        int8_t e = sd[g];
                        if (e < 0) {
                            only_empty = false;
                            is_void = false;
                            buf[idx] = Voxel{types[g], e, 0};
                        } else {
                            if (!(e > VOID_LIMIT)) is_void = false;
                            buf[idx] = Voxel{types[g], e, FLAG_EMPTY};
                        }
                    }
                }
        if (only_empty)
            for (auto& v : buf) v.type = TYPE_DUMMY;
        // classification identical to create_for_generated_voxels
        Chunk c;
        if (!is_void) {
            if (only_empty) {
                c.kind = CK_NONUNIFORM;
                c.flags = CF_ONLY_EMPTY;
            } else {
                Voxel first = buf[0];
                bool uniform = true;
                uint32_t cnt[3][2] = {{0, 0}, {0, 0}, {0, 0}};
                int q = 0;
                for (int a = 0; a < 16; ++a)
                    for (int b = 0; b < 16; ++b)
                        for (int cc = 0; cc < 16; ++cc, ++q) {
                            const Voxel& x = buf[q];
                            if (x.type != first.type || x.flags != first.flags || x.sd != -128) uniform = false;
                            if (x.flags & FLAG_EMPTY) {
                                if (a == 0) cnt[0][0]++; else if (a == 15) cnt[0][1]++;
                                if (b == 0) cnt[1][0]++; else if (b == 15) cnt[1][1]++;
                                if (cc == 0) cnt[2][0]++; else if (cc == 15) cnt[2][1]++;
                            }
                        }
                if (uniform) {
                    c.kind = CK_UNIFORM;
                    first.flags |= FLAG_FULL_ADJ;
                    c.uniform_voxel = first;
                } else {
                    c.kind = CK_NONUNIFORM;
                    for (int d = 0; d < 3; ++d)
                        for (int s = 0; s < 2; ++s)
                            c.face[d][s] = cnt[d][s] == 256 ? FD_EMPTY : (cnt[d][s] == 0 ? FD_FULL : FD_MIXED);
                }
            }
        }
        if (c.kind == CK_NONUNIFORM) {
            c.data_offset = nu++;
            obj.voxels.insert(obj.voxels.end(), buf.begin(), buf.end());
        }
        obj.chunks[ci] = c;
    }
    update_occupied_chunk_ranges(obj);
    update_occupied_voxel_ranges(obj);
    compute_all_derived_state(obj);
    return objp;
}

// VoxelObject::generate for a ChunkedVoxelGenerator given as data (the reference's test fixtures OffsetBoxVoxelGenerator /
// ManualVoxelGenerator, object.rs:3387-3561, or any other): raw chunks → classification → occupied ranges → derived state
void* orc_object_from_generated_chunks(const Voxel* voxels, const uint8_t* sparseness, const uint32_t grid_shape[3],
                                       float voxel_extent, int derive) {
    Object* o = new Object();
    object_from_generated_chunks(voxels, sparseness, grid_shape, voxel_extent, *o);
    if (derive) {
        update_occupied_voxel_ranges(*o);
        compute_all_derived_state(*o);
    }
    return o;
}

void orc_object_info(const void* op, uint32_t chunk_counts[3], uint64_t* n_voxels, uint32_t occ_chunks[6],
                     uint32_t occ_voxels[6]) {
    const Object* obj = (const Object*)op;
    for (int d = 0; d < 3; ++d) {
        chunk_counts[d] = obj->chunk_counts[d];
        occ_chunks[2 * d] = obj->occ_chunks[d][0];
        occ_chunks[2 * d + 1] = obj->occ_chunks[d][1];
        occ_voxels[2 * d] = obj->occ_voxels[d][0];
        occ_voxels[2 * d + 1] = obj->occ_voxels[d][1];
    }
    *n_voxels = obj->voxels.size();
}
void orc_object_chunks(const void* op, OrcChunk* out) {
    const Object* obj = (const Object*)op;
    for (size_t i = 0; i < obj->chunks.size(); ++i) {
        const Chunk& c = obj->chunks[i];
        OrcChunk o{};
        o.kind = c.kind;
        o.flags = c.flags;
        for (int d = 0; d < 3; ++d)
            for (int s = 0; s < 2; ++s) o.face[2 * d + s] = c.face[d][s];
        o.uniform_type = c.uniform_voxel.type;
        o.uniform_sd = c.uniform_voxel.sd;
        o.uniform_flags = c.uniform_voxel.flags;
        o.data_offset = c.data_offset;
        out[i] = o;
    }
}
void orc_object_voxels(const void* op, Voxel* out) {
    const Object* obj = (const Object*)op;
    std::memcpy(out, obj->voxels.data(), obj->voxels.size() * sizeof(Voxel));
}
uint32_t orc_object_dirty(const void* op, uint32_t* out, uint32_t cap) {
    const Object* obj = (const Object*)op;
    uint32_t n = (uint32_t)obj->dirty.size();
    for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = obj->dirty[i];
    return n;
}
void orc_object_clear_dirty(void* op) { ((Object*)op)->dirty.clear(); }
void orc_object_free(void* op) { delete (Object*)op; }

int orc_fill_brick(const void* op, uint32_t ci, uint32_t cj, uint32_t ck, float* values, uint8_t* types,
                   uint8_t adj6[6], uint8_t* flags) {
    static thread_local Brick b;
    std::memset(b.types, 255, sizeof(b.types));
    if (!fill_brick_if_exposed(*(const Object*)op, ci, cj, ck, b, flags)) return 0;
    std::memcpy(values, b.values, sizeof(b.values));
    std::memcpy(types, b.types, sizeof(b.types));
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) adj6[2 * d + s] = b.adj_non_uniform[d][s];
    return 1;
}

void* orc_mesh_create(const void* op, int n_threads, double* t_s) {
    Mesh* m = new Mesh();
    auto t0 = std::chrono::steady_clock::now();
    mesh_object(*(const Object*)op, *m, n_threads);
    auto t1 = std::chrono::steady_clock::now();
    if (t_s) *t_s = std::chrono::duration<double>(t1 - t0).count();
    return m;
}
void orc_mesh_sizes(const void* mp, uint32_t* n_vertices, uint32_t* n_indices, uint32_t* n_submeshes) {
    const Mesh* m = (const Mesh*)mp;
    *n_vertices = (uint32_t)(m->positions.size() / 3);
    *n_indices = (uint32_t)m->indices.size();
    *n_submeshes = (uint32_t)m->submeshes.size();
}
void orc_mesh_copy(const void* mp, float* positions, float* normals, IndexMaterials* index_materials,
                   uint32_t* indices, Submesh* submeshes, uint32_t* vertex_ranges) {
    const Mesh* m = (const Mesh*)mp;
    std::memcpy(positions, m->positions.data(), m->positions.size() * 4);
    std::memcpy(normals, m->normals.data(), m->normals.size() * 4);
    std::memcpy(index_materials, m->index_materials.data(), m->index_materials.size() * sizeof(IndexMaterials));
    std::memcpy(indices, m->indices.data(), m->indices.size() * 4);
    std::memcpy(submeshes, m->submeshes.data(), m->submeshes.size() * sizeof(Submesh));
    std::memcpy(vertex_ranges, m->vertex_ranges.data(), m->vertex_ranges.size() * 4);
}
void orc_mesh_free(void* mp) { delete (Mesh*)mp; }

// RangeAllocator on its own (pinned by the reference's unit tests, range_allocator.rs:150-249)
void* orc_range_allocator_create() { return new RangeAllocator(); }
void orc_range_allocator_free_range(void* a, uint64_t start, uint64_t end) { ((RangeAllocator*)a)->free_range(start, end); }
int orc_range_allocator_allocate(void* a, uint64_t len, uint64_t* start) {
    size_t s = 0;
    const bool ok = ((RangeAllocator*)a)->allocate_range(len, s);
    *start = s;
    return ok ? 1 : 0;
}
void orc_range_allocator_merge(void* a) { ((RangeAllocator*)a)->merge_consecutive_ranges(); }
void orc_range_allocator_destroy(void* a) { delete (RangeAllocator*)a; }

// ---- the mesh kept in sync with a modified object (VoxelObjectMesh::sync_with_voxel_object, mesh.rs:360-456) ----
void* orc_synced_mesh_create(const void* op, int n_threads) {
    SyncedMesh* sm = new SyncedMesh();
    synced_mesh_create(*(const Object*)op, *sm, n_threads);
    return sm;
}
void orc_synced_mesh_sync(void* smp, const void* op, const uint32_t* dirty, uint32_t n_dirty) {
    synced_mesh_sync(*(const Object*)op, *(SyncedMesh*)smp, dirty, n_dirty);
}
const void* orc_synced_mesh_mesh(const void* smp) { return &((const SyncedMesh*)smp)->mesh; }
// VoxelMeshModifications: → number of updated range records (4 words each) since the last report
uint32_t orc_synced_mesh_modifications(const void* smp, uint32_t* out, uint32_t capacity_records, int* chunks_were_removed) {
    const SyncedMesh* sm = (const SyncedMesh*)smp;
    const uint32_t n = (uint32_t)(sm->updated.size() / 4);
    if (out) std::memcpy(out, sm->updated.data(), (size_t)std::min(n, capacity_records) * 16);
    if (chunks_were_removed) *chunks_were_removed = sm->chunks_were_removed ? 1 : 0;
    return n;
}
void orc_synced_mesh_report_synchronized(void* smp) {
    SyncedMesh* sm = (SyncedMesh*)smp;
    sm->updated.clear();
    sm->chunks_were_removed = false;
}
void orc_synced_mesh_free(void* smp) { delete (SyncedMesh*)smp; }

// ---- collision probes (VoxelObjectCollisionProbes, collidable.rs:346-780) ----
void* orc_probes_create(const void* op, const void* mp) {
    CollisionProbes* pr = new CollisionProbes();
    probes_compute_for_all_chunks(*(const Object*)op, *(const Mesh*)mp, *pr);
    return pr;
}
// the same on mesh arrays the caller holds (oracle_lib.Mesh copies the arrays out and frees the C++ mesh)
void* orc_probes_create_arrays(const void* op, const float* positions, const float* normals, const uint32_t* indices,
                               const Submesh* submeshes, const uint32_t* vertex_ranges, uint32_t n_vertices, uint32_t n_indices,
                               uint32_t n_submeshes) {
    Mesh m;
    m.positions.assign(positions, positions + 3 * (size_t)n_vertices);
    m.normals.assign(normals, normals + 3 * (size_t)n_vertices);
    m.indices.assign(indices, indices + n_indices);
    m.submeshes.assign(submeshes, submeshes + n_submeshes);
    m.vertex_ranges.assign(vertex_ranges, vertex_ranges + 2 * (size_t)n_submeshes);
    CollisionProbes* pr = new CollisionProbes();
    probes_compute_for_all_chunks(*(const Object*)op, m, *pr);
    return pr;
}
void orc_probes_sync(void* pp, const void* op, const void* smp, const uint32_t* dirty, uint32_t n_dirty) {
    probes_sync(*(const Object*)op, *(const SyncedMesh*)smp, dirty, n_dirty, *(CollisionProbes*)pp);
}
uint32_t orc_probes_log2_block_size(const void* op) { return probes_log2_block_size(*(const Object*)op); }
void orc_probes_sizes(const void* pp, uint64_t* n_points, uint64_t* n_chunks) {
    const CollisionProbes* pr = (const CollisionProbes*)pp;
    *n_points = pr->points.size() / 3;
    *n_chunks = pr->range_of_chunk.size();
}
// ranges: per chunk with points {linear chunk index, start, end}, sorted by chunk index
void orc_probes_copy(const void* pp, float* points, uint32_t* ranges) {
    const CollisionProbes* pr = (const CollisionProbes*)pp;
    if (points) std::memcpy(points, pr->points.data(), pr->points.size() * sizeof(float));
    if (ranges) {
        std::vector<uint32_t> keys;
        for (const auto& kv : pr->range_of_chunk) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end());
        for (size_t q = 0; q < keys.size(); ++q) {
            const auto& r = pr->range_of_chunk.at(keys[q]);
            ranges[3 * q] = keys[q];
            ranges[3 * q + 1] = (uint32_t)r.first;
            ranges[3 * q + 2] = (uint32_t)r.second;
        }
    }
}
void orc_probes_free(void* pp) { delete (CollisionProbes*)pp; }
// for_each_mutual_voxel_object_contact: isometries as 7 floats (quaternion x, y, z, w, translation), moments as 10 floats
// (mass, moments, ...), ranges as [dim][start, end]. → the two contact lists concatenated; counts[0] = A against B
uint32_t orc_mutual_contacts(const void* oa, const void* pa, const float* moments_a, const float* world_to_a, const void* ob,
                             const void* pb, const float* moments_b, const float* world_to_b, const uint32_t* ranges_in_a,
                             const uint32_t* ranges_in_b, VoxelContact* out, uint32_t capacity, uint32_t* counts) {
    const auto iso = [](const float* f) { return Isometry{Quat{f[0], f[1], f[2], f[3]}, v3(f[4], f[5], f[6])}; };
    InertialMoments ia, ib;
    ia.mass = moments_a[0];
    ib.mass = moments_b[0];
    for (int d = 0; d < 3; ++d) {
        ia.moments[d] = moments_a[1 + d];
        ib.moments[d] = moments_b[1 + d];
    }
    uint32_t ra[3][2], rb[3][2];
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            ra[d][s] = ranges_in_a[2 * d + s];
            rb[d][s] = ranges_in_b[2 * d + s];
        }
    std::vector<VoxelContact> ab, ba;
    mutual_voxel_object_contacts(*(const Object*)oa, *(const CollisionProbes*)pa, ia, iso(world_to_a), *(const Object*)ob,
                                 *(const CollisionProbes*)pb, ib, iso(world_to_b), ra, rb, ab, ba);
    counts[0] = (uint32_t)ab.size();
    counts[1] = (uint32_t)ba.size();
    const uint32_t n = counts[0] + counts[1];
    if (out && n <= capacity) {
        if (!ab.empty()) std::memcpy(out, ab.data(), ab.size() * sizeof(VoxelContact));
        if (!ba.empty()) std::memcpy(out + ab.size(), ba.data(), ba.size() * sizeof(VoxelContact));
    }
    return n;
}
// add_points_for_vertices_in_blocks on caller-provided data (unit tests)
uint32_t orc_probes_points_for_chunk(uint32_t log2_block_size, const uint32_t* chunk_indices, const float* positions,
                                     const float* normals, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices,
                                     uint32_t start_index, float inverse_voxel_extent, float* out, uint32_t capacity_points) {
    std::vector<float> pts;
    probes_points_for_chunk(log2_block_size, chunk_indices, positions, normals, n_vertices, indices, n_indices, start_index,
                            inverse_voxel_extent, pts);
    const uint32_t n = (uint32_t)(pts.size() / 3);
    if (out) std::memcpy(out, pts.data(), (size_t)std::min(n, capacity_points) * 12);
    return n;
}

// Mesh one chunk (one iteration of sync_with_voxel_object). Returns 0 if the
// chunk is not exposed / produced no indices; else fills counts. Buffers sized
// for the worst case: 4913 vertices, 3*6*4913 indices.
int orc_mesh_chunk(const void* op, uint32_t ci, uint32_t cj, uint32_t ck, uint32_t* n_vertices,
                   uint32_t* n_indices, float* positions, float* normals, IndexMaterials* index_materials,
                   uint16_t* indices, uint8_t* flags) {
    ChunkMesh cm;
    if (!mesh_chunk(*(const Object*)op, ci, cj, ck, cm, flags)) return 0;
    *n_vertices = (uint32_t)(cm.positions.size() / 3);
    *n_indices = (uint32_t)cm.indices.size();
    std::memcpy(positions, cm.positions.data(), cm.positions.size() * 4);
    std::memcpy(normals, cm.normals.data(), cm.normals.size() * 4);
    std::memcpy(index_materials, cm.index_materials.data(), cm.index_materials.size() * sizeof(IndexMaterials));
    std::memcpy(indices, cm.indices.data(), cm.indices.size() * 2);
    return 1;
}

void orc_vertex_materials(const uint8_t has[8], const uint8_t mat[8], uint8_t out[16]) {
    bool h[8];
    for (int i = 0; i < 8; ++i) h[i] = has[i] != 0;
    VertexMaterials m = vertex_materials_compute(h, mat);
    std::memcpy(out, m.indices, 8);
    std::memcpy(out + 8, m.weights, 8);
}
void orc_index_materials(const uint8_t vm_in[48], uint8_t out[24]) {
    VertexMaterials vm[3];
    for (int v = 0; v < 3; ++v) {
        std::memcpy(vm[v].indices, vm_in + 16 * v, 8);
        std::memcpy(vm[v].weights, vm_in + 16 * v + 8, 8);
    }
    const VertexMaterials* p[3] = {&vm[0], &vm[1], &vm[2]};
    IndexMaterials im[3];
    index_materials_for_triangle(p, im);
    for (int v = 0; v < 3; ++v) {
        std::memcpy(out + 8 * v, im[v].indices, 4);
        std::memcpy(out + 8 * v + 4, im[v].weights, 4);
    }
}

void orc_absorb_sphere(void* op, const float center[3], float radius, float influence_radius,
                       AbsorbStats* stats) {
    absorb_sphere(*(Object*)op, v3(center[0], center[1], center[2]), radius, influence_radius, stats);
}

void orc_absorb_capsule(void* op, const float start[3], const float vec[3], float radius, float influence_radius,
                        AbsorbStats* stats) {
    absorb_capsule(*(Object*)op, v3(start[0], start[1], start[2]), v3(vec[0], vec[1], vec[2]), radius,
                   influence_radius, stats);
}

// ---- inertial properties (object/inertia.rs); moments = 10 f32: mass, moments[3], moments_of_inertia[3], products[3] ----
static_assert(sizeof(InertialMoments) == 40, "InertialMoments is 10 packed floats");
void orc_moments_for_voxel(float e, const float* densities, const uint32_t ijk[3], uint8_t type, float out[10]) {
    moments_for_voxel(e, e * e, (e * e) * e, densities, ijk, type, *(InertialMoments*)out);
}
void orc_moments_for_non_uniform_chunk(float e, const Voxel* voxels, const float* densities, const uint32_t cc[3],
                                       float out[10]) {
    moments_for_non_uniform_chunk(e, voxels, densities, cc, *(InertialMoments*)out);
}
void orc_moments_for_uniform_chunk(float e, const float* densities, uint8_t type, const uint32_t cc[3], float out[10]) {
    moments_for_uniform_chunk(e, densities, type, cc, *(InertialMoments*)out);
}
// per_chunk: NULL or 10 floats per chunk of the grid (zero for chunks that contribute nothing)
void orc_object_inertial_moments(const void* op, const float* densities, float out[10], float* per_chunk) {
    const Object& o = *(const Object*)op;
    if (per_chunk) std::memset(per_chunk, 0, o.chunks.size() * sizeof(InertialMoments));
    inertial_moments_for_object(o, densities, *(InertialMoments*)out, (InertialMoments*)per_chunk);
}
// absorption with the inertial-property updater attached (apply_*_absorption, absorption.rs:801-889)
void orc_absorb_sphere_inertial(void* op, const float center[3], float radius, float influence_radius,
                                AbsorbStats* stats, const float* densities, float moments[10]) {
    Object& o = *(Object*)op;
    InertialUpdater upd((InertialMoments*)moments, o.voxel_extent, densities);
    absorb_sphere(o, v3(center[0], center[1], center[2]), radius, influence_radius, stats, &upd);
}
void orc_absorb_capsule_inertial(void* op, const float start[3], const float vec[3], float radius,
                                 float influence_radius, AbsorbStats* stats, const float* densities,
                                 float moments[10]) {
    Object& o = *(Object*)op;
    InertialUpdater upd((InertialMoments*)moments, o.voxel_extent, densities);
    absorb_capsule(o, v3(start[0], start[1], start[2]), v3(vec[0], vec[1], vec[2]), radius, influence_radius, stats,
                   &upd);
}

// apply_mutual_absorption (absorption.rs:891-1080); ranges = [dim*2 + {start,end}]; densities NULL → no inertial updaters
void orc_absorb_mutually(void* ap, void* bp, const float q[4], const float t[3], float smoothness, const uint32_t ra[6],
                         const uint32_t rb[6], const float* densities, float moments_a[10], float moments_b[10],
                         AbsorbStats* stats_a, AbsorbStats* stats_b) {
    Object& a = *(Object*)ap;
    Object& b = *(Object*)bp;
    uint32_t r_a[3][2], r_b[3][2];
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            r_a[d][s] = ra[2 * d + s];
            r_b[d][s] = rb[2 * d + s];
        }
    InertialMoments dummy_a, dummy_b;
    InertialUpdater ua(densities ? (InertialMoments*)moments_a : &dummy_a, a.voxel_extent, densities);
    InertialUpdater ub(densities ? (InertialMoments*)moments_b : &dummy_b, b.voxel_extent, densities);
    absorb_mutually(a, b, Isometry{Quat{q[0], q[1], q[2], q[3]}, v3(t[0], t[1], t[2])}, smoothness, r_a, r_b,
                    densities ? &ua : nullptr, densities ? &ub : nullptr, stats_a, stats_b);
}

// for_each_surface_voxel_in_voxel_ranges: returns the number found, writes at most `capacity` 16-byte records
uint64_t orc_surface_voxels_in_ranges(const void* op, const uint32_t ranges[6], SurfaceVoxel* out, uint64_t capacity) {
    uint32_t r[3][2];
    for (int d = 0; d < 3; ++d) {
        r[d][0] = ranges[2 * d];
        r[d][1] = ranges[2 * d + 1];
    }
    std::vector<SurfaceVoxel> found;
    surface_voxels_in_ranges(*(const Object*)op, r, found);
    const size_t n = std::min<size_t>(found.size(), capacity);
    if (n) std::memcpy(out, found.data(), n * sizeof(SurfaceVoxel));
    return found.size();
}

// for_each_sphere_voxel_object_contact: returns the number of contacts, writes at most `capacity` 40-byte records
uint64_t orc_sphere_contacts(const void* op, const float q[4], const float t[3], const float center[3], float radius,
                             VoxelContact* out, uint64_t capacity) {
    std::vector<VoxelContact> found;
    sphere_voxel_object_contacts(*(const Object*)op, Isometry{Quat{q[0], q[1], q[2], q[3]}, v3(t[0], t[1], t[2])},
                                 v3(center[0], center[1], center[2]), radius, found);
    const size_t n = std::min<size_t>(found.size(), capacity);
    if (n) std::memcpy(out, found.data(), n * sizeof(VoxelContact));
    return found.size();
}

uint64_t orc_plane_contacts(const void* op, const float q[4], const float t[3], const float normal[3], float displacement,
                            VoxelContact* out, uint64_t capacity) {
    std::vector<VoxelContact> found;
    plane_voxel_object_contacts(*(const Object*)op, Isometry{Quat{q[0], q[1], q[2], q[3]}, v3(t[0], t[1], t[2])},
                                v3(normal[0], normal[1], normal[2]), displacement, found);
    const size_t n = std::min<size_t>(found.size(), capacity);
    if (n) std::memcpy(out, found.data(), n * sizeof(VoxelContact));
    return found.size();
}
uint64_t orc_capsule_contacts(const void* op, const float q[4], const float t[3], const float seg_start[3], const float seg_vector[3],
                              float radius, VoxelContact* out, uint64_t capacity) {
    std::vector<VoxelContact> found;
    capsule_voxel_object_contacts(*(const Object*)op, Isometry{Quat{q[0], q[1], q[2], q[3]}, v3(t[0], t[1], t[2])},
                                  v3(seg_start[0], seg_start[1], seg_start[2]), v3(seg_vector[0], seg_vector[1], seg_vector[2]),
                                  radius, found);
    const size_t n = std::min<size_t>(found.size(), capacity);
    if (n) std::memcpy(out, found.data(), n * sizeof(VoxelContact));
    return found.size();
}
void orc_voxel_ranges_within_plane(const uint32_t occ[6], const float normal[3], float displacement, uint32_t out[6]) {
    uint32_t o[3][2], r[3][2];
    for (int d = 0; d < 3; ++d) {
        o[d][0] = occ[2 * d];
        o[d][1] = occ[2 * d + 1];
    }
    voxel_ranges_within_plane(o, v3(normal[0], normal[1], normal[2]), displacement, r);
    for (int d = 0; d < 3; ++d) {
        out[2 * d] = r[d][0];
        out[2 * d + 1] = r[d][1];
    }
}

// ---- connected regions ----
// Runs the whole detection on the object's current state. info (u32 x 24): n_regions, has_two, two[0], two[1],
// smallest, overflow, n_region_entries, n_label_bytes, then per candidate region 8 words: chunk_count,
// non_uniform_chunk_count, chunk_min[3], chunk_max[3].
void* orc_split_detect(void* op, uint32_t* info) {
    auto* sd = new SplitDetection();
    resolve_connected_regions(*(Object*)op, *sd);
    info[0] = sd->n_regions;
    info[1] = sd->has_two ? 1u : 0u;
    info[2] = sd->two[0];
    info[3] = sd->two[1];
    info[4] = sd->smallest;
    info[5] = sd->overflow ? 1u : 0u;
    info[6] = (uint32_t)sd->region_root.size();
    info[7] = (uint32_t)sd->voxel_labels.size();
    for (int q = 0; q < 2; ++q) {
        uint32_t* o = info + 8 + 8 * q;
        o[0] = sd->stats[q].chunk_count;
        o[1] = sd->stats[q].non_uniform_chunk_count;
        for (int d = 0; d < 3; ++d) {
            o[2 + d] = sd->stats[q].chunk_min[d];
            o[5 + d] = sd->stats[q].chunk_max[d];
        }
    }
    return sd;
}
// per_chunk: n_chunks x {u16 region_count, u16 boundary_region_count, u32 first_region}
void orc_split_copy(void* sp, uint8_t* voxel_labels, void* per_chunk, uint32_t* region_roots) {
    auto* sd = (SplitDetection*)sp;
    if (voxel_labels) std::memcpy(voxel_labels, sd->voxel_labels.data(), sd->voxel_labels.size());
    if (per_chunk) std::memcpy(per_chunk, sd->per_chunk.data(), sd->per_chunk.size() * sizeof(ChunkRegions));
    if (region_roots) std::memcpy(region_roots, sd->region_root.data(), sd->region_root.size() * 4);
}
void orc_split_free(void* sp) { delete (SplitDetection*)sp; }

// ---- region extraction ----
// info (u32 x 8): found_two, extracted, discarded, single_chunk, region_label, origin_offset_in_parent[3].
// Returns the extracted object (caller frees with orc_object_free) or null.
void* orc_extract_any_disconnected_region(void* op, uint32_t* info) {
    Extraction ex;
    extract_any_disconnected_region(*(Object*)op, ex);
    info[0] = ex.found_two;
    info[1] = ex.extracted;
    info[2] = ex.discarded;
    info[3] = ex.single_chunk;
    info[4] = ex.region_label;
    for (int d = 0; d < 3; ++d) info[5 + d] = ex.origin_offset_in_parent[d];
    if (!ex.extracted) return nullptr;
    return new Object(std::move(ex.object));
}
uint32_t orc_count_regions_brute_force(void* op) { return count_regions_brute_force(*(Object*)op); }

float orc_simplex3(float x, float y, float z, int32_t seed) { return simplex3(x, y, z, seed); }
float orc_fbm3(float x, float y, float z, float lac, float gain, uint32_t oct, int32_t seed) {
    return fbm3(x, y, z, lac, gain, oct, seed);
}
float orc_simplex4(float x, float y, float z, float w, int32_t seed) { return simplex4(x, y, z, w, seed); }
int8_t orc_sd_encode(float v) { return sd_encode(v); }
float orc_sd_decode(int8_t e) { return sd_decode(e); }

}  // extern "C"
