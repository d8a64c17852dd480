// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
//
// Restates (V = engine/crates/impact_voxel/src):
//   fill_sdf_for_chunk_if_exposed + padding fills      V/object/sdf.rs:181-508
//   compute_sdf_gradient_from_corner_samples           V/object/sdf.rs:603-633
//   compute_surface_nets_mesh and helpers              V/object/sdf/surface_nets.rs:131-674
//   VoxelObjectMesh::recreate, ChunkSubmesh            V/mesh.rs:286-354, 559-635
#include <algorithm>
#include <cassert>
#include <cmath>
#include <thread>

#include "oracle.hpp"

namespace orc {

static inline int bidx(int i, int j, int k) { return i * 324 + j * 18 + k; }
static inline int vidx(int i, int j, int k) { return (i << 8) + (j << 4) + k; }

// The 18³ brick is a window of the object's voxel field shifted by -1: the
// reference fills it as interior + 6 faces + 12 edges + 8 corners, each from
// the one adjacent chunk that owns those cells (sdf.rs:215-508). Per adjacent
// chunk kind: Void / out of grid → +2.54 (types untouched), Uniform → -2.56 and
// its type, NonUniform → decoded voxel.
bool fill_brick_if_exposed(const Object& obj, uint32_t ci, uint32_t cj, uint32_t ck, Brick& b,
                           uint8_t* chunk_flags) {
    const Chunk& c = obj.chunks[obj.lin(ci, cj, ck)];
    if (c.kind != CK_NONUNIFORM || (c.flags & CF_OBSCURED_ALL) == CF_OBSCURED_ALL) return false;
    if (chunk_flags) *chunk_flags = c.flags;
    for (int d = 0; d < 3; ++d) b.adj_non_uniform[d][0] = b.adj_non_uniform[d][1] = false;
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj)
            for (int dk = -1; dk <= 1; ++dk) {
                Chunk a = (di == 0 && dj == 0 && dk == 0)
                              ? c
                              : obj.get_chunk((int64_t)ci + di, (int64_t)cj + dj, (int64_t)ck + dk);
                int nz = (di != 0) + (dj != 0) + (dk != 0);
                if (nz == 1 && a.kind == CK_NONUNIFORM) {
                    int dim = di != 0 ? 0 : (dj != 0 ? 1 : 2);
                    int side = (di + dj + dk) > 0 ? 1 : 0;
                    b.adj_non_uniform[dim][side] = true;
                }
                int r0[3], r1[3];  // brick cell ranges covered by this neighbour
                int dd[3] = {di, dj, dk};
                for (int d = 0; d < 3; ++d) {
                    if (dd[d] < 0) { r0[d] = 0; r1[d] = 1; }
                    else if (dd[d] == 0) { r0[d] = 1; r1[d] = 17; }
                    else { r0[d] = 17; r1[d] = 18; }
                }
                const Voxel* av = a.kind == CK_NONUNIFORM ? obj.chunk_voxels(a.data_offset) : nullptr;
                for (int i = r0[0]; i < r1[0]; ++i)
                    for (int j = r0[1]; j < r1[1]; ++j)
                        for (int k = r0[2]; k < r1[2]; ++k) {
                            int cell = bidx(i, j, k);
                            if (a.kind == CK_VOID) {
                                b.values[cell] = sd_decode(127);
                            } else if (a.kind == CK_UNIFORM) {
                                b.values[cell] = sd_decode(-128);
                                b.types[cell] = a.uniform_voxel.type;
                            } else {
                                const Voxel& v = av[vidx((i - 1) & 15, (j - 1) & 15, (k - 1) & 15)];
                                b.values[cell] = sd_decode(v.sd);
                                b.types[cell] = v.type;
                            }
                        }
            }
    return true;
}

// ---- surface_nets.rs:420-538 -------------------------------------------------
VertexMaterials vertex_materials_compute(const bool has_voxel[8], const uint8_t mat[8]) {
    VertexMaterials m{};
    uint8_t map[256];
    std::memset(map, 255, sizeof(map));
    int count = 0;
    for (int c = 0; c < 8; ++c) {
        if (has_voxel[c]) {
            uint8_t idx = map[mat[c]];
            if (idx == 255) {
                m.indices[count] = mat[c];
                m.weights[count] = 1;
                map[mat[c]] = (uint8_t)count;
                count++;
            } else {
                m.weights[idx] += 1;
            }
        }
    }
    m.indices[7] = (uint8_t)count;
    // sorting_network_7 (surface_nets.rs:429-446): 17 compare-swaps, swap iff w[i] < w[j]
    static const int NET[17][2] = {{0, 6}, {1, 5}, {2, 4}, {0, 3}, {1, 2}, {4, 5}, {0, 1}, {2, 3}, {4, 6},
                                   {5, 6}, {1, 4}, {3, 5}, {1, 2}, {3, 4}, {5, 6}, {2, 3}, {4, 5}};
    for (auto& p : NET) {
        int i = p[0], j = p[1];
        if (m.weights[i] < m.weights[j]) {
            std::swap(m.indices[i], m.indices[j]);
            std::swap(m.weights[i], m.weights[j]);
        }
    }
    return m;
}

// ---- surface_nets.rs:556-637 -------------------------------------------------
void index_materials_for_triangle(const VertexMaterials* vm[3], IndexMaterials out[3]) {
    auto count = [](const VertexMaterials* m) { return (int)m->indices[7]; };
    if (count(vm[0]) == 1 && count(vm[1]) == 1 && count(vm[2]) == 1) {
        uint8_t index = vm[0]->indices[0];
        if (vm[1]->indices[0] == index && vm[2]->indices[0] == index) {
            IndexMaterials im{{index, 0, 0, 0}, {1, 0, 0, 0}};
            out[0] = out[1] = out[2] = im;
            return;
        }
    }
    uint8_t top[4] = {0, 0, 0, 0};
    int n_top = 0;
    bool is_top[256] = {false};
    int off[3] = {0, 0, 0};
    for (int t = 0; t < 4; ++t) {
        uint8_t w[3];
        for (int i = 0; i < 3; ++i) w[i] = vm[i]->weights[off[i]];
        int mx = (w[0] >= w[1]) ? ((w[0] >= w[2]) ? 0 : 2) : ((w[1] >= w[2]) ? 1 : 2);
        if (w[mx] == 0) break;
        top[t] = vm[mx]->indices[off[mx]];
        n_top++;
        is_top[top[t]] = true;
        for (int i = 0; i < 3; ++i)
            while (off[i] < count(vm[i]) && is_top[vm[i]->indices[off[i]]]) off[i]++;
    }
    for (int v = 0; v < 3; ++v) {
        IndexMaterials im{};
        for (int i = 0; i < 4; ++i) im.indices[i] = top[i];
        for (int i = 0; i < n_top; ++i)
            for (int j = 0; j < count(vm[v]); ++j)
                if (vm[v]->indices[j] == top[i]) {
                    im.weights[i] = vm[v]->weights[j];
                    break;
                }
        out[v] = im;
    }
}

static const int CUBE_CORNERS[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1},
                                       {1, 0, 0}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1}};
static const int CUBE_EDGES[12][2] = {{0, 1}, {0, 2}, {0, 4}, {1, 3}, {1, 5}, {2, 3},
                                      {2, 6}, {3, 7}, {4, 5}, {4, 6}, {5, 7}, {6, 7}};

static inline bool opposite_signs(float a, float b) { return sign_neg(a) != sign_neg(b); }

// surface_nets.rs:384-418
static V3 centroid_of_edge_intersections(const float d[8]) {
    int count = 0;
    V3 sum = v3s(0.0f);
    for (auto& e : CUBE_EDGES) {
        float d1 = d[e[0]], d2 = d[e[1]];
        if (opposite_signs(d1, d2)) {
            count++;
            float interp1 = d1 / (d1 - d2);
            float interp2 = 1.0f - interp1;
            V3 c1 = v3((float)CUBE_CORNERS[e[0]][0], (float)CUBE_CORNERS[e[0]][1], (float)CUBE_CORNERS[e[0]][2]);
            V3 c2 = v3((float)CUBE_CORNERS[e[1]][0], (float)CUBE_CORNERS[e[1]][1], (float)CUBE_CORNERS[e[1]][2]);
            sum = sum + (interp2 * c1 + interp1 * c2);
        }
    }
    return sum / (float)count;
}

// object/sdf.rs:603-633
static V3 sdf_gradient(const float d[8], V3 o) {
    V3 p00 = v3(d[4], d[2], d[1]), n00 = v3(d[0], d[0], d[0]);
    V3 p01 = v3(d[5], d[6], d[3]), n01 = v3(d[1], d[4], d[2]);
    V3 p10 = v3(d[6], d[3], d[5]), n10 = v3(d[2], d[1], d[4]);
    V3 p11 = v3(d[7], d[7], d[7]), n11 = v3(d[3], d[5], d[6]);
    V3 d00 = p00 - n00, d01 = p01 - n01, d10 = p10 - n10, d11 = p11 - n11;
    V3 r = v3s(1.0f) - o;
    auto yzx = [](V3 a) { return v3(a.y, a.z, a.x); };
    auto zxy = [](V3 a) { return v3(a.z, a.x, a.y); };
    V3 t0 = mulc(mulc(yzx(r), zxy(r)), d00);
    V3 t1 = mulc(mulc(yzx(r), zxy(o)), d01);
    V3 t2 = mulc(mulc(yzx(o), zxy(r)), d10);
    V3 t3 = mulc(mulc(yzx(o), zxy(o)), d11);
    return ((t0 + t1) + t2) + t3;
}

void surface_nets(const Brick& b, float extent, V3 offset, ChunkMesh& out) {
    out.positions.clear();
    out.normals.clear();
    out.vertex_materials.clear();
    out.index_materials.clear();
    out.indices.clear();
    out.surface_lin.clear();
    out.lin_to_vertex.assign(BRICK_CELLS, 0xFFFF);

    // estimate_surface_nets_surface (surface_nets.rs:152-244)
    for (int i = 0; i < 17; ++i)
        for (int j = 0; j < 17; ++j)
            for (int k = 0; k < 17; ++k) {
                int lin = bidx(i, j, k);
                float d[8];
                bool has[8];
                int neg = 0;
                for (int c = 0; c < 8; ++c) {
                    int cl = lin + bidx(CUBE_CORNERS[c][0], CUBE_CORNERS[c][1], CUBE_CORNERS[c][2]);
                    d[c] = b.values[cl];
                    has[c] = sign_neg(d[c]);
                    neg += has[c] ? 1 : 0;
                }
                if (neg == 0 || neg == 8) continue;
                uint8_t mat[8];
                for (int c = 0; c < 8; ++c)
                    mat[c] = b.types[lin + bidx(CUBE_CORNERS[c][0], CUBE_CORNERS[c][1], CUBE_CORNERS[c][2])];
                V3 centroid = centroid_of_edge_intersections(d);
                V3 grad = sdf_gradient(d, centroid);
                V3 normal = grad / norm(grad);  // glam Vec3A::normalize (sse2): v / sqrt(dot)
                VertexMaterials vm = vertex_materials_compute(has, mat);
                V3 pos = extent * (centroid + v3((float)i, (float)j, (float)k)) + offset;
                out.lin_to_vertex[lin] = (uint16_t)(out.positions.size() / 3);
                out.surface_lin.push_back((uint16_t)lin);
                out.positions.push_back(pos.x); out.positions.push_back(pos.y); out.positions.push_back(pos.z);
                out.normals.push_back(normal.x); out.normals.push_back(normal.y); out.normals.push_back(normal.z);
                out.vertex_materials.push_back(vm);
            }

    // make_all_surface_nets_quads (surface_nets.rs:251-381)
    int upper[3] = {17, 17, 17};
    for (int d = 0; d < 3; ++d)
        if (b.adj_non_uniform[d][1]) upper[d] -= 1;
    auto pos_of = [&](uint16_t v) {
        return v3(out.positions[3 * v], out.positions[3 * v + 1], out.positions[3 * v + 2]);
    };
    auto maybe_quad = [&](int p1, int p2, int axb, int axc) {
        float d1 = b.values[p1], d2 = b.values[p2];
        bool n1 = sign_neg(d1), n2 = sign_neg(d2);
        bool negative_face;
        if (n1 && !n2) negative_face = false;
        else if (!n1 && n2) negative_face = true;
        else return;
        uint16_t v1 = out.lin_to_vertex[p1];
        uint16_t v2 = out.lin_to_vertex[p1 - axb];
        uint16_t v3_ = out.lin_to_vertex[p1 - axc];
        uint16_t v4 = out.lin_to_vertex[p1 - axb - axc];
        V3 a1 = pos_of(v1), a2 = pos_of(v2), a3 = pos_of(v3_), a4 = pos_of(v4);
        uint16_t q[6];
        if (norm(a1 - a4) < norm(a2 - a3)) {
            if (negative_face) { q[0]=v1; q[1]=v4; q[2]=v2; q[3]=v1; q[4]=v3_; q[5]=v4; }
            else               { q[0]=v1; q[1]=v2; q[2]=v4; q[3]=v1; q[4]=v4; q[5]=v3_; }
        } else if (negative_face) { q[0]=v2; q[1]=v3_; q[2]=v4; q[3]=v2; q[4]=v1; q[5]=v3_; }
        else                      { q[0]=v2; q[1]=v4; q[2]=v3_; q[3]=v2; q[4]=v3_; q[5]=v1; }
        out.indices.insert(out.indices.end(), q, q + 6);
    };
    for (uint16_t lin16 : out.surface_lin) {
        int lin = lin16;
        int i = lin / 324, j = (lin / 18) % 18, k = lin % 18;
        if (j != 0 && k != 0 && i < upper[0]) maybe_quad(lin, lin + 324, 18, 1);
        if (i != 0 && k != 0 && j < upper[1]) maybe_quad(lin, lin + 18, 1, 324);
        if (i != 0 && j != 0 && k < upper[2]) maybe_quad(lin, lin + 1, 324, 18);
    }

    // calculate_all_index_materials (surface_nets.rs:540-554)
    out.index_materials.resize(out.indices.size());
    for (size_t t = 0; t + 2 < out.indices.size(); t += 3) {
        const VertexMaterials* vm[3] = {&out.vertex_materials[out.indices[t]],
                                        &out.vertex_materials[out.indices[t + 1]],
                                        &out.vertex_materials[out.indices[t + 2]]};
        index_materials_for_triangle(vm, &out.index_materials[t]);
    }
}

// mesh.rs:559-577
static V3 vertex_position_offset_for_chunk(const Object& obj, uint32_t ci, uint32_t cj, uint32_t ck) {
    float extent = obj.voxel_extent;
    float chunk_extent = extent * 16.0f;
    return v3((float)ci * chunk_extent - 0.5f * extent, (float)cj * chunk_extent - 0.5f * extent,
              (float)ck * chunk_extent - 0.5f * extent);
}

bool mesh_chunk(const Object& obj, uint32_t ci, uint32_t cj, uint32_t ck, ChunkMesh& cm,
                uint8_t* chunk_flags) {
    static thread_local Brick brick;
    if (!fill_brick_if_exposed(obj, ci, cj, ck, brick, chunk_flags)) return false;
    surface_nets(brick, obj.voxel_extent, vertex_position_offset_for_chunk(obj, ci, cj, ck), cm);
    return !cm.indices.empty();
}

static void obscuredness_table(uint8_t flags, uint32_t t[2][2][2]) {
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                bool ox = flags & (1u << (i == 0 ? 0 : 3));
                bool oy = flags & (1u << (j == 0 ? 1 : 4));
                bool oz = flags & (1u << (k == 0 ? 2 : 5));
                t[i][j][k] = (ox && oy && oz) ? 1u : 0u;
            }
}

void mesh_object(const Object& obj, Mesh& mesh, int n_threads) {
    mesh = Mesh{};
    const uint32_t total = (uint32_t)obj.chunks.size();
    // The reference meshes chunks serially in i→j→k order. Meshing a chunk is a
    // pure function of the object, so with n_threads > 1 the chunk meshes are
    // computed in parallel and appended in the same order.
    std::vector<ChunkMesh> cms;
    std::vector<uint8_t> ok(total, 0), cflags(total, 0);
    std::vector<uint32_t> exposed;
    for (uint32_t c = 0; c < total; ++c) {
        const Chunk& ch = obj.chunks[c];
        if (ch.kind == CK_NONUNIFORM && (ch.flags & CF_OBSCURED_ALL) != CF_OBSCURED_ALL)
            exposed.push_back(c);
    }
    cms.resize(exposed.size());
    auto work = [&](int t, int nt) {
        for (size_t e = t; e < exposed.size(); e += nt) {
            uint32_t c = exposed[e];
            uint32_t i = c / (obj.chunk_counts[2] * obj.chunk_counts[1]);
            uint32_t j = (c / obj.chunk_counts[2]) % obj.chunk_counts[1];
            uint32_t k = c % obj.chunk_counts[2];
            ok[c] = mesh_chunk(obj, i, j, k, cms[e], &cflags[c]) ? 1 : 0;
        }
    };
    if (n_threads <= 1) {
        work(0, 1);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t, n_threads);
        for (auto& x : th) x.join();
    }
    for (size_t e = 0; e < exposed.size(); ++e) {
        uint32_t c = exposed[e];
        if (!ok[c]) continue;
        const ChunkMesh& cm = cms[e];
        uint32_t vertex_offset = (uint32_t)(mesh.positions.size() / 3);
        uint32_t index_offset = (uint32_t)mesh.indices.size();
        Submesh sm{};
        sm.chunk_indices[0] = c / (obj.chunk_counts[2] * obj.chunk_counts[1]);
        sm.chunk_indices[1] = (c / obj.chunk_counts[2]) % obj.chunk_counts[1];
        sm.chunk_indices[2] = c % obj.chunk_counts[2];
        sm.index_offset = index_offset;
        sm.index_count = (uint32_t)cm.indices.size();
        obscuredness_table(cflags[c], sm.obscured);
        mesh.submeshes.push_back(sm);
        mesh.vertex_ranges.push_back(vertex_offset);
        mesh.vertex_ranges.push_back(vertex_offset + (uint32_t)(cm.positions.size() / 3));
        mesh.positions.insert(mesh.positions.end(), cm.positions.begin(), cm.positions.end());
        mesh.normals.insert(mesh.normals.end(), cm.normals.begin(), cm.normals.end());
        mesh.index_materials.insert(mesh.index_materials.end(), cm.index_materials.begin(),
                                    cm.index_materials.end());
        for (uint16_t idx : cm.indices) mesh.indices.push_back(vertex_offset + (uint32_t)idx);
    }
}

// ---- RangeAllocator (impact_containers/src/range_allocator.rs:24-113) ----
void RangeAllocator::free_range(size_t start, size_t end) {
    if (start < end) free_ranges.emplace(start, end);  // BTreeSet::insert keeps an element that is already there
}
bool RangeAllocator::allocate_range(size_t required_len, size_t& start) {
    auto taken = free_ranges.end();
    size_t best_len = SIZE_MAX;
    for (auto it = free_ranges.begin(); it != free_ranges.end(); ++it) {
        const size_t len = it->second - it->first;
        if (len < best_len && len >= required_len) {
            taken = it;
            best_len = len;
        }
    }
    if (taken == free_ranges.end()) return false;
    const size_t s = taken->first, e = taken->second;
    free_ranges.erase(taken);
    if (s + required_len < e) free_ranges.emplace(s + required_len, e);
    start = s;
    return true;
}
void RangeAllocator::merge_consecutive_ranges() {
    if (free_ranges.size() < 2) return;
    std::map<size_t, size_t> merged;
    auto it = free_ranges.begin();
    size_t s = it->first, e = it->second;
    for (++it; it != free_ranges.end(); ++it) {
        if (it->first == e) {
            e = it->second;
        } else {
            merged.emplace(s, e);
            s = it->first;
            e = it->second;
        }
    }
    merged.emplace(s, e);
    free_ranges.swap(merged);
}

// ---- ChunkSubmeshManager (mesh.rs:703-848) on top of Mesh ----
void synced_mesh_create(const Object& obj, SyncedMesh& sm, int n_threads) {
    sm = SyncedMesh{};
    mesh_object(obj, sm.mesh, n_threads);
    for (uint32_t i = 0; i < sm.mesh.submeshes.size(); ++i) {
        const Submesh& s = sm.mesh.submeshes[i];
        const uint32_t c = (s.chunk_indices[0] * obj.chunk_counts[1] + s.chunk_indices[1]) * obj.chunk_counts[2] + s.chunk_indices[2];
        sm.index_of_chunk[c] = i;
        sm.chunk_at_index.push_back(c);
    }
}

static void remove_chunk_if_present(SyncedMesh& sm, uint32_t c) {
    auto it = sm.index_of_chunk.find(c);
    if (it == sm.index_of_chunk.end()) return;
    const uint32_t idx = it->second;
    sm.index_of_chunk.erase(it);
    // KeyIndexMapper::try_swap_remove_key + Vec::swap_remove on both tables
    const uint32_t last_key = sm.chunk_at_index.back();
    sm.chunk_at_index.pop_back();
    if (last_key != c) {
        sm.chunk_at_index[idx] = last_key;
        sm.index_of_chunk[last_key] = idx;
    }
    const uint32_t v0 = sm.mesh.vertex_ranges[2 * idx], v1 = sm.mesh.vertex_ranges[2 * idx + 1];
    const Submesh removed = sm.mesh.submeshes[idx];
    const size_t last = sm.mesh.submeshes.size() - 1;
    sm.mesh.submeshes[idx] = sm.mesh.submeshes[last];
    sm.mesh.submeshes.pop_back();
    sm.mesh.vertex_ranges[2 * idx] = sm.mesh.vertex_ranges[2 * last];
    sm.mesh.vertex_ranges[2 * idx + 1] = sm.mesh.vertex_ranges[2 * last + 1];
    sm.mesh.vertex_ranges.resize(2 * last);
    sm.free_vertices.free_range(v0, v1);
    sm.free_indices.free_range(removed.index_offset, removed.index_offset + removed.index_count);
    sm.chunks_were_removed = true;
}

void synced_mesh_sync(const Object& obj, SyncedMesh& sm, const uint32_t* dirty, size_t n_dirty) {
    Mesh& m = sm.mesh;
    ChunkMesh cm;
    for (size_t q = 0; q < n_dirty; ++q) {
        const uint32_t c = dirty[q];
        const uint32_t ci = c / (obj.chunk_counts[2] * obj.chunk_counts[1]), cj = (c / obj.chunk_counts[2]) % obj.chunk_counts[1],
                       ck = c % obj.chunk_counts[2];
        uint8_t flags = 0;
        // not exposed any more, or exposed with an empty mesh: the submesh goes (mesh.rs:379-383, 447-452)
        if (!mesh_chunk(obj, ci, cj, ck, cm, &flags)) {
            remove_chunk_if_present(sm, c);
            continue;
        }
        const size_t total_v = m.positions.size() / 3, total_i = m.indices.size();
        const size_t vcount = cm.positions.size() / 3, icount = cm.indices.size();
        // write_chunk (mesh.rs:749-812)
        auto found = sm.index_of_chunk.find(c);
        if (found != sm.index_of_chunk.end()) {
            const uint32_t idx = found->second;
            sm.free_vertices.free_range(m.vertex_ranges[2 * idx], m.vertex_ranges[2 * idx + 1]);
            sm.free_indices.free_range(m.submeshes[idx].index_offset, m.submeshes[idx].index_offset + m.submeshes[idx].index_count);
        }
        size_t v0 = total_v, i0 = total_i;
        if (!sm.free_vertices.allocate_range(vcount, v0)) v0 = total_v;
        if (!sm.free_indices.allocate_range(icount, i0)) i0 = total_i;
        Submesh s{};
        s.chunk_indices[0] = ci;
        s.chunk_indices[1] = cj;
        s.chunk_indices[2] = ck;
        s.index_offset = (uint32_t)i0;
        s.index_count = (uint32_t)icount;
        obscuredness_table(flags, s.obscured);
        if (found != sm.index_of_chunk.end()) {
            m.submeshes[found->second] = s;
            m.vertex_ranges[2 * found->second] = (uint32_t)v0;
            m.vertex_ranges[2 * found->second + 1] = (uint32_t)(v0 + vcount);
        } else {
            sm.index_of_chunk[c] = (uint32_t)m.submeshes.size();
            sm.chunk_at_index.push_back(c);
            m.submeshes.push_back(s);
            m.vertex_ranges.push_back((uint32_t)v0);
            m.vertex_ranges.push_back((uint32_t)(v0 + vcount));
        }
        sm.updated.push_back((uint32_t)v0);
        sm.updated.push_back((uint32_t)(v0 + vcount));
        sm.updated.push_back((uint32_t)i0);
        sm.updated.push_back((uint32_t)(i0 + icount));
        // the data: appended when no free range fitted, else over the obsolete values (mesh.rs:399-445)
        if (v0 == total_v) {
            m.positions.insert(m.positions.end(), cm.positions.begin(), cm.positions.end());
            m.normals.insert(m.normals.end(), cm.normals.begin(), cm.normals.end());
        } else {
            std::copy(cm.positions.begin(), cm.positions.end(), m.positions.begin() + 3 * v0);
            std::copy(cm.normals.begin(), cm.normals.end(), m.normals.begin() + 3 * v0);
        }
        if (i0 == total_i) {
            m.index_materials.insert(m.index_materials.end(), cm.index_materials.begin(), cm.index_materials.end());
            for (uint16_t idx : cm.indices) m.indices.push_back((uint32_t)v0 + (uint32_t)idx);
        } else {
            std::copy(cm.index_materials.begin(), cm.index_materials.end(), m.index_materials.begin() + i0);
            for (size_t t = 0; t < icount; ++t) m.indices[i0 + t] = (uint32_t)v0 + (uint32_t)cm.indices[t];
        }
    }
    // perform_maintainance (mesh.rs:828-831)
    sm.free_vertices.merge_consecutive_ranges();
    sm.free_indices.merge_consecutive_ranges();
}

}  // namespace orc
