// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
//
// Restates (V = engine/crates/impact_voxel/src):
//   SDFVoxelGenerator::new / generate_chunk        V/generation.rs:207-371
//   VoxelTypeGenerator::set_voxel_types_for_chunk  V/generation/voxel_type.rs:54-168
//   VoxelObject::generate*                         V/object.rs:239-404
//   analyze_and_initialize_chunks                  V/object.rs:560-617
//   occupied ranges                                V/object.rs:1149-1280
//   compute_all_derived_state                      V/object.rs:1136-1145
//   boundary adjacencies / obscuredness            V/object.rs:1659-1785, 2077-2652
//   chunk classification                           V/object.rs:1890-1964
//   internal adjacencies / internal state          V/object.rs:2673-2874
// Split detection (connected regions) is out of this round's scope and is not
// restated; it does not influence voxels, chunk kinds, flags or meshes.
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <thread>

#include "oracle.hpp"

namespace orc {

static inline int vidx(int i, int j, int k) { return (i << 8) + (j << 4) + k; }

// ---- SDFVoxelGenerator::new (generation.rs:207-258) -------------------------
void make_voxel_generator(VoxelGenerator& vg, float voxel_extent) {
    vg.voxel_extent = voxel_extent;
    V3 ext = aabb_extents(vg.sdf.domain);
    if (ext.x == 0.0f || ext.y == 0.0f || ext.z == 0.0f) {
        vg.grid_shape[0] = vg.grid_shape[1] = vg.grid_shape[2] = 0;
        vg.shifted_center = v3s(-0.5f);
        return;
    }
    float e[3] = {ext.x, ext.y, ext.z};
    float half[3];
    for (int d = 0; d < 3; ++d) {
        float c = std::ceil(e[d]);
        // `as usize` saturates at 0 for negatives / NaN
        uint32_t n = (c > 0.0f) ? (uint32_t)c : 0u;
        vg.grid_shape[d] = n + 2;
        half[d] = 0.5f * (float)vg.grid_shape[d];
    }
    V3 dc = aabb_center(vg.sdf.domain);
    V3 center_rel_origin = v3(half[0], half[1], half[2]) - dc;
    vg.shifted_center = center_rel_origin - v3s(0.5f);
}

// ---- voxel_type.rs:125-168 (GradientNoise) ----------------------------------
static void set_types_gradient_noise(const TypeGen& tg, Voxel* voxels, V3 chunk_origin,
                                     float* /*scratch*/) {
    // gradient_4d_offset(0, n, o.z, 16, o.y, 16, o.x, 16) with freqs
    // (type_freq, f, f, f). simdnoise walks its x as x_arr[l] = 0 + l, and its
    // y / z / w (our k / j / i) by repeated `+= 1.0`.
    const uint32_t n = tg.n_types;
    int idx = 0;
    float wc = chunk_origin.x;
    for (int i = 0; i < 16; ++i) {
        float zc = chunk_origin.y;
        for (int j = 0; j < 16; ++j) {
            float yc = chunk_origin.z;
            for (int k = 0; k < 16; ++k) {
                float max_noise = 0.0f;
                uint32_t max_idx = 0;
                for (uint32_t t = 0; t < n; ++t) {
                    // simdnoise x: `x_arr[l] = start + l` then `+= VW` per full vector (VW = 8)
                    float xc = 0.0f + (float)(t % 8u);
                    for (uint32_t v = 0; v < t / 8u; ++v) xc = xc + 8.0f;
                    float nv = simplex4(xc * tg.voxel_type_frequency, yc * tg.noise_frequency,
                                        zc * tg.noise_frequency, wc * tg.noise_frequency,
                                        (int32_t)tg.seed);
                    if (t == 0 || nv > max_noise) {
                        max_noise = nv;
                        max_idx = t;
                    }
                }
                // VoxelType::from_idx(max_idx): the INDEX is the type (voxel_type.rs:166)
                voxels[idx].type = (uint8_t)max_idx;
                idx++;
                yc = yc + 1.0f;
            }
            zc = zc + 1.0f;
        }
        wc = wc + 1.0f;
    }
}

// ---- generate_chunk (generation.rs:293-371) ---------------------------------
Sparseness generate_chunk(const VoxelGenerator& vg, const uint32_t origin[3], Voxel* voxels,
                          float* scratch, float* type_scratch) {
    const Voxel outside{TYPE_DUMMY, 127, FLAG_EMPTY};
    if (vg.sdf.empty() || origin[0] >= vg.grid_shape[0] || origin[1] >= vg.grid_shape[1] ||
        origin[2] >= vg.grid_shape[2]) {
        for (int i = 0; i < CHUNK_VOXELS; ++i) voxels[i] = outside;
        return Sparseness{true, true};
    }
    V3 lo = v3((float)origin[0], (float)origin[1], (float)origin[2]) - vg.shifted_center;
    eval_chunk(vg.sdf, lo, scratch, nullptr);
    const float* sd = scratch;
    bool only_empty = true, is_void = true;
    int idx = 0;
    for (int ic = 0; ic < 16; ++ic)
        for (int jc = 0; jc < 16; ++jc)
            for (int kc = 0; kc < 16; ++kc, ++idx) {
                uint32_t i = origin[0] + ic, j = origin[1] + jc, k = origin[2] + kc;
                if (i >= vg.grid_shape[0] || j >= vg.grid_shape[1] || k >= vg.grid_shape[2]) {
                    voxels[idx] = outside;
                } else {
                    int8_t e = sd_encode(sd[idx]);
                    if (e < 0) {
                        only_empty = false;
                        is_void = false;
                        voxels[idx] = Voxel{TYPE_DUMMY, e, 0};
                    } else {
                        if (!(e > VOID_LIMIT)) is_void = false;
                        voxels[idx] = Voxel{TYPE_DUMMY, e, FLAG_EMPTY};
                    }
                }
            }
    if (!only_empty) {
        if (vg.types.kind == 0) {
            for (int i = 0; i < CHUNK_VOXELS; ++i) voxels[i].type = vg.types.same_type;
        } else {
            set_types_gradient_noise(vg.types, voxels, lo, type_scratch);
        }
    }
    return Sparseness{only_empty, is_void};
}

// ---- create_for_generated_voxels (object.rs:1890-1964) ----------------------
static void face_dists_from_empty_counts(const uint32_t cnt[3][2], uint8_t face[3][2]) {
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s)
            face[d][s] = cnt[d][s] == 256 ? FD_EMPTY : (cnt[d][s] == 0 ? FD_FULL : FD_MIXED);
}

static Chunk classify_generated_chunk(const Voxel* v, Sparseness sp) {
    Chunk c;
    if (sp.is_void) return c;
    if (sp.only_empty) {
        c.kind = CK_NONUNIFORM;
        c.flags = CF_ONLY_EMPTY;
        return c;  // all faces Empty
    }
    Voxel first = v[0];
    bool uniform = true;
    uint32_t cnt[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    int idx = 0;
    for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 16; ++j)
            for (int k = 0; k < 16; ++k, ++idx) {
                const Voxel& x = v[idx];
                if (uniform && (!(x.type == first.type && x.flags == first.flags) || x.sd != -128))
                    uniform = false;
                if (x.flags & FLAG_EMPTY) {
                    if (i == 0) cnt[0][0]++; else if (i == 15) cnt[0][1]++;
                    if (j == 0) cnt[1][0]++; else if (j == 15) cnt[1][1]++;
                    if (k == 0) cnt[2][0]++; else if (k == 15) cnt[2][1]++;
                }
            }
    if (uniform) {
        c.kind = CK_UNIFORM;
        first.flags |= FLAG_FULL_ADJ;
        c.uniform_voxel = first;
    } else {
        c.kind = CK_NONUNIFORM;
        face_dists_from_empty_counts(cnt, c.face);
        c.flags = 0;
    }
    return c;
}

void Object::mark_dirty(uint32_t idx) {
    if (std::find(dirty.begin(), dirty.end(), idx) == dirty.end()) dirty.push_back(idx);
}

static inline bool chunk_only_empty(const Chunk& c) {
    return c.kind == CK_VOID || (c.kind == CK_NONUNIFORM && (c.flags & CF_ONLY_EMPTY));
}

// analyze_and_initialize_chunks (object.rs:560-617): data offsets in chunk order, chunk-granular occupied ranges
static void analyze_and_initialize_chunks(Object& obj) {
    const uint32_t total = (uint32_t)obj.chunks.size();
    uint32_t nu = 0;
    uint32_t lo[3] = {UINT32_MAX, UINT32_MAX, UINT32_MAX}, hi[3] = {0, 0, 0};
    bool any = false;
    for (uint32_t ci = 0; ci < total; ++ci) {
        Chunk& c = obj.chunks[ci];
        if (c.kind == CK_NONUNIFORM) c.data_offset = nu++;
        if (!chunk_only_empty(c)) {
            uint32_t ijk[3] = {ci / (obj.chunk_counts[2] * obj.chunk_counts[1]),
                               (ci / obj.chunk_counts[2]) % obj.chunk_counts[1],
                               ci % obj.chunk_counts[2]};
            for (int d = 0; d < 3; ++d) {
                lo[d] = std::min(lo[d], ijk[d]);
                hi[d] = std::max(hi[d], ijk[d] + 1);
            }
            any = true;
        }
    }
    for (int d = 0; d < 3; ++d) {
        obj.occ_chunks[d][0] = any ? lo[d] : 0;
        obj.occ_chunks[d][1] = any ? hi[d] : 0;
        obj.occ_voxels[d][0] = obj.occ_chunks[d][0] * 16;
        obj.occ_voxels[d][1] = obj.occ_chunks[d][1] * 16;
    }
}

// VoxelObject::generate_without_derived_state (object.rs:307-359 → generate_voxels_for_chunks :361-404) for any
// ChunkedVoxelGenerator whose output is given as data: 4096 voxels and the ChunkSparseness per chunk, linear chunk order
void object_from_generated_chunks(const Voxel* voxels, const uint8_t* sparseness, const uint32_t grid_shape[3],
                                  float voxel_extent, Object& obj) {
    obj = Object{};
    obj.voxel_extent = voxel_extent;
    for (int d = 0; d < 3; ++d) obj.chunk_counts[d] = (grid_shape[d] + 15) / 16;
    const uint32_t total = obj.chunk_counts[0] * obj.chunk_counts[1] * obj.chunk_counts[2];
    obj.chunks.assign(total, Chunk{});
    if (total == 0) return;
    for (uint32_t ci = 0; ci < total; ++ci) {
        const Voxel* buf = voxels + (size_t)ci * CHUNK_VOXELS;
        const Sparseness sp{(sparseness[ci] & 1) != 0, (sparseness[ci] & 2) != 0};
        const Chunk c = classify_generated_chunk(buf, sp);
        obj.chunks[ci] = c;
        if (c.kind == CK_NONUNIFORM) obj.voxels.insert(obj.voxels.end(), buf, buf + CHUNK_VOXELS);
    }
    analyze_and_initialize_chunks(obj);
}

// ---- generate_voxels_for_chunks[_in_parallel] + analyze (object.rs:361-617) --
void generate_without_derived_state(const VoxelGenerator& vg, Object& obj, int n_threads) {
    generate_slab_without_derived_state(vg, obj, n_threads, 0, (vg.grid_shape[0] + 15) / 16);
}

void generate_slab_without_derived_state(const VoxelGenerator& vg, Object& obj, int n_threads,
                                         uint32_t plane_begin, uint32_t plane_end) {
    obj = Object{};
    obj.voxel_extent = vg.voxel_extent;
    for (int d = 0; d < 3; ++d) obj.chunk_counts[d] = (vg.grid_shape[d] + 15) / 16;
    if (plane_end > obj.chunk_counts[0]) plane_end = obj.chunk_counts[0];
    if (plane_begin > plane_end) plane_begin = plane_end;
    obj.chunk_counts[0] = plane_end - plane_begin;
    const uint32_t total = obj.chunk_counts[0] * obj.chunk_counts[1] * obj.chunk_counts[2];
    obj.chunks.assign(total, Chunk{});
    if (total == 0) return;

    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > total) n_threads = (int)total;
    struct Part {
        std::vector<Voxel> voxels;
    };
    std::vector<Part> parts(n_threads);
    auto work = [&](int t) {
        // contiguous ranges of the x-major linear chunk index (object.rs:423-427)
        uint32_t per = (total + n_threads - 1) / n_threads;
        uint32_t begin = std::min(total, per * (uint32_t)t);
        uint32_t end = std::min(total, begin + per);
        std::vector<float> scratch((size_t)(vg.sdf.stack_size + 1) * CHUNK_VOXELS);
        std::vector<float> tscratch;
        Voxel buf[CHUNK_VOXELS];
        for (uint32_t ci = begin; ci < end; ++ci) {
            uint32_t i = ci / (obj.chunk_counts[2] * obj.chunk_counts[1]);
            uint32_t j = (ci / obj.chunk_counts[2]) % obj.chunk_counts[1];
            uint32_t k = ci % obj.chunk_counts[2];
            uint32_t origin[3] = {(i + plane_begin) * 16, j * 16, k * 16};
            Sparseness sp = generate_chunk(vg, origin, buf, scratch.data(), tscratch.data());
            Chunk c = classify_generated_chunk(buf, sp);
            obj.chunks[ci] = c;
            if (c.kind == CK_NONUNIFORM)
                parts[t].voxels.insert(parts[t].voxels.end(), buf, buf + CHUNK_VOXELS);
        }
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    size_t nv = 0;
    for (auto& p : parts) nv += p.voxels.size();
    obj.voxels.reserve(nv);
    for (auto& p : parts) obj.voxels.insert(obj.voxels.end(), p.voxels.begin(), p.voxels.end());

    analyze_and_initialize_chunks(obj);
}

void update_occupied_chunk_ranges(Object& obj) {
    uint32_t lo[3] = {UINT32_MAX, UINT32_MAX, UINT32_MAX}, hi[3] = {0, 0, 0};
    bool any = false;
    for (uint32_t i = 0; i < obj.chunk_counts[0]; ++i)
        for (uint32_t j = 0; j < obj.chunk_counts[1]; ++j)
            for (uint32_t k = 0; k < obj.chunk_counts[2]; ++k) {
                if (!chunk_only_empty(obj.chunks[obj.lin(i, j, k)])) {
                    uint32_t ijk[3] = {i, j, k};
                    for (int d = 0; d < 3; ++d) {
                        lo[d] = std::min(lo[d], ijk[d]);
                        hi[d] = std::max(hi[d], ijk[d]);
                    }
                    any = true;
                }
            }
    for (int d = 0; d < 3; ++d) {
        obj.occ_chunks[d][0] = any ? lo[d] : 0;
        obj.occ_chunks[d][1] = any ? hi[d] + 1 : 0;
    }
}

static bool find_bound_in_chunk(const Voxel* v, int dim, int side, uint32_t& out) {
    // Loop3::over_all_from_side(dim, side): primary axis = dim, traversed from `side`
    for (int a = 0; a < 16; ++a) {
        int p = side == 0 ? a : 15 - a;
        for (int b = 0; b < 16; ++b)
            for (int c = 0; c < 16; ++c) {
                int ijk[3];
                ijk[dim] = p;
                ijk[(dim + 1) % 3] = b;
                ijk[(dim + 2) % 3] = c;
                if (!(v[vidx(ijk[0], ijk[1], ijk[2])].flags & FLAG_EMPTY)) {
                    out = (uint32_t)p;
                    return true;
                }
            }
    }
    return false;
}

static uint32_t find_voxel_bound(const Object& obj, int dim, int side) {
    int o0 = dim == 0 ? 1 : 0;
    int o1 = dim == 2 ? 1 : 2;
    uint32_t chunk_l = side == 0 ? obj.occ_chunks[dim][0] : obj.occ_chunks[dim][1] - 1;
    uint32_t bound = side == 0 ? UINT32_MAX : 0;
    uint32_t start = chunk_l * 16;
    for (uint32_t m = obj.occ_chunks[o0][0]; m < obj.occ_chunks[o0][1]; ++m)
        for (uint32_t n = obj.occ_chunks[o1][0]; n < obj.occ_chunks[o1][1]; ++n) {
            uint32_t ci[3];
            ci[dim] = chunk_l;
            ci[o0] = m;
            ci[o1] = n;
            const Chunk& c = obj.chunks[obj.lin(ci[0], ci[1], ci[2])];
            if (c.kind == CK_UNIFORM) return side == 0 ? start : start + 15;
            if (c.kind == CK_NONUNIFORM) {
                uint32_t b;
                if (find_bound_in_chunk(obj.chunk_voxels(c.data_offset), dim, side, b))
                    bound = side == 0 ? std::min(bound, start + b) : std::max(bound, start + b);
            }
        }
    return bound;
}

void update_occupied_voxel_ranges(Object& obj) {
    bool empty = false;
    for (int d = 0; d < 3; ++d)
        if (obj.occ_chunks[d][0] >= obj.occ_chunks[d][1]) empty = true;
    if (empty) {
        for (int d = 0; d < 3; ++d) obj.occ_voxels[d][0] = obj.occ_voxels[d][1] = 0;
        return;
    }
    for (int d = 0; d < 3; ++d) {
        uint32_t first = find_voxel_bound(obj, d, 0);
        uint32_t last = find_voxel_bound(obj, d, 1);
        obj.occ_voxels[d][0] = first;
        obj.occ_voxels[d][1] = last + 1;
    }
}

// ---- internal adjacencies (object.rs:2673-2756) ------------------------------
static const uint8_t UP_FLAG[3] = {FLAG_ADJ_X_UP, FLAG_ADJ_Y_UP, FLAG_ADJ_Z_UP};
static const uint8_t DN_FLAG[3] = {FLAG_ADJ_X_DN, FLAG_ADJ_Y_DN, FLAG_ADJ_Z_DN};

void update_internal_adjacencies(Voxel* v) {
    for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 16; ++j)
            for (int k = 0; k < 16; ++k) {
                int idx = vidx(i, j, k);
                Voxel voxel = v[idx];
                int ijk[3] = {i, j, k};
                if (voxel.flags & FLAG_EMPTY) {
                    for (int d = 0; d < 3; ++d) {
                        if (ijk[d] + 1 < 16) {
                            int a[3] = {i, j, k};
                            a[d] += 1;
                            v[vidx(a[0], a[1], a[2])].flags &= (uint8_t)~DN_FLAG[d];
                        }
                    }
                } else {
                    uint8_t flags = voxel.flags;
                    for (int d = 0; d < 3; ++d) {
                        if (ijk[d] + 1 < 16) {
                            int a[3] = {i, j, k};
                            a[d] += 1;
                            Voxel& adj = v[vidx(a[0], a[1], a[2])];
                            if (adj.flags & FLAG_EMPTY) {
                                flags &= (uint8_t)~UP_FLAG[d];
                            } else {
                                flags |= UP_FLAG[d];
                                adj.flags |= DN_FLAG[d];
                            }
                        }
                    }
                    v[idx].flags = flags;
                }
            }
}

// update_all_internal_state_and_determine_sparseness (object.rs:2761-2874)
Sparseness update_all_internal_state(Chunk& c, Voxel* v) {
    uint32_t cnt[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    bool only_empty = true, is_void = true;
    for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 16; ++j)
            for (int k = 0; k < 16; ++k) {
                int idx = vidx(i, j, k);
                Voxel voxel = v[idx];
                int ijk[3] = {i, j, k};
                if (voxel.flags & FLAG_EMPTY) {
                    if (i == 0) cnt[0][0]++; else if (i == 15) cnt[0][1]++;
                    if (j == 0) cnt[1][0]++; else if (j == 15) cnt[1][1]++;
                    if (k == 0) cnt[2][0]++; else if (k == 15) cnt[2][1]++;
                    for (int d = 0; d < 3; ++d) {
                        if (ijk[d] + 1 < 16) {
                            int a[3] = {i, j, k};
                            a[d] += 1;
                            v[vidx(a[0], a[1], a[2])].flags &= (uint8_t)~DN_FLAG[d];
                        }
                    }
                    if (!(voxel.sd > VOID_LIMIT)) is_void = false;
                } else {
                    uint8_t flags = voxel.flags;
                    for (int d = 0; d < 3; ++d) {
                        if (ijk[d] + 1 < 16) {
                            int a[3] = {i, j, k};
                            a[d] += 1;
                            Voxel& adj = v[vidx(a[0], a[1], a[2])];
                            if (adj.flags & FLAG_EMPTY) {
                                flags &= (uint8_t)~UP_FLAG[d];
                            } else {
                                flags |= UP_FLAG[d];
                                adj.flags |= DN_FLAG[d];
                            }
                        }
                    }
                    v[idx].flags = flags;
                    only_empty = false;
                    is_void = false;
                }
            }
    face_dists_from_empty_counts(cnt, c.face);
    if (only_empty)
        c.flags |= CF_ONLY_EMPTY;
    else
        c.flags &= (uint8_t)~CF_ONLY_EMPTY;
    return Sparseness{only_empty, is_void};
}

// ---- boundary adjacencies (object.rs:2077-2652) ------------------------------
static inline uint8_t adj_flag_for_face(int dim, int side) {
    return side == 0 ? DN_FLAG[dim] : UP_FLAG[dim];
}

template <typename F>
static void for_each_face_voxel(int dim, int side, F&& f) {
    int p = side == 0 ? 0 : 15;
    for (int b = 0; b < 16; ++b)
        for (int c = 0; c < 16; ++c) {
            int ijk[3];
            ijk[dim] = p;
            // remaining dims in ascending order (i before j before k)
            int d0 = dim == 0 ? 1 : 0, d1 = dim == 2 ? 1 : 2;
            ijk[d0] = b;
            ijk[d1] = c;
            f(ijk[0], ijk[1], ijk[2]);
        }
}

static void convert_to_non_uniform_if_uniform(Object& obj, Chunk& c) {
    if (c.kind != CK_UNIFORM) return;
    size_t start = obj.voxels.size();
    obj.voxels.resize(start + CHUNK_VOXELS, c.uniform_voxel);
    c.kind = CK_NONUNIFORM;
    c.data_offset = (uint32_t)(start >> 12);
    for (int d = 0; d < 3; ++d) c.face[d][0] = c.face[d][1] = FD_FULL;
    c.flags = CF_OBSCURED_ALL;
}

static void set_all_outward(Object& obj, uint32_t off, int dim, int side, bool add) {
    Voxel* v = obj.chunk_voxels(off);
    uint8_t flag = adj_flag_for_face(dim, side);
    for_each_face_voxel(dim, side, [&](int i, int j, int k) {
        if (add)
            v[vidx(i, j, k)].flags |= flag;
        else
            v[vidx(i, j, k)].flags &= (uint8_t)~flag;
    });
}

static void update_outward_with_non_uniform(Object& obj, uint32_t cur_off, uint32_t adj_off, int dim,
                                            int side) {
    Voxel* cur = obj.chunk_voxels(cur_off);
    const Voxel* adj = obj.chunk_voxels(adj_off);
    uint8_t flag = adj_flag_for_face(dim, side);
    for_each_face_voxel(dim, side, [&](int i, int j, int k) {
        int a[3] = {i, j, k};
        a[dim] = side == 0 ? 15 : 0;
        Voxel& cv = cur[vidx(i, j, k)];
        if (!(cv.flags & FLAG_EMPTY)) {
            if (adj[vidx(a[0], a[1], a[2])].flags & FLAG_EMPTY)
                cv.flags &= (uint8_t)~flag;
            else
                cv.flags |= flag;
        }
    });
}

static inline void mark_obscured(Chunk& c, int dim, int side, bool obscured) {
    if (c.kind != CK_NONUNIFORM) {
        assert(!(c.kind == CK_UNIFORM && !obscured));
        return;
    }
    uint8_t bit = (uint8_t)(1u << (side == 0 ? dim : 3 + dim));
    if (obscured)
        c.flags |= bit;
    else
        c.flags &= (uint8_t)~bit;
}

// update_mutual_face_adjacencies; lower / upper = -1 means "outside the grid" (Void).
static void update_mutual_face_adjacencies(Object& obj, int64_t lower, int64_t upper, int dim) {
    Chunk lc = lower >= 0 ? obj.chunks[lower] : Chunk{};
    Chunk uc = upper >= 0 ? obj.chunks[upper] : Chunk{};
    const int LO = 0, UP = 1;
    if (lc.kind == CK_VOID && uc.kind == CK_VOID) return;
    if (lc.kind == CK_UNIFORM && uc.kind == CK_UNIFORM) return;
    if (lc.kind == CK_UNIFORM && uc.kind == CK_VOID) {
        Chunk& c = obj.chunks[lower];
        convert_to_non_uniform_if_uniform(obj, c);
        set_all_outward(obj, c.data_offset, dim, UP, false);
        mark_obscured(c, dim, UP, false);
        return;
    }
    if (lc.kind == CK_VOID && uc.kind == CK_UNIFORM) {
        Chunk& c = obj.chunks[upper];
        convert_to_non_uniform_if_uniform(obj, c);
        set_all_outward(obj, c.data_offset, dim, LO, false);
        mark_obscured(c, dim, LO, false);
        return;
    }
    if (lc.kind == CK_NONUNIFORM && uc.kind == CK_VOID) {
        if (lc.face[dim][1] != FD_EMPTY) set_all_outward(obj, lc.data_offset, dim, UP, false);
        mark_obscured(obj.chunks[lower], dim, UP, false);
        return;
    }
    if (lc.kind == CK_VOID && uc.kind == CK_NONUNIFORM) {
        if (uc.face[dim][0] != FD_EMPTY) set_all_outward(obj, uc.data_offset, dim, LO, false);
        mark_obscured(obj.chunks[upper], dim, LO, false);
        return;
    }
    if (lc.kind == CK_NONUNIFORM && uc.kind == CK_UNIFORM) {
        uint8_t fd = lc.face[dim][1];
        if (fd != FD_EMPTY) set_all_outward(obj, lc.data_offset, dim, UP, true);
        mark_obscured(obj.chunks[lower], dim, UP, true);
        if (fd == FD_EMPTY) {
            Chunk& c = obj.chunks[upper];
            convert_to_non_uniform_if_uniform(obj, c);
            set_all_outward(obj, c.data_offset, dim, LO, false);
            mark_obscured(c, dim, LO, false);
        } else if (fd == FD_MIXED) {
            Chunk& c = obj.chunks[upper];
            convert_to_non_uniform_if_uniform(obj, c);
            update_outward_with_non_uniform(obj, c.data_offset, lc.data_offset, dim, LO);
            mark_obscured(c, dim, LO, false);
        }
        return;
    }
    if (lc.kind == CK_UNIFORM && uc.kind == CK_NONUNIFORM) {
        uint8_t fd = uc.face[dim][0];
        if (fd != FD_EMPTY) set_all_outward(obj, uc.data_offset, dim, LO, true);
        mark_obscured(obj.chunks[upper], dim, LO, true);
        if (fd == FD_EMPTY) {
            Chunk& c = obj.chunks[lower];
            convert_to_non_uniform_if_uniform(obj, c);
            set_all_outward(obj, c.data_offset, dim, UP, false);
            mark_obscured(c, dim, UP, false);
        } else if (fd == FD_MIXED) {
            Chunk& c = obj.chunks[lower];
            convert_to_non_uniform_if_uniform(obj, c);
            update_outward_with_non_uniform(obj, c.data_offset, uc.data_offset, dim, UP);
            mark_obscured(c, dim, UP, false);
        }
        return;
    }
    // both non-uniform
    uint8_t lfd = lc.face[dim][1], ufd = uc.face[dim][0];
    if (lfd != FD_EMPTY) {
        if (ufd == FD_EMPTY)
            set_all_outward(obj, lc.data_offset, dim, UP, false);
        else if (ufd == FD_FULL)
            set_all_outward(obj, lc.data_offset, dim, UP, true);
        else
            update_outward_with_non_uniform(obj, lc.data_offset, uc.data_offset, dim, UP);
    }
    if (ufd != FD_EMPTY) {
        if (lfd == FD_EMPTY)
            set_all_outward(obj, uc.data_offset, dim, LO, false);
        else if (lfd == FD_FULL)
            set_all_outward(obj, uc.data_offset, dim, LO, true);
        else
            update_outward_with_non_uniform(obj, uc.data_offset, lc.data_offset, dim, LO);
    }
    mark_obscured(obj.chunks[lower], dim, UP, ufd == FD_FULL);
    mark_obscured(obj.chunks[upper], dim, LO, lfd == FD_FULL);
}

void update_upper_boundary_adjacencies_in_ranges(Object& obj, const uint32_t r[3][2]) {
    for (uint32_t i = r[0][0]; i < r[0][1]; ++i)
        for (uint32_t j = r[1][0]; j < r[1][1]; ++j)
            for (uint32_t k = r[2][0]; k < r[2][1]; ++k) {
                uint32_t ci = obj.lin(i, j, k);
                uint32_t adj[3][3] = {{i + 1, j, k}, {i, j + 1, k}, {i, j, k + 1}};
                for (int d = 0; d < 3; ++d) {
                    int64_t up = -1;
                    if (adj[d][d] < obj.chunk_counts[d]) up = obj.lin(adj[d][0], adj[d][1], adj[d][2]);
                    update_mutual_face_adjacencies(obj, ci, up, d);
                }
            }
}

void update_all_chunk_boundary_adjacencies(Object& obj) {
    uint32_t r[3][2] = {{0, obj.chunk_counts[0]}, {0, obj.chunk_counts[1]}, {0, obj.chunk_counts[2]}};
    update_upper_boundary_adjacencies_in_ranges(obj, r);
    for (uint32_t j = 0; j < obj.chunk_counts[1]; ++j)
        for (uint32_t k = 0; k < obj.chunk_counts[2]; ++k)
            update_mutual_face_adjacencies(obj, -1, obj.lin(0, j, k), 0);
    for (uint32_t i = 0; i < obj.chunk_counts[0]; ++i)
        for (uint32_t k = 0; k < obj.chunk_counts[2]; ++k)
            update_mutual_face_adjacencies(obj, -1, obj.lin(i, 0, k), 1);
    for (uint32_t i = 0; i < obj.chunk_counts[0]; ++i)
        for (uint32_t j = 0; j < obj.chunk_counts[1]; ++j)
            update_mutual_face_adjacencies(obj, -1, obj.lin(i, j, 0), 2);
}

void compute_all_derived_state(Object& obj) {
    for (const Chunk& c : obj.chunks)
        if (c.kind == CK_NONUNIFORM) update_internal_adjacencies(obj.chunk_voxels(c.data_offset));
    update_all_chunk_boundary_adjacencies(obj);
}

void generate_object(const VoxelGenerator& vg, Object& obj, int n_threads, double* t_gen,
                     double* t_der) {
    auto t0 = std::chrono::steady_clock::now();
    generate_without_derived_state(vg, obj, n_threads);
    auto t1 = std::chrono::steady_clock::now();
    update_occupied_voxel_ranges(obj);
    compute_all_derived_state(obj);
    auto t2 = std::chrono::steady_clock::now();
    if (t_gen) *t_gen = std::chrono::duration<double>(t1 - t0).count();
    if (t_der) *t_der = std::chrono::duration<double>(t2 - t1).count();
}

}  // namespace orc
