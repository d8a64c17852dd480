// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
//
// Restates (V = engine/crates/impact_voxel/src):
//   VoxelObject::modify_voxels_within_sphere       V/object/intersection.rs:283-394
//   VoxelObject::modify_voxels_within_capsule      V/object/intersection.rs:417-537
//   Capsule::compute_aabb / trim_segment_outside_aab / CapsulePointContainmentTester
//                                                  impact_geometry/src/capsule.rs:132-250
//   AxisAlignedBox::find_contained_subsegment      impact_geometry/src/axis_aligned_box.rs:385-415
//   VoxelAbsorbingCapsule::compute_new_signed_distance, apply_capsule_absorption closure
//                                                  V/interaction/absorption.rs:226-237, 869-888
//   handle_chunk_voxels_modified                   V/object/intersection.rs:539-598
//   voxel_ranges_touching_aab                      V/object/intersection.rs:766-784
//   VoxelAbsorbingSphere::compute_new_signed_distance  V/interaction/absorption.rs:170-179
//   apply_sphere_absorption closure                V/interaction/absorption.rs:823-843
//   Voxel::set_signed_distance                     V/lib.rs:451-461
#include <algorithm>
#include <cmath>

#include "oracle.hpp"

namespace orc {

static inline int vidx(int i, int j, int k) { return (i << 8) + (j << 4) + k; }

namespace {

struct Shape {
    bool capsule;
    V3 start;    // sphere centre / capsule segment start
    V3 vec;      // capsule segment vector
    float radius, influence_radius;
};

inline float cmp3(const V3& v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }

// Sphere::compute_aabb (sphere.rs:245-248); Capsule::compute_aabb (capsule.rs:132-137)
void shape_aabb(const Shape& s, V3 start, V3 vec, float lo[3], float hi[3]) {
    const float R = s.influence_radius;
    V3 end = start + vec;
    for (int d = 0; d < 3; ++d) {
        const float a = cmp3(start, d);
        lo[d] = a - R;
        hi[d] = a + R;
        if (s.capsule) {
            const float b = cmp3(end, d);
            lo[d] = std::fmin(lo[d], b - R);
            hi[d] = std::fmax(hi[d], b + R);
        }
    }
}

// voxel_ranges_touching_aab (intersection.rs:766-782); `as usize` saturates
void ranges_touching(const uint32_t max_r[3][2], const float lo[3], const float hi[3], uint32_t out[3][2]) {
    for (int d = 0; d < 3; ++d) {
        float fl = std::fmax(std::floor(lo[d]), 0.0f);
        float ce = std::ceil(hi[d]);
        uint32_t s = fl >= 4294967296.0f ? UINT32_MAX : (uint32_t)fl;
        uint32_t e = !(ce > 0.0f) ? 0u : (ce >= 4294967296.0f ? UINT32_MAX : (uint32_t)ce);
        out[d][0] = std::max(max_r[d][0], s);
        out[d][1] = std::min(max_r[d][1], e);
    }
}

// Capsule::trim_segment_outside_aab against the chunk's box (capsule.rs:144-165)
bool trim_capsule(const Shape& s, const uint32_t cc[3], V3& t_start, V3& t_vec) {
    float t_min = 0.0f, t_max = 1.0f;
    for (int d = 0; d < 3; ++d) {
        const float lo = (float)(cc[d] * 16u) - s.influence_radius;
        const float hi = (float)((cc[d] + 1u) * 16u) + s.influence_radius;
        const float v = cmp3(s.vec, d), o = cmp3(s.start, d);
        if (std::fabs(v) > 1e-8f) {
            const float recip = 1.0f / v;
            const float t1 = (lo - o) * recip, t2 = (hi - o) * recip;
            const float t_entry = t1 < t2 ? t1 : t2, t_exit = t1 < t2 ? t2 : t1;
            t_min = std::fmax(t_min, t_entry);
            t_max = std::fmin(t_max, t_exit);
        } else if (o < lo || o > hi) {
            return false;
        }
    }
    if (!(t_min <= t_max)) return false;
    t_start = s.start + t_min * s.vec;
    t_vec = (t_max - t_min) * s.vec;
    return true;
}

// handle_chunk_voxels_modified (intersection.rs:539-598)
void handle_chunk_voxels_modified(Object& obj, Chunk& chunk, Voxel* v, const uint32_t cc[3], uint32_t cidx,
                                  const uint32_t vr[3][2], const uint32_t tv[3][2], AbsorbStats& st, bool& removed_chunks) {
    st.touched_chunks++;
    Sparseness sp = update_all_internal_state(chunk, v);
    if (sp.is_void) {
        chunk = Chunk{};
        removed_chunks = true;
        st.removed_chunks++;
    }
    obj.mark_dirty(cidx);
    for (int d = 0; d < 3; ++d) {
        if (cc[d] > 0 && tv[d][0] - vr[d][0] < 2) {
            uint32_t a[3] = {cc[0], cc[1], cc[2]};
            a[d] -= 1;
            obj.mark_dirty(obj.lin(a[0], a[1], a[2]));
        }
        if (cc[d] + 1 < obj.chunk_counts[d] && vr[d][1] - tv[d][1] < 2) {
            uint32_t a[3] = {cc[0], cc[1], cc[2]};
            a[d] += 1;
            obj.mark_dirty(obj.lin(a[0], a[1], a[2]));
        }
    }
}

void absorb_shape(Object& obj, const Shape& shape, AbsorbStats* stats, InertialUpdater* updater) {
    AbsorbStats st{0, 0, 0, 0};
    const V3 center = shape.start;
    const float radius = shape.radius, influence_radius = shape.influence_radius;
    float lo[3], hi[3];
    shape_aabb(shape, shape.start, shape.vec, lo, hi);
    // CapsulePointContainmentTester (capsule.rs:168-181)
    const float len2 = dot(shape.vec, shape.vec);
    const V3 vec_over_len2 = len2 > 1e-8f ? v3(shape.vec.x / len2, shape.vec.y / len2, shape.vec.z / len2)
                                          : v3(0.0f, 0.0f, 0.0f);
    uint32_t tr[3][2];
    bool empty = false;
    for (int d = 0; d < 3; ++d) {
        float fl = std::fmax(std::floor(lo[d]), 0.0f);
        float ce = std::ceil(hi[d]);
        // `as usize` saturating casts
        uint32_t s = fl >= 4294967296.0f ? UINT32_MAX : (uint32_t)fl;
        uint32_t e = !(ce > 0.0f) ? 0u : (ce >= 4294967296.0f ? UINT32_MAX : (uint32_t)ce);
        tr[d][0] = std::max(obj.occ_voxels[d][0], s);
        tr[d][1] = std::min(obj.occ_voxels[d][1], e);
        if (tr[d][0] >= tr[d][1]) empty = true;
    }
    if (empty) {
        if (stats) *stats = st;
        return;
    }
    uint32_t cr[3][2];
    for (int d = 0; d < 3; ++d) {
        cr[d][0] = tr[d][0] / 16;
        cr[d][1] = (tr[d][1] + 15) / 16;
    }
    const float r2 = influence_radius * influence_radius;
    bool removed_chunks = false;

    for (uint32_t ci = cr[0][0]; ci < cr[0][1]; ++ci)
        for (uint32_t cj = cr[1][0]; cj < cr[1][1]; ++cj)
            for (uint32_t ck = cr[2][0]; ck < cr[2][1]; ++ck) {
                uint32_t cidx = obj.lin(ci, cj, ck);
                Chunk& chunk = obj.chunks[cidx];
                uint32_t cc[3] = {ci, cj, ck};
                uint32_t vr[3][2], tv[3][2];
                for (int d = 0; d < 3; ++d) {
                    vr[d][0] = cc[d] * 16;
                    vr[d][1] = (cc[d] + 1) * 16;
                }
                if (shape.capsule) {
                    // the capsule trimmed to the chunk decides the touched voxels, before any conversion
                    // (intersection.rs:445-461)
                    V3 ts, tvec;
                    if (!trim_capsule(shape, cc, ts, tvec)) continue;
                    float tlo[3], thi[3];
                    shape_aabb(shape, ts, tvec, tlo, thi);
                    ranges_touching(vr, tlo, thi, tv);
                    if (tv[0][0] >= tv[0][1] || tv[1][0] >= tv[1][1] || tv[2][0] >= tv[2][1]) continue;
                }
                if (chunk.kind == CK_VOID) continue;
                if (chunk.kind == CK_UNIFORM) {
                    size_t start = obj.voxels.size();
                    obj.voxels.resize(start + CHUNK_VOXELS, chunk.uniform_voxel);
                    chunk.kind = CK_NONUNIFORM;
                    chunk.data_offset = (uint32_t)(start >> 12);
                    for (int d = 0; d < 3; ++d) chunk.face[d][0] = chunk.face[d][1] = FD_FULL;
                    chunk.flags = CF_OBSCURED_ALL;
                }
                if (!shape.capsule) {
                    for (int d = 0; d < 3; ++d) {
                        tv[d][0] = std::max(vr[d][0], tr[d][0]);
                        tv[d][1] = std::min(vr[d][1], tr[d][1]);
                    }
                }
                Voxel* v = obj.chunk_voxels(chunk.data_offset);
                bool touched = false;
                for (uint32_t i = tv[0][0]; i < tv[0][1]; ++i)
                    for (uint32_t j = tv[1][0]; j < tv[1][1]; ++j)
                        for (uint32_t k = tv[2][0]; k < tv[2][1]; ++k) {
                            V3 p = v3((float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f);
                            float d2;
                            bool inside;
                            if (shape.capsule) {
                                // shortest_squared_distance_from_point_to_segment (capsule.rs:240-250); boundary included
                                V3 sp = p - shape.start;
                                float t = dot(sp, vec_over_len2);
                                if (t < 0.0f) t = 0.0f;
                                if (t > 1.0f) t = 1.0f;
                                V3 closest = shape.start + t * shape.vec;
                                V3 df = p - closest;
                                d2 = dot(df, df);
                                inside = d2 <= r2;
                            } else {
                                V3 df = center - p;
                                d2 = dot(df, df);
                                inside = d2 < r2;
                            }
                            if (inside) {
                                Voxel& vx = v[vidx(i & 15, j & 15, k & 15)];
                                bool was_empty = vx.flags & FLAG_EMPTY;
                                float sphere_sd = std::sqrt(d2) - radius;
                                float nsd = std::fmax(sd_decode(vx.sd), -sphere_sd);
                                vx.sd = sd_encode(nsd);
                                if (!(vx.sd < 0)) {
                                    vx.flags |= FLAG_EMPTY;
                                    if (!was_empty) {
                                        st.emptied_voxels++;
                                        // the closure's callback (absorption.rs:836-840): remove_voxel with the
                                        // object voxel indices and the voxel's (unchanged) type
                                        if (updater) {
                                            const uint32_t ijk[3] = {i, j, k};
                                            updater->remove_voxel(ijk, vx.type);
                                        }
                                    }
                                }
                                st.touched_voxels++;
                                touched = true;
                            }
                        }
                if (touched) handle_chunk_voxels_modified(obj, chunk, v, cc, cidx, vr, tv, st, removed_chunks);
            }
    if (removed_chunks) {
        update_occupied_chunk_ranges(obj);
        update_occupied_voxel_ranges(obj);
    }
    uint32_t br[3][2];
    for (int d = 0; d < 3; ++d) {
        br[d][0] = cr[d][0] > 0 ? cr[d][0] - 1 : 0;
        br[d][1] = cr[d][1];
    }
    update_upper_boundary_adjacencies_in_ranges(obj, br);
    if (stats) *stats = st;
}

}  // namespace

void absorb_sphere(Object& obj, V3 center, float radius, float influence_radius, AbsorbStats* stats,
                   InertialUpdater* updater) {
    absorb_shape(obj, Shape{false, center, v3(0.0f, 0.0f, 0.0f), radius, influence_radius}, stats, updater);
}

void absorb_capsule(Object& obj, V3 segment_start, V3 segment_vector, float radius, float influence_radius,
                    AbsorbStats* stats, InertialUpdater* updater) {
    absorb_shape(obj, Shape{true, segment_start, segment_vector, radius, influence_radius}, stats, updater);
}

// ---- mutual absorption between two voxel objects ---------------------------------------------------------------
//   VoxelObject::modify_voxels_within_ranges          V/object/intersection.rs:167-261
//   VoxelObject::voxel                                V/object.rs:1050-1062
//   sample_voxel_object_sdf, evaluate_sdf_from_corner_samples   V/object/sdf.rs:579-592, 636-675
//   apply_mutual_absorption, compute_subtracted_signed_distance V/interaction/absorption.rs:891-1094
//   Isometry3::transform_point / inverse_transform_point        impact_math/src/transform/isometry.rs:146-172
namespace {

template <typename F>
void modify_voxels_within_ranges(Object& obj, const uint32_t r[3][2], AbsorbStats& st, F&& modify_voxel) {
    for (int d = 0; d < 3; ++d)
        if (r[d][0] >= r[d][1]) return;
    uint32_t cr[3][2];
    for (int d = 0; d < 3; ++d) {
        cr[d][0] = r[d][0] / 16;
        cr[d][1] = (r[d][1] + 15) / 16;
    }
    bool removed_chunks = false;
    for (uint32_t ci = cr[0][0]; ci < cr[0][1]; ++ci)
        for (uint32_t cj = cr[1][0]; cj < cr[1][1]; ++cj)
            for (uint32_t ck = cr[2][0]; ck < cr[2][1]; ++ck) {
                const uint32_t cidx = obj.lin(ci, cj, ck);
                Chunk& chunk = obj.chunks[cidx];
                if (chunk.kind == CK_VOID) continue;
                if (chunk.kind == CK_UNIFORM) {
                    size_t start = obj.voxels.size();
                    obj.voxels.resize(start + CHUNK_VOXELS, chunk.uniform_voxel);
                    chunk.kind = CK_NONUNIFORM;
                    chunk.data_offset = (uint32_t)(start >> 12);
                    for (int d = 0; d < 3; ++d) chunk.face[d][0] = chunk.face[d][1] = FD_FULL;
                    chunk.flags = CF_OBSCURED_ALL;
                }
                const uint32_t cc[3] = {ci, cj, ck};
                uint32_t vr[3][2], tv[3][2];
                for (int d = 0; d < 3; ++d) {
                    vr[d][0] = cc[d] * 16;
                    vr[d][1] = (cc[d] + 1) * 16;
                    tv[d][0] = std::max(vr[d][0], r[d][0]);
                    tv[d][1] = std::min(vr[d][1], r[d][1]);
                }
                Voxel* v = obj.chunk_voxels(chunk.data_offset);
                bool touched = false;
                for (uint32_t i = tv[0][0]; i < tv[0][1]; ++i)
                    for (uint32_t j = tv[1][0]; j < tv[1][1]; ++j)
                        for (uint32_t k = tv[2][0]; k < tv[2][1]; ++k) {
                            const uint32_t ijk[3] = {i, j, k};
                            if (modify_voxel(ijk, v[vidx(i & 15, j & 15, k & 15)])) {
                                touched = true;
                                st.touched_voxels++;
                            }
                        }
                if (touched) handle_chunk_voxels_modified(obj, chunk, v, cc, cidx, vr, tv, st, removed_chunks);
            }
    if (removed_chunks) {
        update_occupied_chunk_ranges(obj);
        update_occupied_voxel_ranges(obj);
    }
    uint32_t br[3][2];
    for (int d = 0; d < 3; ++d) {
        br[d][0] = cr[d][0] > 0 ? cr[d][0] - 1 : 0;
        br[d][1] = cr[d][1];
    }
    update_upper_boundary_adjacencies_in_ranges(obj, br);
}

inline float voxel_sd(const Object& obj, uint32_t i, uint32_t j, uint32_t k) {
    const Chunk& c = obj.chunks[obj.lin(i >> 4, j >> 4, k >> 4)];
    if (c.kind == CK_VOID) return sd_decode(127);
    if (c.kind == CK_UNIFORM) return sd_decode(c.uniform_voxel.sd);
    return sd_decode(obj.chunk_voxels(c.data_offset)[vidx(i & 15, j & 15, k & 15)].sd);
}

inline float corner_samples(const float d[8], V3 o) {
    const V3 ro = v3(1.0f - o.x, 1.0f - o.y, 1.0f - o.z);
    const float d00 = d[0] * ro.x + d[4] * o.x;
    const float d01 = d[1] * ro.x + d[5] * o.x;
    const float d10 = d[2] * ro.x + d[6] * o.x;
    const float d11 = d[3] * ro.x + d[7] * o.x;
    const float d0 = d00 * ro.y + d10 * o.y;
    const float d1 = d01 * ro.y + d11 * o.y;
    return d0 * ro.z + d1 * o.z;
}

// `idx as isize` of a floored float (saturating) kept in 64 bits
inline int64_t floor_index(float f) {
    if (f != f) return 0;
    if (f >= 9.2e18f) return INT64_MAX;
    if (f <= -9.2e18f) return INT64_MIN;
    return (int64_t)f;
}

float sample_object_sdf(const Object& obj, const uint32_t dims[3], V3 p) {
    const V3 lc = v3(p.x - 0.5f, p.y - 0.5f, p.z - 0.5f);
    const V3 li = v3(std::floor(lc.x), std::floor(lc.y), std::floor(lc.z));
    const V3 fo = lc - li;
    if (sign_neg(li.x) || sign_neg(li.y) || sign_neg(li.z)) return SD_MAX_F32;  // has_negative_component: sign bits
    const uint64_t i = (uint64_t)floor_index(li.x), j = (uint64_t)floor_index(li.y), k = (uint64_t)floor_index(li.z);
    if (i + 1 >= dims[0] || j + 1 >= dims[1] || k + 1 >= dims[2]) return SD_MAX_F32;
    const uint32_t a = (uint32_t)i, b = (uint32_t)j, c = (uint32_t)k;
    const float d[8] = {voxel_sd(obj, a, b, c),         voxel_sd(obj, a, b, c + 1),     voxel_sd(obj, a, b + 1, c),
                        voxel_sd(obj, a, b + 1, c + 1), voxel_sd(obj, a + 1, b, c),     voxel_sd(obj, a + 1, b, c + 1),
                        voxel_sd(obj, a + 1, b + 1, c), voxel_sd(obj, a + 1, b + 1, c + 1)};
    return corner_samples(d, fo);
}

inline float subtracted_sd(float sd, float inside_other, float k, float qik) {
    const float inter = std::fmax(sd, inside_other);
    if (k == 0.0f) return std::fmax(sd, -inter);
    // -smooth_sdf_union(-d1, d2)
    const float d1 = -sd, d2 = inter;
    const float h = std::fmax(k - std::fabs(d1 - d2), 0.0f);
    return -(std::fmin(d1, d2) - (h * h) * qik);
}

// Voxel::set_signed_distance + the closures' callback; true if the voxel went from non-empty to empty
inline void set_sd(Voxel& vx, float nsd, const uint32_t ijk[3], InertialUpdater* upd, AbsorbStats& st) {
    const bool was_empty = vx.flags & FLAG_EMPTY;
    vx.sd = sd_encode(nsd);
    if (!(vx.sd < 0)) {
        vx.flags |= FLAG_EMPTY;
        if (!was_empty) {
            st.emptied_voxels++;
            if (upd) upd->remove_voxel(ijk, vx.type);
        }
    }
}

}  // namespace

void absorb_mutually(Object& a, Object& b, Isometry b_to_a, float smoothness, const uint32_t ranges_a[3][2],
                     const uint32_t ranges_b[3][2], InertialUpdater* upd_a, InertialUpdater* upd_b, AbsorbStats* stats_a,
                     AbsorbStats* stats_b) {
    AbsorbStats sa{0, 0, 0, 0}, sb{0, 0, 0, 0};
    const float ea = a.voxel_extent, eb = b.voxel_extent;
    const float inv_ea = 1.0f / ea, inv_eb = 1.0f / eb;  // VoxelObject::inverse_voxel_extent = voxel_extent.recip()
    const float b_dist_to_a = eb * inv_ea, a_dist_to_b = ea * inv_eb;
    const float qik = 0.25f / smoothness;
    uint32_t dims_a[3], dims_b[3];
    for (int d = 0; d < 3; ++d) {
        dims_a[d] = a.chunk_counts[d] * 16;
        dims_b[d] = b.chunk_counts[d] * 16;
    }
    const float pad_f = std::ceil(b_dist_to_a);
    const uint32_t pad = !(pad_f > 0.0f) ? 0u : (pad_f >= 4294967296.0f ? UINT32_MAX : (uint32_t)pad_f);
    uint32_t sr[3][2];
    size_t n_snap = 1;
    for (int d = 0; d < 3; ++d) {
        sr[d][0] = ranges_a[d][0] > pad ? ranges_a[d][0] - pad : 0;
        sr[d][1] = std::min<uint64_t>((uint64_t)ranges_a[d][1] + pad, dims_a[d]);
        n_snap *= sr[d][1] > sr[d][0] ? sr[d][1] - sr[d][0] : 0;
    }
    std::vector<float> snapshot(n_snap, SD_MAX_F32);
    const size_t sj = sr[1][1] > sr[1][0] ? sr[1][1] - sr[1][0] : 0, sk = sr[2][1] > sr[2][0] ? sr[2][1] - sr[2][0] : 0;
    auto snap_idx = [&](uint32_t i, uint32_t j, uint32_t k) {
        return ((size_t)(i - sr[0][0]) * sj + (j - sr[1][0])) * sk + (k - sr[2][0]);
    };
    const Quat q = b_to_a.q, qc = quat_conj(b_to_a.q);
    const V3 t = b_to_a.t;

    modify_voxels_within_ranges(a, sr, sa, [&](const uint32_t ijk[3], Voxel& vx) {
        if (vx.sd == 127) return false;
        const float sd_a = sd_decode(vx.sd);
        snapshot[snap_idx(ijk[0], ijk[1], ijk[2])] = sd_a;
        const V3 center_a = v3(((float)ijk[0] + 0.5f) * ea, ((float)ijk[1] + 0.5f) * ea, ((float)ijk[2] + 0.5f) * ea);
        const V3 center_b = inv_eb * quat_rotate(qc, center_a - t);
        const float inside_b = sample_object_sdf(b, dims_b, center_b) * b_dist_to_a;
        set_sd(vx, subtracted_sd(sd_a, inside_b, smoothness, qik), ijk, upd_a, sa);
        return true;
    });

    modify_voxels_within_ranges(b, ranges_b, sb, [&](const uint32_t ijk[3], Voxel& vx) {
        if (vx.sd == 127) return false;
        const V3 center_b = v3(((float)ijk[0] + 0.5f) * eb, ((float)ijk[1] + 0.5f) * eb, ((float)ijk[2] + 0.5f) * eb);
        const V3 center_a = inv_ea * (quat_rotate(q, center_b) + t);
        const V3 lc = v3(center_a.x - 0.5f, center_a.y - 0.5f, center_a.z - 0.5f);
        const V3 lf = v3(std::floor(lc.x), std::floor(lc.y), std::floor(lc.z));
        const V3 fo = lc - lf;
        const int64_t li[3] = {floor_index(lf.x), floor_index(lf.y), floor_index(lf.z)};
        for (int d = 0; d < 3; ++d)
            if (li[d] < (int64_t)sr[d][0] || li[d] + 1 >= (int64_t)sr[d][1]) return false;
        const uint32_t i = (uint32_t)li[0], j = (uint32_t)li[1], k = (uint32_t)li[2];
        const float dd[8] = {snapshot[snap_idx(i, j, k)],         snapshot[snap_idx(i, j, k + 1)],
                             snapshot[snap_idx(i, j + 1, k)],     snapshot[snap_idx(i, j + 1, k + 1)],
                             snapshot[snap_idx(i + 1, j, k)],     snapshot[snap_idx(i + 1, j, k + 1)],
                             snapshot[snap_idx(i + 1, j + 1, k)], snapshot[snap_idx(i + 1, j + 1, k + 1)]};
        const float inside_a = corner_samples(dd, fo) * a_dist_to_b;
        set_sd(vx, subtracted_sd(sd_decode(vx.sd), inside_a, smoothness, qik), ijk, upd_b, sb);
        return true;
    });
    if (stats_a) *stats_a = sa;
    if (stats_b) *stats_b = sb;
}

// ---- surface voxel queries (object/intersection.rs:51-151) ---------------------------------------------------------
void surface_voxels_in_ranges(const Object& obj, const uint32_t r[3][2], std::vector<SurfaceVoxel>& out) {
    out.clear();
    for (int d = 0; d < 3; ++d)
        if (r[d][0] >= r[d][1]) return;
    for (uint32_t ci = r[0][0] / 16; ci < (r[0][1] + 15) / 16; ++ci)
        for (uint32_t cj = r[1][0] / 16; cj < (r[1][1] + 15) / 16; ++cj)
            for (uint32_t ck = r[2][0] / 16; ck < (r[2][1] + 15) / 16; ++ck) {
                const Chunk& c = obj.chunks[obj.lin(ci, cj, ck)];
                if (c.kind != CK_NONUNIFORM) continue;  // only non-uniform chunks can have surface voxels
                const Voxel* v = obj.chunk_voxels(c.data_offset);
                const uint32_t cc[3] = {ci, cj, ck};
                uint32_t t0[3], t1[3];
                for (int d = 0; d < 3; ++d) {
                    t0[d] = std::max(cc[d] * 16, r[d][0]);
                    t1[d] = std::min((cc[d] + 1) * 16, r[d][1]);
                }
                for (uint32_t i = t0[0]; i < t1[0]; ++i)
                    for (uint32_t j = t0[1]; j < t1[1]; ++j)
                        for (uint32_t k = t0[2]; k < t1[2]; ++k) {
                            const Voxel& vx = v[vidx(i & 15, j & 15, k & 15)];
                            if (vx.flags & FLAG_EMPTY) continue;  // Voxel::placement: None
                            const int blocked = __builtin_popcount(vx.flags & FLAG_FULL_ADJ);
                            if (blocked == 6) continue;  // VoxelPlacement::Interior
                            out.push_back(SurfaceVoxel{{i, j, k}, vx, (uint8_t)(blocked == 5 ? 0 : (blocked == 4 ? 1 : 2))});
                        }
            }
}

// ---- sphere contacts (collidable.rs:1097-1127, 1453-1455; impact_physics sphere.rs:105-136) ------------------------
void sphere_voxel_object_contacts(const Object& obj, const Isometry& T, V3 center, float radius, std::vector<VoxelContact>& out) {
    out.clear();
    const float e = obj.voxel_extent, inv_e = 1.0f / e;
    // sphere.iso_transformed(transform_to_object_space), then .scaled(inverse_voxel_extent) and its box clipped to the
    // occupied ranges (for_each_surface_voxel_maybe_intersecting_sphere)
    const V3 c_obj = quat_rotate(T.q, center) + T.t;
    const V3 cn = inv_e * c_obj;
    const float rn = inv_e * radius;
    const float lo[3] = {cn.x - rn, cn.y - rn, cn.z - rn}, hi[3] = {cn.x + rn, cn.y + rn, cn.z + rn};
    uint32_t r[3][2];
    ranges_touching(obj.occ_voxels, lo, hi, r);
    std::vector<SurfaceVoxel> sv;
    surface_voxels_in_ranges(obj, r, sv);
    const Quat qc = quat_conj(T.q);
    for (const SurfaceVoxel& v : sv) {
        const V3 c_voxel = v3(((float)v.ijk[0] + 0.5f) * e, ((float)v.ijk[1] + 0.5f) * e, ((float)v.ijk[2] + 0.5f) * e);
        const V3 vc = quat_rotate(qc, c_voxel - T.t);           // inverse_transform_point
        const float vr = -sd_decode(v.voxel.sd) * e;            // compute_voxel_radius
        const V3 disp = center - vc;
        const float d2 = dot(disp, disp);
        const float max_d = radius + vr;
        if (d2 > max_d * max_d) continue;
        const float dist = std::sqrt(d2);
        const V3 n = dist > 1e-8f ? v3(disp.x / dist, disp.y / dist, disp.z / dist) : v3(0.0f, 0.0f, 1.0f);
        const V3 pos = vc + vr * n;
        out.push_back(VoxelContact{{v.ijk[0], v.ijk[1], v.ijk[2]}, {pos.x, pos.y, pos.z}, {n.x, n.y, n.z},
                                   std::fmax(0.0f, max_d - dist)});
    }
}

// ---- plane queries and contacts (object/intersection.rs:30-40, 751-761; collidable.rs:1176-1209) -------------------
void voxel_ranges_within_plane(const uint32_t occ[3][2], V3 n, float displacement, uint32_t out[3][2]) {
    const float c[2][3] = {{(float)occ[0][0], (float)occ[1][0], (float)occ[2][0]},
                           {(float)occ[0][1], (float)occ[1][1], (float)occ[2][1]}};
    float lo[3] = {c[0][0], c[0][1], c[0][2]}, hi[3] = {c[1][0], c[1][1], c[1][2]};
    const float nn[3] = {n.x, n.y, n.z};
    const int perm[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
    for (const auto& p : perm) {
        const int i = p[0], j = p[1], k = p[2];
        if (std::fabs(nn[k]) > 1e-8f) {
            const float a = nn[i] * c[0][i] + nn[j] * c[0][j];
            const float b = nn[i] * c[0][i] + nn[j] * c[1][j];
            const float cc = nn[i] * c[1][i] + nn[j] * c[0][j];
            const float d = nn[i] * c[1][i] + nn[j] * c[1][j];
            const float extremal = (displacement - std::fmin(std::fmin(std::fmin(a, b), cc), d)) / nn[k];
            if (!std::signbit(nn[k])) {
                lo[k] = std::fmin(lo[k], extremal);
                hi[k] = std::fmin(hi[k], extremal);
            } else {
                lo[k] = std::fmax(lo[k], extremal);
                hi[k] = std::fmax(hi[k], extremal);
            }
        }
    }
    ranges_touching(occ, lo, hi, out);
}

void plane_voxel_object_contacts(const Object& obj, const Isometry& T, V3 normal, float displacement,
                                 std::vector<VoxelContact>& out) {
    out.clear();
    const float e = obj.voxel_extent, inv_e = 1.0f / e;
    // plane.iso_transformed(transform_to_object_space) (plane.rs:197-203), then .scaled(inverse_voxel_extent)
    const V3 point_in_plane = normal * displacement;
    const V3 tp = quat_rotate(T.q, point_in_plane) + T.t;
    const V3 tn = quat_rotate(T.q, normal);
    const float td = dot(tn, tp);
    uint32_t r[3][2];
    voxel_ranges_within_plane(obj.occ_voxels, tn, td * inv_e, r);
    std::vector<SurfaceVoxel> sv;
    surface_voxels_in_ranges(obj, r, sv);
    const Quat qc = quat_conj(T.q);
    for (const SurfaceVoxel& v : sv) {
        if (v.placement != 2) continue;  // only the corner voxels matter against a plane
        const V3 c_voxel = v3(((float)v.ijk[0] + 0.5f) * e, ((float)v.ijk[1] + 0.5f) * e, ((float)v.ijk[2] + 0.5f) * e);
        const V3 vc = quat_rotate(qc, c_voxel - T.t);
        const float vr = -sd_decode(v.voxel.sd) * e;
        const float sd = dot(normal, vc) - displacement;
        const float depth = vr - sd;
        if (depth < 0.0f) continue;
        const V3 pos = vc - sd * normal;
        out.push_back(VoxelContact{{v.ijk[0], v.ijk[1], v.ijk[2]}, {pos.x, pos.y, pos.z}, {normal.x, normal.y, normal.z}, depth});
    }
}

// ---- capsule contacts (collidable.rs:1257-1288; impact_physics capsule.rs:212-270; impact_geometry line.rs:26-45) ------
void capsule_voxel_object_contacts(const Object& obj, const Isometry& T, V3 seg_start, V3 seg_vector, float radius,
                                   std::vector<VoxelContact>& out) {
    out.clear();
    const float e = obj.voxel_extent, inv_e = 1.0f / e;
    // capsule.iso_transformed(transform_to_object_space) (capsule.rs:122-128), .scaled(inverse_voxel_extent) (:100-106),
    // .compute_aabb() (:132-137) clipped to the occupied ranges (for_each_surface_voxel_maybe_intersecting_capsule)
    const V3 s_obj = quat_rotate(T.q, seg_start) + T.t;
    const V3 v_obj = quat_rotate(T.q, seg_vector);
    const V3 sn = inv_e * s_obj, vn = inv_e * v_obj;
    const V3 en = sn + vn;
    const float rn = inv_e * radius;
    const float lo[3] = {std::fmin(sn.x - rn, en.x - rn), std::fmin(sn.y - rn, en.y - rn), std::fmin(sn.z - rn, en.z - rn)};
    const float hi[3] = {std::fmax(sn.x + rn, en.x + rn), std::fmax(sn.y + rn, en.y + rn), std::fmax(sn.z + rn, en.z + rn)};
    uint32_t r[3][2];
    ranges_touching(obj.occ_voxels, lo, hi, r);
    std::vector<SurfaceVoxel> sv;
    surface_voxels_in_ranges(obj, r, sv);
    const Quat qc = quat_conj(T.q);
    const float len2 = dot(seg_vector, seg_vector);
    for (const SurfaceVoxel& v : sv) {
        const V3 c_voxel = v3(((float)v.ijk[0] + 0.5f) * e, ((float)v.ijk[1] + 0.5f) * e, ((float)v.ijk[2] + 0.5f) * e);
        const V3 vc = quat_rotate(qc, c_voxel - T.t);
        const float vr = -sd_decode(v.voxel.sd) * e;
        float param = 0.0f;
        if (!(len2 <= 1e-8f)) {
            const V3 sp = vc - seg_start;
            param = std::fmin(std::fmax(dot(seg_vector, sp) / len2, 0.0f), 1.0f);
        }
        const V3 closest = seg_start + param * seg_vector;
        const V3 sdisp = vc - closest;
        const float sd2 = dot(sdisp, sdisp);
        const float max_sd = vr + radius;
        if (sd2 > max_sd * max_sd) continue;
        const float sdist = std::sqrt(sd2);
        V3 cn;
        float depth;
        if (sdist > 1e-8f) {
            cn = v3(sdisp.x / sdist, sdisp.y / sdist, sdisp.z / sdist);
            depth = std::fmax(0.0f, max_sd - sdist);
        } else {
            // glam Vec3A::any_orthogonal_vector (third party): cross with Y when |x| > |y|, else with X
            const V3 o = std::fabs(seg_vector.x) > std::fabs(seg_vector.y) ? v3(-seg_vector.z, 0.0f, seg_vector.x)
                                                                          : v3(0.0f, seg_vector.z, -seg_vector.y);
            const float on = std::sqrt(dot(o, o));
            cn = on > 1e-8f ? v3(o.x / on, o.y / on, o.z / on) : v3(0.0f, 0.0f, 1.0f);
            depth = std::fmax(0.0f, max_sd);
        }
        const V3 n = v3(-cn.x, -cn.y, -cn.z);
        const V3 pos = vc + vr * n;
        out.push_back(VoxelContact{{v.ijk[0], v.ijk[1], v.ijk[2]}, {pos.x, pos.y, pos.z}, {n.x, n.y, n.z}, depth});
    }
}

}  // namespace orc
