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
                if (touched) {
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
                            uint32_t a[3] = {ci, cj, ck};
                            a[d] -= 1;
                            obj.mark_dirty(obj.lin(a[0], a[1], a[2]));
                        }
                        if (cc[d] + 1 < obj.chunk_counts[d] && vr[d][1] - tv[d][1] < 2) {
                            uint32_t a[3] = {ci, cj, ck};
                            a[d] += 1;
                            obj.mark_dirty(obj.lin(a[0], a[1], a[2]));
                        }
                    }
                }
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

}  // namespace orc
