// Host-side geometry of the mutual absorption path (no device work): which voxel ranges of two objects can overlap.
//
// Replaces
//   VoxelObject::determine_voxel_ranges_encompassing_intersection   (object/intersection.rs:707-745)
//   normalized_aabb_from_voxel_ranges, voxel_ranges_touching_aab    (object.rs:3311-3325, object/intersection.rs:766-782)
//   compute_box_intersection_bounds                                 (impact_geometry/src/oriented_box.rs:315-431)
//   OrientedBox::from_axis_aligned_box / iso_transformed / compute_corners / transform_point_{to,from}_box_frame
//                                                                   (oriented_box.rs:57-64, 149-214)
//   AxisAlignedBox::find_contained_subsegment / corner              (axis_aligned_box.rs:152-190, 385-415)
// f32 throughout, in the reference's operation order; quaternion rotation and the quaternion → axes conversion follow
// glam (Quat::mul_vec3a, Mat3A::from_quat), third party, restated like in program.cpp.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/impact_voxel_cuda.h"

namespace {

struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 l, V3 r) { return V3{l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x}; }
inline float comp(const V3& a, int d) { return d == 0 ? a.x : (d == 1 ? a.y : a.z); }
inline V3 vmin(V3 a, V3 b) { return V3{std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return V3{std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }

struct Quat {
    float x, y, z, w;
};
inline Quat conj(Quat q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
inline V3 rotate(Quat q, V3 v) {
    const V3 b{q.x, q.y, q.z};
    const float b2 = dot(b, b);
    return ((v * (q.w * q.w - b2)) + (b * (dot(v, b) * 2.0f))) + (cross(b, v) * (q.w * 2.0f));
}
inline void axes_of(Quat q, V3& ax, V3& ay, V3& az) {
    const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
    const float xx = q.x * x2, xy = q.x * y2, xz = q.x * z2;
    const float yy = q.y * y2, yz = q.y * z2, zz = q.z * z2;
    const float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
    ax = V3{1.0f - (yy + zz), xy + wz, xz - wy};
    ay = V3{xy - wz, 1.0f - (xx + zz), yz + wx};
    az = V3{xz + wy, yz - wx, 1.0f - (xx + yy)};
}

struct Aabb {
    V3 lo, hi;
};
struct Obb {
    V3 center;
    Quat q;
    V3 half;
};
inline V3 to_box_frame(const Obb& b, V3 p) { return rotate(conj(b.q), p - b.center); }
inline V3 from_box_frame(const Obb& b, V3 p) { return b.center + rotate(b.q, p); }

bool contained_subsegment(const Aabb& box, V3 start, V3 vec, float& t_min, float& t_max) {
    t_min = 0.0f;
    t_max = 1.0f;
    for (int d = 0; d < 3; ++d) {
        const float v = comp(vec, d), o = comp(start, d), lo = comp(box.lo, d), hi = comp(box.hi, d);
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
    return t_min <= t_max;
}

constexpr int EDGES[12][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

bool box_intersection_bounds(const Aabb& a, const Obb& b, Aabb& in_a, Aabb& in_b) {
    const float inf = INFINITY;
    in_a = Aabb{V3{inf, inf, inf}, V3{-inf, -inf, -inf}};
    in_b = in_a;
    bool intersect = false;
    auto expand = [&](V3 pa, V3 pb) {
        in_a.lo = vmin(in_a.lo, pa);
        in_a.hi = vmax(in_a.hi, pa);
        in_b.lo = vmin(in_b.lo, pb);
        in_b.hi = vmax(in_b.hi, pb);
        intersect = true;
    };
    // edges of box B cut to box A
    V3 ax, ay, az;
    axes_of(b.q, ax, ay, az);
    const V3 hw = b.half.x * ax, hh = b.half.y * ay, hd = b.half.z * az;
    const V3 bc[8] = {((b.center - hw) - hh) - hd, ((b.center - hw) - hh) + hd, ((b.center - hw) + hh) - hd,
                      ((b.center - hw) + hh) + hd, ((b.center + hw) - hh) - hd, ((b.center + hw) - hh) + hd,
                      ((b.center + hw) + hh) - hd, ((b.center + hw) + hh) + hd};
    for (const auto& e : EDGES) {
        const V3 s = bc[e[0]], v = bc[e[1]] - bc[e[0]];
        float t0, t1;
        if (contained_subsegment(a, s, v, t0, t1)) {
            const V3 p0 = s + v * t0, p1 = s + v * t1;
            expand(p0, to_box_frame(b, p0));
            expand(p1, to_box_frame(b, p1));
        }
    }
    // edges of box A, in B's frame, cut to box B
    V3 ac[8];
    for (int c = 0; c < 8; ++c)
        ac[c] = to_box_frame(b, V3{(c >> 2) & 1 ? a.hi.x : a.lo.x, (c >> 1) & 1 ? a.hi.y : a.lo.y, c & 1 ? a.hi.z : a.lo.z});
    const Aabb b_own{V3{-b.half.x, -b.half.y, -b.half.z}, b.half};
    for (const auto& e : EDGES) {
        const V3 s = ac[e[0]], v = ac[e[1]] - ac[e[0]];
        float t0, t1;
        if (contained_subsegment(b_own, s, v, t0, t1)) {
            const V3 p0 = s + v * t0, p1 = s + v * t1;
            expand(from_box_frame(b, p0), p0);
            expand(from_box_frame(b, p1), p1);
        }
    }
    return intersect;
}

// voxel_ranges_touching_aab: `as usize` casts saturate
void ranges_touching(const uint32_t occ[6], V3 lo, V3 hi, uint32_t out[6]) {
    for (int d = 0; d < 3; ++d) {
        const float fl = std::fmax(std::floor(comp(lo, d)), 0.0f), ce = std::ceil(comp(hi, d));
        const uint32_t s = fl >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)fl;
        const uint32_t e = !(ce > 0.0f) ? 0u : (ce >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)ce);
        out[2 * d] = occ[2 * d] > s ? occ[2 * d] : s;
        out[2 * d + 1] = occ[2 * d + 1] < e ? occ[2 * d + 1] : e;
    }
}

}  // namespace

extern "C" {

int ivx_box_intersection_bounds(const float a_lower[3], const float a_upper[3], const float b_center[3],
                                const float b_orientation[4], const float b_half_extents[3], float out_in_a[6],
                                float out_in_b[6], int* out_intersect) {
    if (!a_lower || !a_upper || !b_center || !b_orientation || !b_half_extents || !out_in_a || !out_in_b || !out_intersect)
        return IVX_ERR_INVALID_ARGUMENT;
    const Aabb a{V3{a_lower[0], a_lower[1], a_lower[2]}, V3{a_upper[0], a_upper[1], a_upper[2]}};
    const Obb b{V3{b_center[0], b_center[1], b_center[2]},
                Quat{b_orientation[0], b_orientation[1], b_orientation[2], b_orientation[3]},
                V3{b_half_extents[0], b_half_extents[1], b_half_extents[2]}};
    Aabb ia, ib;
    *out_intersect = box_intersection_bounds(a, b, ia, ib) ? 1 : 0;
    const float ra[6] = {ia.lo.x, ia.lo.y, ia.lo.z, ia.hi.x, ia.hi.y, ia.hi.z};
    const float rb[6] = {ib.lo.x, ib.lo.y, ib.lo.z, ib.hi.x, ib.hi.y, ib.hi.z};
    std::memcpy(out_in_a, ra, sizeof(ra));
    std::memcpy(out_in_b, rb, sizeof(rb));
    return IVX_OK;
}

int ivx_intersection_voxel_ranges(const uint32_t occupied_a[6], float voxel_extent_a, const uint32_t occupied_b[6],
                                  float voxel_extent_b, const ivx_isometry* transform_from_b_to_a, uint32_t out_ranges_in_a[6],
                                  uint32_t out_ranges_in_b[6], int* out_intersect) {
    if (!occupied_a || !occupied_b || !transform_from_b_to_a || !out_ranges_in_a || !out_ranges_in_b || !out_intersect)
        return IVX_ERR_INVALID_ARGUMENT;
    if (!(voxel_extent_a > 0.0f) || !(voxel_extent_b > 0.0f)) return IVX_ERR_INVALID_ARGUMENT;
    auto box_of = [](const uint32_t occ[6], float e) {
        return Aabb{e * V3{(float)occ[0], (float)occ[2], (float)occ[4]}, e * V3{(float)occ[1], (float)occ[3], (float)occ[5]}};
    };
    const Aabb a = box_of(occupied_a, voxel_extent_a), b = box_of(occupied_b, voxel_extent_b);
    const Quat q{transform_from_b_to_a->rotation[0], transform_from_b_to_a->rotation[1], transform_from_b_to_a->rotation[2],
                 transform_from_b_to_a->rotation[3]};
    const V3 t{transform_from_b_to_a->translation[0], transform_from_b_to_a->translation[1], transform_from_b_to_a->translation[2]};
    // OrientedBox::from_axis_aligned_box(b).iso_transformed(transform): the identity orientation times the rotation is the
    // rotation (glam's quaternion product with (0, 0, 0, 1) only adds zeros)
    const V3 b_center = 0.5f * (b.lo + b.hi);
    const Obb b_in_a{rotate(q, b_center) + t, q, 0.5f * (b.hi - b.lo)};
    Aabb in_a, in_b_rel;
    std::memset(out_ranges_in_a, 0, 6 * sizeof(uint32_t));
    std::memset(out_ranges_in_b, 0, 6 * sizeof(uint32_t));
    *out_intersect = 0;
    if (!box_intersection_bounds(a, b_in_a, in_a, in_b_rel)) return IVX_OK;
    *out_intersect = 1;
    // the second bounds are relative to B's centre: back to B's lower corner, then both to normalized voxel space
    const Aabb in_b{in_b_rel.lo + b_center, in_b_rel.hi + b_center};
    const float inv_a = 1.0f / voxel_extent_a, inv_b = 1.0f / voxel_extent_b;
    ranges_touching(occupied_a, inv_a * in_a.lo, inv_a * in_a.hi, out_ranges_in_a);
    ranges_touching(occupied_b, inv_b * in_b.lo, inv_b * in_b.hi, out_ranges_in_b);
    return IVX_OK;
}

}  // extern "C"
