// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the f32 vector / matrix / box arithmetic the reference's
// voxel hot path relies on. Nothing under oracle/ is product code: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference leg may load it. The product (impact_b200/csrc) never links or
// calls anything here.
//
// The reference wraps glam 0.30.10 (engine/Cargo.lock:1072-1073), which is not
// vendored under /root/reference; the operation ORDER below restates glam's
// SSE2 code paths as published (mul, then add, no FMA contraction). The oracle
// must be compiled with -ffp-contract=off.
//
//   Matrix4::transform_point  -> glam Mat4::transform_point3a
//        (impact_math/src/matrix.rs:678-682)
//   Matrix4::translate_transform / scale_transform (matrix.rs:652-676)
//   AxisAlignedBox::{aabb_of_transformed, box_lies_outside, contains_box,
//        expanded_about_center, translated, scaled, aabb_from_pair,
//        compute_overlap_with} (impact_geometry/src/axis_aligned_box.rs:253-364)
//   OrientedBox::{from_axis_aligned_box, rotated, compute_corners}
//        (impact_geometry/src/oriented_box.rs:62-68, 189-214)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct V3 {
    float x, y, z;
};

static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 v3s(float s) { return V3{s, s, s}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
static inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 mulc(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
static inline V3 vabs(V3 a) { return V3{std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
static inline V3 vmin(V3 a, V3 b) {
    return V3{std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)};
}
static inline V3 vmax(V3 a, V3 b) {
    return V3{std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)};
}
// glam dot3_in_x: (x*x + y*y) + z*z
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
// glam Vec3A::cross: (self.zxy()*rhs - self*rhs.zxy()).zxy()
static inline V3 cross(V3 l, V3 r) {
    return V3{l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x};
}
static inline float max_component(V3 a) { return std::fmax(std::fmax(a.x, a.z), a.y); }
static inline bool sign_neg(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return (u >> 31) != 0;
}
static inline uint32_t neg_mask(V3 a) {
    return (sign_neg(a.x) ? 1u : 0u) | (sign_neg(a.y) ? 2u : 0u) | (sign_neg(a.z) ? 4u : 0u);
}
static inline float comp(V3 a, int d) { return d == 0 ? a.x : (d == 1 ? a.y : a.z); }

struct Quat {
    float x, y, z, w;
};

// Column-major 4x4; c[col][row].
struct M4 {
    float c[4][4];
};

static inline M4 m4_identity() {
    M4 m{};
    for (int i = 0; i < 4; ++i) m.c[i][i] = 1.0f;
    return m;
}
static inline V3 col3(const M4& m, int j) { return V3{m.c[j][0], m.c[j][1], m.c[j][2]}; }

// glam Mat4::transform_point3a: res = X*p.x; res = Y*p.y + res; res = Z*p.z + res; res = W + res
static inline V3 transform_point(const M4& m, V3 p) {
    float r[3];
    for (int i = 0; i < 3; ++i) {
        float v = m.c[0][i] * p.x;
        v = m.c[1][i] * p.y + v;
        v = m.c[2][i] * p.z + v;
        v = m.c[3][i] + v;
        r[i] = v;
    }
    return V3{r[0], r[1], r[2]};
}

// glam Mat4::mul_mat4 → per column mul_vec4: ((X*v.x + Y*v.y) + Z*v.z) + W*v.w
static inline M4 m4_mul(const M4& a, const M4& b) {
    M4 r{};
    for (int j = 0; j < 4; ++j) {
        for (int i = 0; i < 4; ++i) {
            float v = a.c[0][i] * b.c[j][0];
            v = v + a.c[1][i] * b.c[j][1];
            v = v + a.c[2][i] * b.c[j][2];
            v = v + a.c[3][i] * b.c[j][3];
            r.c[j][i] = v;
        }
    }
    return r;
}

// glam Mat4::from_quat / Mat3A::from_quat (quat_to_axes)
static inline void quat_axes(Quat q, V3& ax, V3& ay, V3& az) {
    float x = q.x, y = q.y, z = q.z, w = q.w;
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx = x * x2, xy = x * y2, xz = x * z2;
    float yy = y * y2, yz = y * z2, zz = z * z2;
    float wx = w * x2, wy = w * y2, wz = w * z2;
    ax = V3{1.0f - (yy + zz), xy + wz, xz - wy};
    ay = V3{xy - wz, 1.0f - (xx + zz), yz + wx};
    az = V3{xz + wy, yz - wx, 1.0f - (xx + yy)};
}
static inline M4 m4_from_quat(Quat q) {
    V3 ax, ay, az;
    quat_axes(q, ax, ay, az);
    M4 m{};
    m.c[0][0] = ax.x; m.c[0][1] = ax.y; m.c[0][2] = ax.z;
    m.c[1][0] = ay.x; m.c[1][1] = ay.y; m.c[1][2] = ay.z;
    m.c[2][0] = az.x; m.c[2][1] = az.y; m.c[2][2] = az.z;
    m.c[3][3] = 1.0f;
    return m;
}
static inline Quat quat_conj(Quat q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
// glam Quat::mul_vec3a: v*(w*w - b.b) + b*(2*(v.b)) + (b x v)*(2*w)
static inline V3 quat_rotate(Quat q, V3 v) {
    V3 b{q.x, q.y, q.z};
    float b2 = dot(b, b);
    V3 t1 = v * (q.w * q.w - b2);
    V3 t2 = b * (dot(v, b) * 2.0f);
    V3 t3 = cross(b, v) * (q.w * 2.0f);
    return (t1 + t2) + t3;
}

struct Aabb {
    V3 lo, hi;
};

static inline V3 aabb_center(const Aabb& b) { return 0.5f * (b.lo + b.hi); }
static inline V3 aabb_extents(const Aabb& b) { return b.hi - b.lo; }
static inline V3 aabb_half_extents(const Aabb& b) { return 0.5f * aabb_extents(b); }
static inline Aabb aabb_expanded(const Aabb& b, float margin) {
    V3 m = v3s(margin);
    return Aabb{b.lo - m, b.hi + m};
}
static inline Aabb aabb_translated(const Aabb& b, V3 d) { return Aabb{b.lo + d, b.hi + d}; }
static inline Aabb aabb_scaled(const Aabb& b, float s) { return Aabb{s * b.lo, s * b.hi}; }
static inline Aabb aabb_from_pair(const Aabb& a, const Aabb& b) {
    return Aabb{vmin(a.lo, b.lo), vmax(a.hi, b.hi)};
}
static inline bool aabb_overlap(const Aabb& a, const Aabb& b, Aabb& out) {
    V3 lo = vmax(a.lo, b.lo);
    V3 hi = vmin(a.hi, b.hi);
    if (neg_mask(hi - lo) != 0) return false;
    out = Aabb{lo, hi};
    return true;
}
// self.box_lies_outside(other)
static inline bool aabb_box_lies_outside(const Aabb& self, const Aabb& other) {
    return (neg_mask(other.hi - self.lo) | neg_mask(self.hi - other.lo)) != 0;
}
// self.contains_box(other)
static inline bool aabb_contains_box(const Aabb& self, const Aabb& other) {
    return (neg_mask(other.lo - self.lo) | neg_mask(self.hi - other.hi)) == 0;
}
static inline Aabb aabb_of_transformed(const Aabb& b, const M4& m) {
    V3 c = transform_point(m, aabb_center(b));
    V3 h = aabb_half_extents(b);
    V3 ax = vabs(col3(m, 0)), ay = vabs(col3(m, 1)), az = vabs(col3(m, 2));
    // glam Mat3A::mul_vec3a: (X*h.x + Y*h.y) + Z*h.z
    V3 th = (ax * h.x + ay * h.y) + az * h.z;
    return Aabb{c - th, c + th};
}
// OrientedBox::from_axis_aligned_box(b).rotated(q).compute_corners() → aabb_for_point_array
static inline Aabb aabb_of_rotated_obb(const Aabb& b, Quat q) {
    V3 center = quat_rotate(q, aabb_center(b));
    V3 half = aabb_half_extents(b);
    V3 ax, ay, az;
    quat_axes(q, ax, ay, az);  // rotation * identity == rotation
    V3 hw = half.x * ax, hh = half.y * ay, hd = half.z * az;
    V3 pts[8] = {
        ((center - hw) - hh) - hd, ((center - hw) - hh) + hd, ((center - hw) + hh) - hd,
        ((center - hw) + hh) + hd, ((center + hw) - hh) - hd, ((center + hw) - hh) + hd,
        ((center + hw) + hh) - hd, ((center + hw) + hh) + hd,
    };
    V3 lo = pts[0], hi = pts[0];
    for (int i = 1; i < 8; ++i) {
        lo = vmin(lo, pts[i]);
        hi = vmax(hi, pts[i]);
    }
    return Aabb{lo, hi};
}

// compiler-rt __powisf2 (what Rust's f32::powi lowers to)
static inline float powi(float a, int b) {
    const bool recip = b < 0;
    float r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}

}  // namespace orc
