// Device-side f32 building blocks of the voxel hot path.
//
// Everything here reproduces the reference's arithmetic operation for
// operation (IEEE f32, round-to-nearest, no FMA contraction), because the
// i8 quantisation `(sd * 50.0) as i8` (lib.rs:195-201) turns a 1-ulp difference
// at any multiple of 0.02 into a different stored byte. This file must be
// compiled with -fmad=false (the Makefile passes -DIVX_FMAD_OFF to prove it).
#pragma once
#ifndef IVX_FMAD_OFF
#error "compile with -fmad=false -DIVX_FMAD_OFF: FMA contraction changes quantised voxels"
#endif
#include <cuda_runtime.h>
#include <stdint.h>

#include "ivx_internal.h"

namespace ivx {

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return f3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float norm3(f3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ bool sign_neg(float v) { return (__float_as_uint(v) >> 31) != 0u; }

// Matrix4::transform_point → glam Mat4::transform_point3a (matrix.rs:678-682)
__device__ __forceinline__ f3 transform_point(const float* __restrict__ M, f3 p) {
    f3 r;
    r.x = M[0] * p.x;  r.y = M[1] * p.x;  r.z = M[2] * p.x;
    r.x = M[4] * p.y + r.x;  r.y = M[5] * p.y + r.y;  r.z = M[6] * p.y + r.z;
    r.x = M[8] * p.z + r.x;  r.y = M[9] * p.z + r.y;  r.z = M[10] * p.z + r.z;
    r.x = M[12] + r.x;  r.y = M[13] + r.y;  r.z = M[14] + r.z;
    return r;
}

// AxisAlignedBox::aabb_of_transformed (axis_aligned_box.rs:350-364) of the
// root-space box [lo, lo+extent].
__device__ __forceinline__ void aabb_of_transformed(const float* __restrict__ M, f3 lo, f3 hi, f3& out_lo,
                                                    f3& out_hi) {
    f3 c = 0.5f * (lo + hi);
    f3 h = 0.5f * (hi - lo);
    f3 tc = transform_point(M, c);
    f3 th;
    th.x = (fabsf(M[0]) * h.x + fabsf(M[4]) * h.y) + fabsf(M[8]) * h.z;
    th.y = (fabsf(M[1]) * h.x + fabsf(M[5]) * h.y) + fabsf(M[9]) * h.z;
    th.z = (fabsf(M[2]) * h.x + fabsf(M[6]) * h.y) + fabsf(M[10]) * h.z;
    out_lo = tc - th;
    out_hi = tc + th;
}

// domain.box_lies_outside(block) (axis_aligned_box.rs:262-267)
__device__ __forceinline__ bool box_lies_outside(const float* dlo, const float* dhi, f3 blo, f3 bhi) {
    return sign_neg(bhi.x - dlo[0]) | sign_neg(bhi.y - dlo[1]) | sign_neg(bhi.z - dlo[2]) |
           sign_neg(dhi[0] - blo.x) | sign_neg(dhi[1] - blo.y) | sign_neg(dhi[2] - blo.z);
}
// [-h, h].contains_box(block) (axis_aligned_box.rs:253-258)
__device__ __forceinline__ bool sym_box_contains(f3 h, f3 blo, f3 bhi) {
    return !(sign_neg(blo.x - (-h.x)) | sign_neg(blo.y - (-h.y)) | sign_neg(blo.z - (-h.z)) |
             sign_neg(h.x - bhi.x) | sign_neg(h.y - bhi.y) | sign_neg(h.z - bhi.z));
}

// expanded_interior_domain_bounds(-margin) / expanded_domain_bounds(-margin)
// (atomic.rs:1171-1180, 1224-1234, 1279-1285): half extents of the box inside
// which a leaf is assumed <= -margin.
__device__ __forceinline__ f3 leaf_interior_half_extents(const ivx_node& n) {
    const float FRAC_1_SQRT_3 = 0.57735026f;
    float m = -n.margin;
    if (n.kind == IVX_SPHERE) {
        float e = n.p[0] * FRAC_1_SQRT_3 + m;
        return mk3(e, e, e);
    } else if (n.kind == IVX_CAPSULE) {
        float e = n.p[1] * FRAC_1_SQRT_3 + m;
        return mk3(e, e + n.p[0], e);
    }
    return mk3(n.p[0] + m, n.p[1] + m, n.p[2] + m);
}

// Primitive signed distances (atomic.rs:1183-1190, 1237-1252, 1288-1291)
__device__ __forceinline__ float sd_leaf(uint32_t kind, const float* __restrict__ p, f3 q) {
    if (kind == IVX_SPHERE) {
        return norm3(q) - p[0];
    } else if (kind == IVX_CAPSULE) {
        float h = p[0];
        float c = q.y;
        if (c < -h) c = -h;
        if (c > h) c = h;
        q.y -= c;
        return norm3(q) - p[1];
    }
    f3 d = mk3(fabsf(q.x) - p[0], fabsf(q.y) - p[1], fabsf(q.z) - p[2]);
    f3 m = mk3(fmaxf(d.x, 0.0f), fmaxf(d.y, 0.0f), fmaxf(d.z, 0.0f));
    float mc = fmaxf(fmaxf(d.x, d.z), d.y);
    return norm3(m) + fminf(mc, 0.0f);
}

// CSG operators (generation/sdf.rs:46-102)
__device__ __forceinline__ float smooth_union(float d1, float d2, float k, float qik) {
    float h = fmaxf(k - fabsf(d1 - d2), 0.0f);
    return fminf(d1, d2) - (h * h) * qik;
}
__device__ __forceinline__ float op_combine(uint32_t kind, float d1, float d2, float k, float qik) {
    if (kind == IVX_UNION) return k == 0.0f ? fminf(d1, d2) : smooth_union(d1, d2, k, qik);
    if (kind == IVX_SUBTRACTION) return k == 0.0f ? fmaxf(d1, -d2) : -smooth_union(-d1, d2, k, qik);
    return k == 0.0f ? fmaxf(d1, d2) : -smooth_union(-d1, -d2, k, qik);
}

// VoxelSignedDistance::from_f32 (lib.rs:195-201): saturating truncating cast, NaN → 0
__device__ __forceinline__ int sd_encode(float v) {
    float s = v * 50.0f;
    if (s != s) return 0;
    if (s >= 127.0f) return 127;
    if (s <= -128.0f) return -128;
    return (int)s;  // cvt.rzi
}
__device__ __forceinline__ float sd_decode(int e) { return (float)e * 0.02f; }

// ---- simplex noise ---------------------------------------------------------
// Restates simdnoise 3.1.x (third-party, not under /root/reference; see
// DESIGN.md "noise parity"): 3-D hashed-gradient simplex + fBm, 4-D
// permutation-table simplex. Bit-identical to oracle/noise.cpp.
__device__ __forceinline__ float xor_sign(float v, uint32_t bits) {
    return __uint_as_float(__float_as_uint(v) ^ bits);
}
__device__ __forceinline__ float grad3d_dot(int32_t seed, uint32_t i, uint32_t j, uint32_t k, float x, float y,
                                            float z) {
    uint32_t hash = i ^ (uint32_t)seed;
    hash = j ^ hash;
    hash = k ^ hash;
    hash = ((hash * hash) * 60493u) * hash;
    hash = (uint32_t)((int32_t)hash >> 13) ^ hash;
    uint32_t h13 = hash & 13u;
    float u = (h13 < 8u) ? x : y;
    float v = (h13 < 2u) ? y : ((h13 == 12u) ? x : z);
    return xor_sign(u, hash << 31) + xor_sign(v, (hash & 2u) << 30);
}

__device__ __forceinline__ float simplex3(float x, float y, float z, int32_t seed) {
    const float F3 = 1.0f / 3.0f, G3 = 1.0f / 6.0f, G33 = 3.0f / 6.0f - 1.0f;
    float f = F3 * ((x + y) + z);
    float x0 = floorf(x + f), y0 = floorf(y + f), z0 = floorf(z + f);
    uint32_t i = (uint32_t)(int32_t)x0 * 1619u;
    uint32_t j = (uint32_t)(int32_t)y0 * 31337u;
    uint32_t k = (uint32_t)(int32_t)z0 * 6791u;
    float g = G3 * ((x0 + y0) + z0);
    x0 = x - (x0 - g);
    y0 = y - (y0 - g);
    z0 = z - (z0 - g);
    bool x_ge_y = x0 >= y0, y_ge_z = y0 >= z0, x_ge_z = x0 >= z0;
    bool i1 = x_ge_y && x_ge_z, j1 = !x_ge_y && y_ge_z, k1 = !x_ge_z && !y_ge_z;
    bool i2 = x_ge_y || x_ge_z, j2 = !x_ge_y || y_ge_z, k2 = !(x_ge_z && y_ge_z);
    float x1 = (x0 - (i1 ? 1.0f : 0.0f)) + G3, y1 = (y0 - (j1 ? 1.0f : 0.0f)) + G3,
          z1 = (z0 - (k1 ? 1.0f : 0.0f)) + G3;
    float x2 = (x0 - (i2 ? 1.0f : 0.0f)) + F3, y2 = (y0 - (j2 ? 1.0f : 0.0f)) + F3,
          z2 = (z0 - (k2 ? 1.0f : 0.0f)) + F3;
    float x3 = x0 + G33, y3 = y0 + G33, z3 = z0 + G33;
    float t0 = ((0.6f - x0 * x0) - y0 * y0) - z0 * z0;
    float t1 = ((0.6f - x1 * x1) - y1 * y1) - z1 * z1;
    float t2 = ((0.6f - x2 * x2) - y2 * y2) - z2 * z2;
    float t3 = ((0.6f - x3 * x3) - y3 * y3) - z3 * z3;
    t0 = (t0 >= 0.0f) ? t0 : 0.0f;
    t1 = (t1 >= 0.0f) ? t1 : 0.0f;
    t2 = (t2 >= 0.0f) ? t2 : 0.0f;
    t3 = (t3 >= 0.0f) ? t3 : 0.0f;
    t0 = t0 * t0; t1 = t1 * t1; t2 = t2 * t2; t3 = t3 * t3;
    t0 = t0 * t0; t1 = t1 * t1; t2 = t2 * t2; t3 = t3 * t3;
    float v0 = t0 * grad3d_dot(seed, i, j, k, x0, y0, z0);
    float v1 = t1 * grad3d_dot(seed, i + (i1 ? 1619u : 0u), j + (j1 ? 31337u : 0u), k + (k1 ? 6791u : 0u), x1,
                               y1, z1);
    float v2 = t2 * grad3d_dot(seed, i + (i2 ? 1619u : 0u), j + (j2 ? 31337u : 0u), k + (k2 ? 6791u : 0u), x2,
                               y2, z2);
    float v3 = t3 * grad3d_dot(seed, i + 1619u, j + 31337u, k + 6791u, x3, y3, z3);
    float p1 = v3 + v2;
    float p2 = p1 + v1;
    return (p2 + v0) * 32.69428253173828125f;
}

__device__ __forceinline__ float fbm3(float x, float y, float z, float lacunarity, float gain, uint32_t octaves,
                                      int32_t seed) {
    uint32_t oct = octaves & 0xFFu;
    float amp = 1.0f;
    float result = simplex3(x, y, z, seed);
    for (uint32_t o = 1; o < oct; ++o) {
        x = x * lacunarity;
        y = y * lacunarity;
        z = z * lacunarity;
        amp = amp * gain;
        result = simplex3(x, y, z, seed) * amp + result;
    }
    return result;
}

__device__ __forceinline__ float grad4(int32_t seed, int32_t hash, float x, float y, float z, float t) {
    int32_t h = (seed ^ hash) & 31;
    float u = (24 > h) ? x : y;
    float v = (16 > h) ? y : z;
    float w = (8 > h) ? z : t;
    float a = ((h & 1) == 0) ? u : (0.0f - u);
    float b = ((h & 2) == 0) ? v : (0.0f - v);
    float c = ((h & 4) == 0) ? w : (0.0f - w);
    return a + (b + c);
}

template <typename PermT>
__device__ __forceinline__ float simplex4_t(float x, float y, float z, float w, int32_t seed,
                                            const PermT* __restrict__ perm) {
    const float F4 = 0.309016994f, G4 = 0.138196601f;
    const float G24 = 2.0f * 0.138196601f, G34 = 3.0f * 0.138196601f, G44 = 4.0f * 0.138196601f;
    float s = F4 * (x + (y + (z + w)));
    float ips = floorf(x + s), jps = floorf(y + s), kps = floorf(z + s), lps = floorf(w + s);
    int32_t i = (int32_t)ips, j = (int32_t)jps, k = (int32_t)kps, l = (int32_t)lps;
    float t = (float)(i + (j + (k + l))) * G4;
    float x0 = x - (ips - t), y0 = y - (jps - t), z0 = z - (kps - t), w0 = w - (lps - t);
    int rx = 0, ry = 0, rz = 0, rw = 0;
    if (x0 > y0) rx++; else ry++;
    if (x0 > z0) rx++; else rz++;
    if (x0 > w0) rx++; else rw++;
    if (y0 > z0) ry++; else rz++;
    if (y0 > w0) ry++; else rw++;
    if (z0 > w0) rz++; else rw++;
    int32_t i1 = rx > 2, j1 = ry > 2, k1 = rz > 2, l1 = rw > 2;
    int32_t i2 = rx > 1, j2 = ry > 1, k2 = rz > 1, l2 = rw > 1;
    int32_t i3 = rx > 0, j3 = ry > 0, k3 = rz > 0, l3 = rw > 0;
    float x1 = (x0 - (float)i1) + G4, y1 = (y0 - (float)j1) + G4, z1 = (z0 - (float)k1) + G4,
          w1 = (w0 - (float)l1) + G4;
    float x2 = (x0 - (float)i2) + G24, y2 = (y0 - (float)j2) + G24, z2 = (z0 - (float)k2) + G24,
          w2 = (w0 - (float)l2) + G24;
    float x3 = (x0 - (float)i3) + G34, y3 = (y0 - (float)j3) + G34, z3 = (z0 - (float)k3) + G34,
          w3 = (w0 - (float)l3) + G34;
    float x4 = (x0 - 1.0f) + G44, y4 = (y0 - 1.0f) + G44, z4 = (z0 - 1.0f) + G44, w4 = (w0 - 1.0f) + G44;
    int32_t ii = i & 0xff, jj = j & 0xff, kk = k & 0xff, ll = l & 0xff;
#define IVX_GI(di, dj, dk, dl) \
    ((int32_t)perm[(ii + (di) + (int32_t)perm[(jj + (dj) + (int32_t)perm[(kk + (dk) + (int32_t)perm[(ll + (dl)) & 255]) & 255]) & 255]) & 255])
    int32_t gi0 = IVX_GI(0, 0, 0, 0), gi1 = IVX_GI(i1, j1, k1, l1), gi2 = IVX_GI(i2, j2, k2, l2),
            gi3 = IVX_GI(i3, j3, k3, l3), gi4 = IVX_GI(1, 1, 1, 1);
#undef IVX_GI
    float t0 = (((0.5f - x0 * x0) - y0 * y0) - z0 * z0) - w0 * w0;
    float t1 = (((0.5f - x1 * x1) - y1 * y1) - z1 * z1) - w1 * w1;
    float t2 = (((0.5f - x2 * x2) - y2 * y2) - z2 * z2) - w2 * w2;
    float t3 = (((0.5f - x3 * x3) - y3 * y3) - z3 * z3) - w3 * w3;
    float t4 = (((0.5f - x4 * x4) - y4 * y4) - z4 * z4) - w4 * w4;
    float q0 = t0 * t0, q1 = t1 * t1, q2 = t2 * t2, q3 = t3 * t3, q4 = t4 * t4;
    q0 = q0 * q0; q1 = q1 * q1; q2 = q2 * q2; q3 = q3 * q3; q4 = q4 * q4;
    float n0 = q0 * grad4(seed, gi0, x0, y0, z0, w0);
    float n1 = q1 * grad4(seed, gi1, x1, y1, z1, w1);
    float n2 = q2 * grad4(seed, gi2, x2, y2, z2, w2);
    float n3 = q3 * grad4(seed, gi3, x3, y3, z3, w3);
    float n4 = q4 * grad4(seed, gi4, x4, y4, z4, w4);
    if (t0 < 0.0f) n0 = 0.0f;
    if (t1 < 0.0f) n1 = 0.0f;
    if (t2 < 0.0f) n2 = 0.0f;
    if (t3 < 0.0f) n3 = 0.0f;
    if (t4 < 0.0f) n4 = 0.0f;
    return (n0 + (n1 + (n2 + (n3 + n4)))) * 62.77772078955791f;
}

// simdnoise's permutation table of the 4-D simplex noise (one copy per translation unit)
static __constant__ uint8_t c_perm[256] = {
    151, 160, 137, 91,  90,  15,  131, 13,  201, 95,  96,  53,  194, 233, 7,   225, 140, 36,  103,
    30,  69,  142, 8,   99,  37,  240, 21,  10,  23,  190, 6,   148, 247, 120, 234, 75,  0,   26,
    197, 62,  94,  252, 219, 203, 117, 35,  11,  32,  57,  177, 33,  88,  237, 149, 56,  87,  174,
    20,  125, 136, 171, 168, 68,  175, 74,  165, 71,  134, 139, 48,  27,  166, 77,  146, 158, 231,
    83,  111, 229, 122, 60,  211, 133, 230, 220, 105, 92,  41,  55,  46,  245, 40,  244, 102, 143,
    54,  65,  25,  63,  161, 1,   216, 80,  73,  209, 76,  132, 187, 208, 89,  18,  169, 200, 196,
    135, 130, 116, 188, 159, 86,  164, 100, 109, 198, 173, 186, 3,   64,  52,  217, 226, 250, 124,
    123, 5,   202, 38,  147, 118, 126, 255, 82,  85,  212, 207, 206, 59,  227, 47,  16,  58,  17,
    182, 189, 28,  42,  223, 183, 170, 213, 119, 248, 152, 2,   44,  154, 163, 70,  221, 153, 101,
    155, 167, 43,  172, 9,   129, 22,  39,  253, 19,  98,  108, 110, 79,  113, 224, 232, 178, 185,
    112, 104, 218, 246, 97,  228, 251, 34,  242, 193, 238, 210, 144, 12,  191, 179, 162, 241, 81,
    51,  145, 235, 249, 14,  239, 107, 49,  192, 214, 31,  181, 199, 106, 157, 184, 84,  204, 176,
    115, 121, 50,  45,  127, 4,   150, 254, 138, 236, 205, 93,  222, 114, 67,  29,  24,  72,  243,
    141, 128, 195, 78,  66,  215, 61,  156, 180};

// simdnoise walks the slow axes of a block call by repeated += 1.0 (not start + n)
__device__ __forceinline__ float accumulate_ones(float start, int n) {
    for (int t = 0; t < n; ++t) start = start + 1.0f;
    return start;
}

// Linear voxel index in a chunk / cell index in an 18³ brick
__device__ __forceinline__ int vidx(int i, int j, int k) { return (i << 8) + (j << 4) + k; }
__device__ __forceinline__ int bidx(int i, int j, int k) { return i * 324 + j * 18 + k; }

}  // namespace ivx
