// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
//
// PARITY UNPINNED. Scalar restatement of the simplex noise used by the
// reference through the third-party crate `simdnoise` 3.1.7
// (git+https://github.com/lars-frogner/rust-simd-noise#e59c958e, see
// engine/Cargo.lock:3167-3173), which is NOT present under /root/reference.
// Call sites: V/generation/sdf/atomic.rs:1460-1475, 1487-1502, 1547-1562
// (`fbm_3d_offset`) and V/generation/voxel_type.rs:133-152
// (`gradient_4d_offset`). No test, fixture or snapshot in the reference pins
// noise output, so this file follows the published simdnoise 3.1.x algorithm
// (FastNoiseSIMD-style hashed-gradient 3-D simplex; permutation-table 4-D
// simplex) as recalled; constants cannot be verified in this environment.
// What IS pinned: GPU kernel == this restatement, bit for bit.
#include <cmath>

#include "oracle.hpp"

namespace orc {

static const float F3 = 1.0f / 3.0f;
static const float G3 = 1.0f / 6.0f;
static const float G33 = 3.0f / 6.0f - 1.0f;
static const float F4 = 0.309016994f;
static const float G4 = 0.138196601f;
static const float G24 = 2.0f * 0.138196601f;
static const float G34 = 3.0f * 0.138196601f;
static const float G44 = 4.0f * 0.138196601f;
static const int32_t X_PRIME = 1619;
static const int32_t Y_PRIME = 31337;
static const int32_t Z_PRIME = 6791;
static const float SIMPLEX3_SCALE = 32.69428253173828125f;
static const float SIMPLEX4_SCALE = 62.77772078955791f;

static const uint8_t PERM[256] = {
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
static inline int32_t perm(int32_t i) { return (int32_t)PERM[i & 255]; }

static inline float xor_sign(float v, uint32_t signbits) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    u ^= signbits;
    std::memcpy(&v, &u, 4);
    return v;
}

// hash3d + grad3d_dot: gradient towards an edge midpoint of the double-unit cube.
static inline float grad3d_dot(int32_t seed, int32_t i, int32_t j, int32_t k, float x, float y,
                               float z) {
    uint32_t hash = (uint32_t)i ^ (uint32_t)seed;
    hash = (uint32_t)j ^ hash;
    hash = (uint32_t)k ^ hash;
    hash = ((hash * hash) * 60493u) * hash;
    hash = (uint32_t)((int32_t)hash >> 13) ^ hash;
    uint32_t h13 = hash & 13u;
    bool l8 = h13 < 8u;
    bool l4 = h13 < 2u;
    bool h12 = h13 == 12u;
    float u = l8 ? x : y;
    float v = l4 ? y : (h12 ? x : z);
    uint32_t h1 = hash << 31;
    uint32_t h2 = (hash & 2u) << 30;
    return xor_sign(u, h1) + xor_sign(v, h2);
}

float simplex3(float x, float y, float z, int32_t seed) {
    float f = F3 * ((x + y) + z);
    float x0 = std::floor(x + f);
    float y0 = std::floor(y + f);
    float z0 = std::floor(z + f);
    int32_t i = (int32_t)((uint32_t)(int32_t)x0 * (uint32_t)X_PRIME);
    int32_t j = (int32_t)((uint32_t)(int32_t)y0 * (uint32_t)Y_PRIME);
    int32_t k = (int32_t)((uint32_t)(int32_t)z0 * (uint32_t)Z_PRIME);
    float g = G3 * ((x0 + y0) + z0);
    x0 = x - (x0 - g);
    y0 = y - (y0 - g);
    z0 = z - (z0 - g);
    bool x_ge_y = x0 >= y0, y_ge_z = y0 >= z0, x_ge_z = x0 >= z0;
    bool i1 = x_ge_y && x_ge_z;
    bool j1 = !x_ge_y && y_ge_z;
    bool k1 = !x_ge_z && !y_ge_z;
    bool i2 = x_ge_y || x_ge_z;
    bool j2 = !x_ge_y || y_ge_z;
    bool k2 = !(x_ge_z && y_ge_z);
    float x1 = (x0 - (i1 ? 1.0f : 0.0f)) + G3;
    float y1 = (y0 - (j1 ? 1.0f : 0.0f)) + G3;
    float z1 = (z0 - (k1 ? 1.0f : 0.0f)) + G3;
    float x2 = (x0 - (i2 ? 1.0f : 0.0f)) + F3;
    float y2 = (y0 - (j2 ? 1.0f : 0.0f)) + F3;
    float z2 = (z0 - (k2 ? 1.0f : 0.0f)) + F3;
    float x3 = x0 + G33;
    float y3 = y0 + G33;
    float z3 = z0 + G33;
    float t0 = ((0.6f - x0 * x0) - y0 * y0) - z0 * z0;
    float t1 = ((0.6f - x1 * x1) - y1 * y1) - z1 * z1;
    float t2 = ((0.6f - x2 * x2) - y2 * y2) - z2 * z2;
    float t3 = ((0.6f - x3 * x3) - y3 * y3) - z3 * z3;
    if (!(t0 >= 0.0f)) t0 = 0.0f;
    if (!(t1 >= 0.0f)) t1 = 0.0f;
    if (!(t2 >= 0.0f)) t2 = 0.0f;
    if (!(t3 >= 0.0f)) t3 = 0.0f;
    float t20 = t0 * t0, t21 = t1 * t1, t22 = t2 * t2, t23 = t3 * t3;
    float t40 = t20 * t20, t41 = t21 * t21, t42 = t22 * t22, t43 = t23 * t23;
    auto wadd = [](int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); };
    float v0 = t40 * grad3d_dot(seed, i, j, k, x0, y0, z0);
    float v1 = t41 * grad3d_dot(seed, wadd(i, i1 ? X_PRIME : 0), wadd(j, j1 ? Y_PRIME : 0),
                                wadd(k, k1 ? Z_PRIME : 0), x1, y1, z1);
    float v2 = t42 * grad3d_dot(seed, wadd(i, i2 ? X_PRIME : 0), wadd(j, j2 ? Y_PRIME : 0),
                                wadd(k, k2 ? Z_PRIME : 0), x2, y2, z2);
    float v3 = t43 * grad3d_dot(seed, wadd(i, X_PRIME), wadd(j, Y_PRIME), wadd(k, Z_PRIME), x3, y3, z3);
    float p1 = v3 + v2;
    float p2 = p1 + v1;
    return (p2 + v0) * SIMPLEX3_SCALE;
}

float fbm3(float x, float y, float z, float lacunarity, float gain, uint32_t octaves,
           int32_t seed) {
    // `octaves as u8` at the call site (atomic.rs:1470)
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

static inline float grad4(int32_t seed, int32_t hash, float x, float y, float z, float t) {
    int32_t h = (seed ^ hash) & 31;
    float u = (24 > h) ? x : y;
    float v = (16 > h) ? y : z;
    float w = (8 > h) ? z : t;
    float a = ((h & 1) == 0) ? u : (0.0f - u);
    float b = ((h & 2) == 0) ? v : (0.0f - v);
    float c = ((h & 4) == 0) ? w : (0.0f - w);
    return a + (b + c);
}

float simplex4(float x, float y, float z, float w, int32_t seed) {
    float s = F4 * (x + (y + (z + w)));
    float ips = std::floor(x + s), jps = std::floor(y + s), kps = std::floor(z + s),
          lps = std::floor(w + s);
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
    float x4 = (x0 - 1.0f) + G44, y4 = (y0 - 1.0f) + G44, z4 = (z0 - 1.0f) + G44,
          w4 = (w0 - 1.0f) + G44;
    int32_t ii = i & 0xff, jj = j & 0xff, kk = k & 0xff, ll = l & 0xff;
    auto gi = [&](int32_t di, int32_t dj, int32_t dk, int32_t dl) {
        int32_t lp = perm(ll + dl);
        int32_t kp = perm(kk + dk + lp);
        int32_t jp = perm(jj + dj + kp);
        return perm(ii + di + jp);
    };
    int32_t gi0 = gi(0, 0, 0, 0), gi1 = gi(i1, j1, k1, l1), gi2 = gi(i2, j2, k2, l2),
            gi3 = gi(i3, j3, k3, l3), gi4 = gi(1, 1, 1, 1);
    auto tw = [](float a, float b, float c, float d) {
        return (((0.5f - a * a) - b * b) - c * c) - d * d;
    };
    float t0 = tw(x0, y0, z0, w0), t1 = tw(x1, y1, z1, w1), t2 = tw(x2, y2, z2, w2),
          t3 = tw(x3, y3, z3, w3), t4 = tw(x4, y4, z4, w4);
    auto q = [](float v) {
        float v2 = v * v;
        return v2 * v2;
    };
    float n0 = q(t0) * grad4(seed, gi0, x0, y0, z0, w0);
    float n1 = q(t1) * grad4(seed, gi1, x1, y1, z1, w1);
    float n2 = q(t2) * grad4(seed, gi2, x2, y2, z2, w2);
    float n3 = q(t3) * grad4(seed, gi3, x3, y3, z3, w3);
    float n4 = q(t4) * grad4(seed, gi4, x4, y4, z4, w4);
    if (t0 < 0.0f) n0 = 0.0f;
    if (t1 < 0.0f) n1 = 0.0f;
    if (t2 < 0.0f) n2 = 0.0f;
    if (t3 < 0.0f) n3 = 0.0f;
    if (t4 < 0.0f) n4 = 0.0f;
    return (n0 + (n1 + (n2 + (n3 + n4)))) * SIMPLEX4_SCALE;
}

}  // namespace orc
