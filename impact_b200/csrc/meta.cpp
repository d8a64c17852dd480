// Meta-graph compiler: `MetaSDFGraph::build_in` (engine/crates/impact_voxel/src/generation/sdf/meta.rs:741-896) with the
// `resolve` of all 21 node kinds (meta.rs:1194-2260), parameter sampling (meta/params.rs:85-264), the seeded RNG
// (impact_math/src/random.rs:11-100 over fastrand 2.3.0's wyrand, random/splitmix.rs) and the surface probes of the
// three nodes that sample an SDF (meta.rs:2411-2769). Host C++; the probes run on the device through
// ivx_program_build / ivx_program_eval_blocks, all instances of a node in one batch.
//
// Input: the meta nodes as PODs (include/impact_voxel_cuda.h `ivx_meta_node`: the RON formats are read by the host —
// impact_b200/meta.py here, serde in the engine). Output: the atomic `SDFNode` list + root, ready for
// ivx_program_build.
//
// Third-party arithmetic restated from the published algorithms (not under /root/reference; unpinned like the noise,
// DESIGN.md §2): fastrand's wyrand generator and its Lemire range reduction; glam's Quat::from_rotation_arc /
// mul_vec3a / any_orthonormal_vector operation order; libm sin / cos / acos / pow are evaluated in double and rounded to
// f32 (the reference calls the f32 versions).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../include/impact_voxel_cuda.h"

namespace {

typedef float f32;
const f32 F32_EPS = 1.1920929e-07f;
const f32 PI_F = 3.14159274f;  // std::f32::consts::PI

struct V3 {
    f32 x, y, z;
};
V3 v3(f32 x, f32 y, f32 z) { return V3{x, y, z}; }
V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
V3 operator*(f32 s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
V3 operator*(V3 a, f32 s) { return v3(a.x * s, a.y * s, a.z * s); }
V3 operator/(V3 a, f32 s) { return v3(a.x / s, a.y / s, a.z / s); }
f32 dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
f32 norm(V3 a) { return std::sqrt(dot(a, a)); }
V3 cross(V3 l, V3 r) { return v3(l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x); }
f32 comp(V3 a, int d) { return d == 0 ? a.x : (d == 1 ? a.y : a.z); }

struct Quat {
    f32 x, y, z, w;
};
const Quat QID = {0.0f, 0.0f, 0.0f, 1.0f};
Quat quat_mul(Quat a, Quat b) {
    return Quat{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
                a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
// glam Quat::mul_vec3a: v (w^2 - b.b) + b (2 v.b) + (b x v)(2 w)
V3 quat_rotate(Quat q, V3 v) {
    const V3 b = v3(q.x, q.y, q.z);
    const f32 b2 = dot(b, b);
    const V3 t1 = v * (q.w * q.w - b2);
    const V3 t2 = b * (dot(v, b) * 2.0f);
    const V3 t3 = cross(b, v) * (q.w * 2.0f);
    return (t1 + t2) + t3;
}
Quat quat_conj(Quat q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
f32 sin_f(f32 a) { return (f32)std::sin((double)a); }
f32 cos_f(f32 a) { return (f32)std::cos((double)a); }
f32 acos_f(f32 a) { return (f32)std::acos((double)a); }
f32 pow_f(f32 a, f32 b) { return (f32)std::pow((double)a, (double)b); }
f32 clampf(f32 v, f32 lo, f32 hi) { return std::fmin(std::fmax(v, lo), hi); }
Quat quat_from_axis_angle(V3 axis, f32 angle) {
    const f32 half = angle * 0.5f;
    const f32 s = sin_f(half), c = cos_f(half);
    return Quat{axis.x * s, axis.y * s, axis.z * s, c};
}
V3 any_orthonormal_vector(V3 v) {  // glam Vec3::any_orthonormal_vector
    const f32 sign = std::copysign(1.0f, v.z);
    const f32 a = -1.0f / (sign + v.z);
    const f32 b = v.x * v.y * a;
    return v3(b, sign + v.y * v.y * a, -v.y);
}
Quat quat_from_rotation_arc(V3 from, V3 to) {  // glam Quat::from_rotation_arc
    const f32 one_minus_eps = 1.0f - 2.0f * F32_EPS;
    const f32 d = dot(from, to);
    if (d > one_minus_eps) return QID;
    if (d < -one_minus_eps) return quat_from_axis_angle(any_orthonormal_vector(from), PI_F);
    const V3 c = cross(from, to);
    const Quat q = {c.x, c.y, c.z, 1.0f + d};
    const f32 n = std::sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    return Quat{q.x / n, q.y / n, q.z / n, q.w / n};
}

// `Similarity3` (impact_math/src/transform/similarity.rs:22-26): scaling, then rotation, then translation
struct Sim {
    V3 t = {0.0f, 0.0f, 0.0f};
    Quat r = QID;
    f32 s = 1.0f;
    Sim translated(V3 d) const { return Sim{t + d, r, s}; }
    Sim rotated(Quat q) const { return Sim{quat_rotate(q, t), quat_mul(q, r), s}; }
    Sim scaled(f32 k) const { return Sim{k * t, r, k * s}; }
    Sim applied_to_translation(V3 d) const { return Sim{quat_rotate(r, s * d) + t, r, s}; }
    Sim applied_to_rotation(Quat q) const { return Sim{t, quat_mul(r, q), s}; }
    Sim applied_to_scaling(f32 k) const { return Sim{t, r, s * k}; }
    Sim mul(const Sim& b) const { return Sim{quat_rotate(r, s * b.t) + t, quat_mul(r, b.r), s * b.s}; }
    V3 transform_point(V3 p) const { return quat_rotate(r, s * p) + t; }
    V3 transform_vector(V3 v) const { return quat_rotate(r, s * v); }
    V3 inverse_transform_point(V3 p) const { return quat_rotate(quat_conj(r), p - t) / s; }
    V3 inverse_transform_vector(V3 v) const { return quat_rotate(quat_conj(r), v) / s; }
};

// ---- randomness ---------------------------------------------------------------------------------------------------
uint64_t splitmix(uint64_t state) {
    state += 0x9E3779B97F4A7C15ull;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t splitmix2(uint64_t a, uint64_t b) { return splitmix(a ^ splitmix(b)); }
uint64_t splitmix3(uint64_t a, uint64_t b, uint64_t c) { return splitmix2(splitmix2(a, b), c); }

struct Rng {  // fastrand::Rng::with_seed
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t gen_u64() {
        s += 0x2D358DCCAA6C78A5ull;
        const unsigned __int128 t = (unsigned __int128)s * (unsigned __int128)(s ^ 0x8BB84B93962EACC9ull);
        return (uint64_t)t ^ (uint64_t)(t >> 64);
    }
    uint32_t gen_u32() { return (uint32_t)gen_u64(); }
    uint32_t mod_u32(uint32_t n) {  // Lemire's nearly divisionless reduction
        uint32_t r = gen_u32();
        uint64_t m = (uint64_t)r * n;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo < n) {
            const uint32_t t = (0u - n) % n;
            while (lo < t) {
                r = gen_u32();
                m = (uint64_t)r * n;
                hi = (uint32_t)(m >> 32);
                lo = (uint32_t)m;
            }
        }
        return hi;
    }
    uint64_t mod_u64(uint64_t n) {
        uint64_t r = gen_u64();
        unsigned __int128 m = (unsigned __int128)r * n;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo < n) {
            const uint64_t t = (0ull - n) % n;
            while (lo < t) {
                r = gen_u64();
                m = (unsigned __int128)r * n;
                hi = (uint64_t)(m >> 64);
                lo = (uint64_t)m;
            }
        }
        return hi;
    }
    uint32_t u32_inclusive(uint32_t lo, uint32_t hi) {
        if (lo == 0u && hi == 0xFFFFFFFFu) return gen_u32();
        return lo + mod_u32(hi - lo + 1u);
    }
    uint64_t usize_inclusive(uint64_t lo, uint64_t hi) { return lo + mod_u64(hi - lo + 1ull); }
    f32 f32_fraction() {
        const uint32_t bits = 0x3F800000u + (gen_u32() >> 9);
        f32 v;
        std::memcpy(&v, &bits, 4);
        return v - 1.0f;
    }
    f32 f32_in_range(f32 start, f32 end) {
        const f32 t = f32_fraction();
        return start + t * (end - start);
    }
};

// ---- parameters (meta/params.rs) ------------------------------------------------------------------------------------
f32 src_eval(const ivx_meta_source& s, const f32* values) {
    if (s.kind == 0) return s.value;
    return s.value + s.scale * values[s.idx];  // offset + scale * parameter
}
uint32_t src_eval_discrete(const ivx_meta_source& s, const f32* values) {
    if (s.kind == 0) return (uint32_t)(int64_t)s.value;
    const f32 r = std::fmax(std::nearbyint(src_eval(s, values)), 0.0f);
    return (uint32_t)r;
}
int spec_sources(const ivx_meta_param& p) { return p.dist == 0 ? 1 : (p.dist == 3 ? 3 : 2); }

f32 sample_spec(const ivx_meta_param& p, const f32* values, Rng& rng, bool discrete) {
    if (discrete) {
        if (p.dist == 0) return (f32)src_eval_discrete(p.src[0], values);
        const uint32_t lo = src_eval_discrete(p.src[0], values);
        const uint32_t hi = std::max(src_eval_discrete(p.src[1], values), lo);
        return (f32)rng.u32_inclusive(lo, hi);
    }
    switch (p.dist) {
        case 0: return src_eval(p.src[0], values);
        case 1: {
            const f32 lo = src_eval(p.src[0], values);
            const f32 hi = std::fmax(src_eval(p.src[1], values), lo);
            return rng.f32_in_range(lo, hi);
        }
        case 2: {
            const f32 d2r = PI_F / 180.0f;
            f32 lo = src_eval(p.src[0], values) * d2r, hi = src_eval(p.src[1], values) * d2r;
            lo = clampf(lo, 0.0f, PI_F);
            hi = std::fmin(std::fmax(hi, lo), PI_F);
            const f32 min_cos = cos_f(hi), max_cos = cos_f(lo);
            const f32 c = rng.f32_in_range(min_cos, max_cos);
            return acos_f(c) * (180.0f / PI_F);
        }
        default: {
            const f32 lo = src_eval(p.src[0], values);
            const f32 hi = std::fmax(src_eval(p.src[1], values), lo);
            const f32 ex = src_eval(p.src[2], values);
            const f32 frac = rng.f32_fraction();
            const f32 a = 1.0f - ex;
            if (std::fabs(a) <= F32_EPS) return lo * pow_f(hi / lo, frac);
            const f32 lp = pow_f(lo, a), hp = pow_f(hi, a);
            return pow_f(lp + frac * (hp - lp), 1.0f / a);
        }
    }
}

// evaluate_params_for_node (params.rs:246-264): topological order, FIFO among ready parameters
bool sample_params(const ivx_meta_node& node, int n, uint32_t discrete_mask, Rng& rng, f32* values, std::string& err) {
    int dep_counts[IVX_META_MAX_PARAMS] = {0};
    std::vector<int> rev[IVX_META_MAX_PARAMS];
    for (int i = 0; i < n; ++i) {
        const ivx_meta_param& p = node.params[i];
        for (int q = 0; q < spec_sources(p); ++q)
            if (p.src[q].kind != 0) {
                if ((int)p.src[q].idx >= n) {
                    err = "Parameter " + std::to_string(i) + " depends on out-of-range parameter " + std::to_string(p.src[q].idx);
                    return false;
                }
                dep_counts[i]++;
                rev[p.src[q].idx].push_back(i);
            }
    }
    std::deque<int> queue;
    for (int i = 0; i < n; ++i) {
        values[i] = 0.0f;
        if (dep_counts[i] == 0) queue.push_back(i);
    }
    int done = 0;
    while (!queue.empty()) {
        const int i = queue.front();
        queue.pop_front();
        values[i] = sample_spec(node.params[i], values, rng, (discrete_mask >> i) & 1u);
        done++;
        for (int r : rev[i])
            if (--dep_counts[r] == 0) queue.push_back(r);
    }
    if (done != n) {
        err = "Cycle in parameter dependencies";
        return false;
    }
    return true;
}

// ---- compile state ----------------------------------------------------------------------------------------------------
struct Instance {
    int shape = -1;  // -1 none (Points), else IVX_META_SPHERES / CAPSULES / BOXES
    f32 sp[6] = {0, 0, 0, 0, 0, 0};  // the kind's parameters in declaration order (already scaled)
    Sim transform;
};
enum OutKind { OUT_SDF = 0, OUT_GROUP = 1, OUT_INSTANCES = 2 };
struct Output {
    OutKind kind = OUT_SDF;
    bool has_sdf = false;
    uint32_t sdf = 0;
    std::vector<uint32_t> group;
    std::vector<Instance> instances;
    const char* label() const { return kind == OUT_SDF ? "sdf" : (kind == OUT_GROUP ? "group" : "instances"); }
};

struct Compiler {
    ivx_ctx* ctx;
    const ivx_meta_node* nodes;
    uint32_t n;
    f32 S;
    uint64_t seed;
    std::vector<ivx_sdf_node> graph;
    std::string err;
    int err_code = IVX_ERR_GRAPH;

    uint32_t add(uint32_t kind, uint32_t c0, uint32_t c1, uint32_t octaves, uint32_t nseed, const f32* p, int np) {
        ivx_sdf_node nd{};
        nd.kind = kind;
        nd.child[0] = c0;
        nd.child[1] = c1;
        nd.octaves = octaves;
        nd.seed = nseed;
        for (int i = 0; i < np; ++i) nd.p[i] = p[i];
        graph.push_back(nd);
        return (uint32_t)graph.size() - 1u;
    }
    uint32_t translation(uint32_t c, V3 t) {
        const f32 p[3] = {t.x, t.y, t.z};
        return add(IVX_TRANSLATION, c, 0, 0, 0, p, 3);
    }
    uint32_t rotation(uint32_t c, Quat q) {
        const f32 p[4] = {q.x, q.y, q.z, q.w};
        return add(IVX_ROTATION, c, 0, 0, 0, p, 4);
    }
    uint32_t scaling(uint32_t c, f32 s) { return add(IVX_SCALING, c, 0, 0, 0, &s, 1); }
    uint32_t combine(uint32_t kind, uint32_t a, uint32_t b, f32 k) { return add(kind, a, b, 0, 0, &k, 1); }

    bool fail(const std::string& m) {
        err = m;
        return false;
    }
    static const char* kind_name(uint32_t k) {
        static const char* names[] = {"Points", "Spheres", "Capsules", "Boxes", "Translation", "Rotation", "Scaling", "Similarity",
                                      "StratifiedGridTransforms", "SphereSurfaceTransforms", "ClosestTranslationToSurface",
                                      "RayTranslationToSurface", "RotationToGradient", "StochasticSelection", "SDFInstantiation",
                                      "TransformApplication", "MultifractalNoiseSDFModifier", "SDFUnion", "SDFSubtraction",
                                      "SDFIntersection", "SDFGroupUnion"};
        return k < 21 ? names[k] : "?";
    }

    // ---- stable seeds (meta.rs:1099-1169) ----
    static uint64_t stable_seed(const ivx_meta_node& nd, const std::vector<uint64_t>& seeds) {
        switch (nd.kind) {
            case IVX_META_POINTS: return splitmix(0x00);
            case IVX_META_SPHERES: return splitmix2(0x01, nd.seed);
            case IVX_META_CAPSULES: return splitmix2(0x02, nd.seed);
            case IVX_META_BOXES: return splitmix2(0x03, nd.seed);
            case IVX_META_TRANSLATION: return splitmix3(0x10, nd.seed, seeds[nd.child[0]]);
            case IVX_META_ROTATION: return splitmix3(0x11, nd.seed, seeds[nd.child[0]]);
            case IVX_META_SCALING: return splitmix3(0x12, nd.seed, seeds[nd.child[0]]);
            case IVX_META_SIMILARITY: return splitmix3(0x13, nd.seed, seeds[nd.child[0]]);
            case IVX_META_STRATIFIED_GRID_TRANSFORMS: return splitmix3(0x14, nd.seed, seeds[nd.child[0]]);
            case IVX_META_SPHERE_SURFACE_TRANSFORMS: return splitmix3(0x15, nd.seed, seeds[nd.child[0]]);
            case IVX_META_STOCHASTIC_SELECTION: return splitmix3(0x30, nd.seed, seeds[nd.child[0]]);
            case IVX_META_MULTIFRACTAL_NOISE: return splitmix3(0x50, nd.seed, seeds[nd.child[0]]);
            case IVX_META_SDF_INSTANTIATION: return splitmix2(0x40, seeds[nd.child[0]]);
            case IVX_META_SDF_GROUP_UNION: return splitmix2(0x63, seeds[nd.child[0]]);
            case IVX_META_CLOSEST_TRANSLATION_TO_SURFACE: return splitmix3(0x20, seeds[nd.child[0]], seeds[nd.child[1]]);
            case IVX_META_RAY_TRANSLATION_TO_SURFACE: return splitmix3(0x21, seeds[nd.child[0]], seeds[nd.child[1]]);
            case IVX_META_ROTATION_TO_GRADIENT: return splitmix3(0x22, seeds[nd.child[0]], seeds[nd.child[1]]);
            case IVX_META_TRANSFORM_APPLICATION: return splitmix3(0x41, seeds[nd.child[0]], seeds[nd.child[1]]);
            case IVX_META_SDF_SUBTRACTION: return splitmix3(0x61, seeds[nd.child[0]], seeds[nd.child[1]]);
            default: {  // commutative: SDFUnion 0x60, SDFIntersection 0x62
                const uint64_t s1 = seeds[nd.child[0]], s2 = seeds[nd.child[1]];
                return splitmix3(nd.kind == IVX_META_SDF_UNION ? 0x60 : 0x62, std::min(s1, s2), std::max(s1, s2));
            }
        }
    }
    static int n_children(uint32_t kind) {
        if (kind <= IVX_META_BOXES) return 0;
        switch (kind) {
            case IVX_META_CLOSEST_TRANSLATION_TO_SURFACE:
            case IVX_META_RAY_TRANSLATION_TO_SURFACE:
            case IVX_META_ROTATION_TO_GRADIENT:
            case IVX_META_TRANSFORM_APPLICATION:
            case IVX_META_SDF_UNION:
            case IVX_META_SDF_SUBTRACTION:
            case IVX_META_SDF_INTERSECTION: return 2;
            default: return 1;
        }
    }

    bool instances_of(const Output& o, const char* name, const std::vector<Instance>*& out) {
        if (o.kind != OUT_INSTANCES) return fail(std::string(name) + " node expects Instances as input, got " + o.label());
        out = &o.instances;
        return true;
    }

    // unit_quaternion_from_tilt_turn_roll (meta.rs:2810-2834)
    static Quat tilt_turn_roll(f32 tilt_deg, f32 turn_deg, f32 roll_deg) {
        const f32 d2r = PI_F / 180.0f;
        const f32 polar = tilt_deg * d2r, azim = turn_deg * d2r, roll = roll_deg * d2r;
        const f32 sp = sin_f(polar), cp = cos_f(polar), sa = sin_f(azim), ca = cos_f(azim);
        const V3 direction = v3(sp * ca, cp, sp * sa);
        const Quat without_roll = quat_from_rotation_arc(v3(0, 1, 0), direction);
        return quat_mul(quat_from_axis_angle(direction, roll), without_roll);
    }
    // compute_uniformly_distributed_radial_directions (impact_geometry/src/lib.rs:59-87)
    static std::vector<V3> radial_directions(size_t count) {
        const f32 idx_norm = 1.0f / (count > 1 ? (f32)(count - 1) : 1.0f);
        const f32 golden = PI_F * (3.0f - std::sqrt(5.0f));
        std::vector<V3> out;
        for (size_t i = 0; i < count; ++i) {
            const f32 fi = (f32)i;
            const f32 z = 1.0f - 2.0f * fi * idx_norm;
            const f32 hr = std::sqrt(1.0f - z * z);
            const f32 az = fi * golden;
            const V3 v = v3(hr * cos_f(az), hr * sin_f(az), z);
            out.push_back(v / norm(v));
        }
        return out;
    }
    // compute_jittered_direction (meta.rs:2772-2808)
    static V3 jittered_direction(V3 direction, f32 max_angle, Rng& rng) {
        if (std::fabs(max_angle) <= F32_EPS) return direction;
        const f32 angle = rng.f32_in_range(0.0f, max_angle);
        V3 axis;
        axis.x = rng.f32_in_range(-1.0f, 1.0f);
        axis.y = rng.f32_in_range(-1.0f, 1.0f);
        axis.z = rng.f32_in_range(-1.0f, 1.0f);
        axis = axis - dot(axis, direction) * direction;
        const f32 n2 = dot(axis, axis);
        if (n2 > 1e-8f * 1e-8f) {
            axis = axis / std::sqrt(n2);
        } else {
            axis = std::fabs(direction.z) < 0.9f ? v3(0, 0, 1) : v3(1, 0, 0);
            axis = axis - dot(axis, direction) * direction;
            axis = axis / norm(axis);
        }
        return quat_rotate(quat_from_axis_angle(axis, angle), direction);
    }

    template <typename Make>
    bool per_instance(const ivx_meta_node& nd, const std::vector<Output>& outs, uint64_t rseed, int n_params, Make make, Output& out) {
        const std::vector<Instance>* inst;
        if (!instances_of(outs[nd.child[0]], kind_name(nd.kind), inst)) return false;
        Rng rng(rseed);
        const bool per = nd.sampling == 1;
        f32 p[IVX_META_MAX_PARAMS];
        if (!sample_params(nd, n_params, 0, rng, p, err)) return false;
        out.kind = OUT_INSTANCES;
        for (size_t i = 0; i < inst->size(); ++i) {
            out.instances.push_back(make(p, (*inst)[i]));
            if (per && i + 1 < inst->size())
                if (!sample_params(nd, n_params, 0, rng, p, err)) return false;
        }
        return true;
    }

    // ---- the device side of the probing nodes ----
    struct Probe {
        ivx_ctx* ctx = nullptr;
        ivx_program* prog = nullptr;
        Sim surf;
        f32 dom_lo[3], dom_hi[3];
        ~Probe() {
            if (prog) ivx_program_free(ctx, prog);
        }
        bool eval(const std::vector<V3>& origins, uint32_t size, std::vector<f32>& out, std::string& err) {
            out.assign(origins.size() * size * size * size, 0.0f);
            if (origins.empty()) return true;
            std::vector<f32> o(origins.size() * 3);
            for (size_t i = 0; i < origins.size(); ++i) {
                o[3 * i] = origins[i].x;
                o[3 * i + 1] = origins[i].y;
                o[3 * i + 2] = origins[i].z;
            }
            if (ivx_program_eval_blocks(ctx, prog, o.data(), (uint32_t)origins.size(), size, out.data()) != IVX_OK) {
                err = ivx_last_error(ctx);
                return false;
            }
            return true;
        }
    };
    bool make_probe(uint32_t sdf_id, const char* what, Probe& pr) {
        if (!ctx) {
            err_code = IVX_ERR_INVALID_ARGUMENT;
            return fail(std::string(what) + " needs a device context for its SDF probes");
        }
        pr.ctx = ctx;
        const int rc = ivx_program_build(ctx, graph.data(), (uint32_t)graph.size(), sdf_id, &pr.prog);
        if (rc != IVX_OK) {
            err_code = rc;
            return fail(ivx_last_error(ctx));
        }
        ivx_program_info info;
        ivx_program_info_get(ctx, pr.prog, &info);
        for (int d = 0; d < 3; ++d) {
            pr.dom_lo[d] = info.domain_lo[d];
            pr.dom_hi[d] = info.domain_hi[d];
        }
        const ivx_sdf_node& sn = graph[sdf_id];  // node_to_parent_transform (atomic.rs:1138-1148)
        if (sn.kind == IVX_TRANSLATION) pr.surf.t = v3(sn.p[0], sn.p[1], sn.p[2]);
        else if (sn.kind == IVX_ROTATION) pr.surf.r = Quat{sn.p[0], sn.p[1], sn.p[2], sn.p[3]};
        else if (sn.kind == IVX_SCALING) pr.surf.s = sn.p[0];
        return true;
    }
    // sample_signed_distance_with_gradient (meta.rs:2728-2769) for many positions
    bool sample_with_gradient(Probe& pr, const std::vector<V3>& pos, std::vector<f32>& sd, std::vector<V3>& grad) {
        std::vector<V3> org(pos.size());
        for (size_t i = 0; i < pos.size(); ++i) org[i] = v3(pos[i].x - 0.5f, pos[i].y - 0.5f, pos[i].z - 0.5f);
        std::vector<f32> d;
        if (!pr.eval(org, 2, d, err)) return false;
        sd.resize(pos.size());
        grad.resize(pos.size());
        for (size_t i = 0; i < pos.size(); ++i) {
            const f32* q = &d[8 * i];
            f32 total = 0.0f;
            for (int k = 0; k < 8; ++k) total = total + q[k];
            sd[i] = total * 0.125f;
            const f32 d000 = q[0], d001 = q[1], d010 = q[2], d011 = q[3], d100 = q[4], d101 = q[5], d110 = q[6], d111 = q[7];
            grad[i] = 0.25f * v3((((d100 + d110) + d101) + d111) - (((d000 + d010) + d001) + d011),
                                 (((d010 + d110) + d011) + d111) - (((d000 + d100) + d001) + d101),
                                 (((d001 + d101) + d011) + d111) - (((d000 + d100) + d010) + d110));
        }
        return true;
    }

    bool resolve(const ivx_meta_node& nd, const std::vector<Output>& outs, uint64_t rseed, Output& out);
    bool ray_translation(const ivx_meta_node& nd, const std::vector<Output>& outs, Output& out);
    bool closest_translation(const ivx_meta_node& nd, const std::vector<Output>& outs, Output& out);
    bool rotation_to_gradient(const ivx_meta_node& nd, const std::vector<Output>& outs, Output& out);

    bool build(uint32_t& root, bool& empty) {
        empty = true;
        if (n == 0) return true;
        std::vector<Output> outputs(n);
        std::vector<uint8_t> state(n, 0);
        std::vector<uint64_t> seeds(n, 0);
        std::vector<std::pair<int, uint32_t>> stack;  // (0 visit | 1 process, node)
        stack.push_back({0, n - 1});                  // root = last meta node (meta.rs:766)
        while (!stack.empty()) {
            const auto [op, idx] = stack.back();
            stack.pop_back();
            if (op == 0) {
                if (idx >= n) return fail("Missing meta SDF node " + std::to_string(idx));
                if (state[idx] == 2) continue;
                if (state[idx] == 1) return fail("Detected cycle in meta SDF node graph");
                state[idx] = 1;
                stack.push_back({1, idx});
                for (int c = n_children(nodes[idx].kind) - 1; c >= 0; --c) stack.push_back({0, nodes[idx].child[c]});
            } else {
                state[idx] = 2;
                for (int c = 0; c < n_children(nodes[idx].kind); ++c)
                    if (nodes[idx].child[c] >= n) return fail("Missing meta SDF node " + std::to_string(nodes[idx].child[c]));
                seeds[idx] = stable_seed(nodes[idx], seeds);
                if (!resolve(nodes[idx], outputs, splitmix2(seed, seeds[idx]), outputs[idx])) return false;
            }
        }
        const Output& r = outputs[n - 1];
        if (r.kind != OUT_SDF) return fail("Root meta node must have single SDF output");
        if (!r.has_sdf) return true;  // an empty graph
        empty = false;
        root = r.sdf;
        return true;
    }
};

bool Compiler::resolve(const ivx_meta_node& nd, const std::vector<Output>& outs, uint64_t rseed, Output& out) {
    const f32 eps = F32_EPS;
    switch (nd.kind) {
        case IVX_META_POINTS: {
            out.kind = OUT_INSTANCES;
            out.instances.assign(nd.count, Instance{});
            return true;
        }
        case IVX_META_SPHERES:
        case IVX_META_CAPSULES:
        case IVX_META_BOXES: {
            const int np = nd.kind == IVX_META_SPHERES ? 4 : (nd.kind == IVX_META_CAPSULES ? 5 : 6);
            Rng rng(rseed);
            const bool per = nd.sampling == 1;
            f32 p[IVX_META_MAX_PARAMS];
            if (!sample_params(nd, np, 0, rng, p, err)) return false;
            out.kind = OUT_INSTANCES;
            for (uint32_t i = 0; i < nd.count; ++i) {
                Instance ins;
                ins.shape = (int)nd.kind;
                for (int q = 0; q < np; ++q) ins.sp[q] = p[q] * S;
                out.instances.push_back(ins);
                if (per && i + 1 < nd.count)
                    if (!sample_params(nd, np, 0, rng, p, err)) return false;
            }
            return true;
        }
        case IVX_META_TRANSLATION:
            return per_instance(nd, outs, rseed, 3, [&](const f32* p, const Instance& ins) {
                const V3 tr = v3(p[0] * S, p[1] * S, p[2] * S);
                Instance r = ins;
                r.transform = nd.composition == 0 ? ins.transform.translated(tr) : ins.transform.applied_to_translation(tr);
                return r;
            }, out);
        case IVX_META_ROTATION:
            return per_instance(nd, outs, rseed, 3, [&](const f32* p, const Instance& ins) {
                const Quat q = tilt_turn_roll(p[0], p[1], p[2]);
                Instance r = ins;
                r.transform = nd.composition == 0 ? ins.transform.rotated(q) : ins.transform.applied_to_rotation(q);
                return r;
            }, out);
        case IVX_META_SCALING:
            return per_instance(nd, outs, rseed, 1, [&](const f32* p, const Instance& ins) {
                const f32 s = std::fmax(p[0], eps);
                Instance r = ins;
                r.transform = nd.composition == 0 ? ins.transform.scaled(s) : ins.transform.applied_to_scaling(s);
                return r;
            }, out);
        case IVX_META_SIMILARITY:
            return per_instance(nd, outs, rseed, 7, [&](const f32* p, const Instance& ins) {
                Sim tf;
                tf.s = std::fmax(p[0], eps);
                tf.r = tilt_turn_roll(p[1], p[2], p[3]);
                tf.t = v3(p[4] * S, p[5] * S, p[6] * S);
                Instance r = ins;
                r.transform = nd.composition == 0 ? tf.mul(ins.transform) : ins.transform.mul(tf);
                return r;
            }, out);
        case IVX_META_STRATIFIED_GRID_TRANSFORMS: {
            const std::vector<Instance>* inst;
            if (!instances_of(outs[nd.child[0]], "StratifiedGridTransforms", inst)) return false;
            out.kind = OUT_INSTANCES;
            if (inst->empty()) return true;
            Rng rng(rseed);
            f32 p[IVX_META_MAX_PARAMS];
            if (!sample_params(nd, 7, 0x7u, rng, p, err)) return false;
            const uint64_t shape[3] = {(uint64_t)p[0], (uint64_t)p[1], (uint64_t)p[2]};
            f32 ext[3];
            for (int d = 0; d < 3; ++d) ext[d] = std::fmax(p[3 + d] * S, 0.0f);
            const f32 jf = clampf(p[6], 0.0f, 1.0f);
            const uint64_t cells = shape[0] * shape[1] * shape[2];
            if (cells == 0) {
                out.instances = *inst;
                return true;
            }
            f32 start[3];
            for (int d = 0; d < 3; ++d) start[d] = -0.5f * ((f32)shape[d] * ext[d]) + 0.5f * ext[d];
            for (size_t idx = 0; idx < inst->size(); ++idx) {
                const uint64_t c = (idx * cells) / inst->size();
                const uint64_t ijk[3] = {c / (shape[1] * shape[2]), (c / shape[2]) % shape[1], c % shape[2]};
                f32 pos[3], jit[3];
                for (int d = 0; d < 3; ++d) pos[d] = start[d] + (f32)ijk[d] * ext[d];
                for (int d = 0; d < 3; ++d) jit[d] = rng.f32_in_range(-0.5f, 0.5f) * jf * ext[d];
                Sim tf;
                tf.t = v3(pos[0] + jit[0], pos[1] + jit[1], pos[2] + jit[2]);
                Instance r = (*inst)[idx];
                r.transform = tf.mul((*inst)[idx].transform);
                out.instances.push_back(r);
            }
            return true;
        }
        case IVX_META_SPHERE_SURFACE_TRANSFORMS: {
            const std::vector<Instance>* inst;
            if (!instances_of(outs[nd.child[0]], "SphereSurfaceTransforms", inst)) return false;
            out.kind = OUT_INSTANCES;
            if (inst->empty()) return true;
            Rng rng(rseed);
            f32 p[IVX_META_MAX_PARAMS];
            if (!sample_params(nd, 2, 0, rng, p, err)) return false;
            const f32 radius = std::fmax(p[0] * S, 0.0f);
            const f32 jf = clampf(p[1], 0.0f, 1.0f);
            const size_t count = inst->size();
            const f32 solid = 4.0f * PI_F / (f32)count;
            const f32 x = clampf(1.0f - solid / (2.0f * PI_F), -1.0f, 1.0f);
            const f32 max_angle = clampf(jf * acos_f(x), 0.0f, 0.5f * PI_F);
            const std::vector<V3> dirs = radial_directions(count);
            for (size_t i = 0; i < count; ++i) {
                const V3 jd = jittered_direction(dirs[i], max_angle, rng);
                Quat rot = QID;
                if (nd.rotation == 1) rot = quat_from_rotation_arc(v3(0, 1, 0), jd);
                else if (nd.rotation == 2) rot = quat_from_rotation_arc(v3(0, -1, 0), jd);
                Sim tf;
                tf.t = radius * jd;
                tf.r = rot;
                Instance r = (*inst)[i];
                r.transform = tf.mul((*inst)[i].transform);
                out.instances.push_back(r);
            }
            return true;
        }
        case IVX_META_STOCHASTIC_SELECTION: {
            Rng rng(rseed);
            const uint32_t lo = nd.min_pick_count, hi = std::max(nd.max_pick_count, lo);
            const f32 prob = clampf(nd.pick_probability, 0.0f, 1.0f);
            const Output& in = outs[nd.child[0]];
            out.kind = in.kind;
            if (in.kind == OUT_SDF) {
                if (!in.has_sdf) return true;
                if (lo > 0 && rng.f32_fraction() < prob) {
                    out.has_sdf = true;
                    out.sdf = in.sdf;
                }
                return true;
            }
            const uint32_t count = rng.u32_inclusive(lo, hi);
            // Rng::clone_random_subset_from_slice (impact_math/src/random.rs:58-93): reservoir sampling
            auto subset = [&](auto& dest, const auto& source) {
                const size_t take = std::min<size_t>(count, source.size());
                dest.assign(source.begin(), source.begin() + take);
                if (take == 0 || take >= source.size()) return;
                uint64_t idx = take;
                for (size_t q = take; q < source.size(); ++q) {
                    const uint64_t x = rng.usize_inclusive(0, idx);
                    if (x < take) dest[x] = source[q];
                    idx++;
                }
            };
            if (in.kind == OUT_GROUP) {
                std::vector<uint32_t> sel;
                subset(sel, in.group);
                for (uint32_t s : sel)
                    if (rng.f32_fraction() < prob) out.group.push_back(s);
            } else {
                std::vector<Instance> sel;
                subset(sel, in.instances);
                for (const Instance& s : sel)
                    if (rng.f32_fraction() < prob) out.instances.push_back(s);
            }
            return true;
        }
        case IVX_META_SDF_INSTANTIATION: {
            const std::vector<Instance>* inst;
            if (!instances_of(outs[nd.child[0]], "SDFInstantiation", inst)) return false;
            out.kind = OUT_GROUP;
            for (const Instance& ins : *inst) {
                if (ins.shape < 0) continue;
                uint32_t id;
                V3 center;
                if (ins.shape == IVX_META_SPHERES) {
                    id = add(IVX_SPHERE, 0, 0, 0, 0, &ins.sp[0], 1);
                    center = v3(ins.sp[1], ins.sp[2], ins.sp[3]);
                } else if (ins.shape == IVX_META_CAPSULES) {
                    id = add(IVX_CAPSULE, 0, 0, 0, 0, &ins.sp[0], 2);
                    center = v3(ins.sp[2], ins.sp[3], ins.sp[4]);
                } else {
                    id = add(IVX_BOX, 0, 0, 0, 0, &ins.sp[0], 3);
                    center = v3(ins.sp[3], ins.sp[4], ins.sp[5]);
                }
                const Sim& tf = ins.transform;
                if (std::fabs(center.x) > eps || std::fabs(center.y) > eps || std::fabs(center.z) > eps) id = translation(id, center);
                if (std::fabs(tf.s - 1.0f) > eps) id = scaling(id, tf.s);
                if (std::fabs(tf.r.x) > eps || std::fabs(tf.r.y) > eps || std::fabs(tf.r.z) > eps || std::fabs(tf.r.w - 1.0f) > eps)
                    id = rotation(id, tf.r);
                if (std::fabs(tf.t.x) > eps || std::fabs(tf.t.y) > eps || std::fabs(tf.t.z) > eps) id = translation(id, tf.t);
                out.group.push_back(id);
            }
            return true;
        }
        case IVX_META_TRANSFORM_APPLICATION: {
            const Output& a = outs[nd.child[0]];
            if (a.kind == OUT_INSTANCES) return fail("TransformApplication node expects SingleSDF or GroupSDF as input 1, got Instances");
            std::vector<uint32_t> ids;
            if (a.kind == OUT_SDF) {
                if (a.has_sdf) ids.push_back(a.sdf);
            } else {
                ids = a.group;
            }
            const Output& b = outs[nd.child[1]];
            if (b.kind != OUT_INSTANCES) return fail(std::string("TransformApplication node expects Instances as input 2, got ") + b.label());
            out.kind = OUT_GROUP;
            for (uint32_t sid : ids)
                for (const Instance& ins : b.instances) {
                    const Sim& tf = ins.transform;
                    uint32_t id = sid;
                    if (std::fabs(tf.s - 1.0f) > eps) id = scaling(id, tf.s);
                    if (std::fabs(tf.r.x) > eps || std::fabs(tf.r.y) > eps || std::fabs(tf.r.z) > eps || std::fabs(tf.r.w - 1.0f) > eps)
                        id = rotation(id, tf.r);
                    if (std::fabs(tf.t.x) > eps || std::fabs(tf.t.y) > eps || std::fabs(tf.t.z) > eps) id = translation(id, tf.t);
                    out.group.push_back(id);
                }
            return true;
        }
        case IVX_META_MULTIFRACTAL_NOISE: {
            const Output& in = outs[nd.child[0]];
            struct Drawn {
                f32 p[IVX_META_MAX_PARAMS];
                uint32_t nseed;
            };
            auto draw = [&](Rng& rng, Drawn& d) {
                if (!sample_params(nd, 5, 0x1u, rng, d.p, err)) return false;
                d.nseed = rng.u32_inclusive(0u, 0xFFFFFFFFu);
                return true;
            };
            auto make = [&](const Drawn& d, uint32_t child) {
                const f32 p[4] = {d.p[1] / S, d.p[2], d.p[3], d.p[4] * S};
                return add(IVX_MULTIFRACTAL_NOISE, child, 0, (uint32_t)d.p[0], d.nseed, p, 4);
            };
            if (in.kind == OUT_SDF) {
                out.kind = OUT_SDF;
                if (!in.has_sdf) return true;
                Rng rng(rseed);
                Drawn d;
                if (!draw(rng, d)) return false;
                out.has_sdf = true;
                out.sdf = make(d, in.sdf);
                return true;
            }
            if (in.kind == OUT_GROUP) {
                Rng rng(rseed);
                const bool per = nd.sampling == 1;
                Drawn d;
                if (!draw(rng, d)) return false;
                out.kind = OUT_GROUP;
                for (size_t i = 0; i < in.group.size(); ++i) {
                    out.group.push_back(make(d, in.group[i]));
                    if (per && i + 1 < in.group.size())
                        if (!draw(rng, d)) return false;
                }
                return true;
            }
            return fail("MultifractalNoiseSDFModifier node expects SingleSDF or SDFGroup input, got Instances");
        }
        case IVX_META_SDF_UNION:
        case IVX_META_SDF_SUBTRACTION:
        case IVX_META_SDF_INTERSECTION: {
            const Output &a = outs[nd.child[0]], &b = outs[nd.child[1]];
            if (a.kind != OUT_SDF || b.kind != OUT_SDF)
                return fail(std::string(kind_name(nd.kind)) + " node expects two SingleSDF inputs, got " + a.label() + " and " + b.label());
            const f32 k = std::fmax(nd.smoothness * S, 0.0f);
            out.kind = OUT_SDF;
            if (nd.kind == IVX_META_SDF_UNION) {
                if (!a.has_sdf || !b.has_sdf) {
                    const Output& keep = b.has_sdf ? b : a;  // `a if b is None else b`
                    out.has_sdf = keep.has_sdf;
                    out.sdf = keep.sdf;
                    return true;
                }
                out.has_sdf = true;
                out.sdf = combine(IVX_UNION, a.sdf, b.sdf, k);
            } else if (nd.kind == IVX_META_SDF_SUBTRACTION) {
                if (!a.has_sdf) return true;
                out.has_sdf = true;
                out.sdf = b.has_sdf ? combine(IVX_SUBTRACTION, a.sdf, b.sdf, k) : a.sdf;
            } else {
                if (!a.has_sdf || !b.has_sdf) return true;
                out.has_sdf = true;
                out.sdf = combine(IVX_INTERSECTION, a.sdf, b.sdf, k);
            }
            return true;
        }
        case IVX_META_SDF_GROUP_UNION: {
            const Output& in = outs[nd.child[0]];
            out.kind = OUT_SDF;
            if (in.kind == OUT_SDF) {
                out.has_sdf = in.has_sdf;
                out.sdf = in.sdf;
                return true;
            }
            if (in.kind != OUT_GROUP) return fail("SDFGroupUnion node expects SDFGroup or SingleSDF input, got Instances");
            const f32 k = std::fmax(nd.smoothness * S, 0.0f);
            std::deque<uint32_t> queue(in.group.begin(), in.group.end());  // emit_balanced_binary_tree (meta.rs:2390-2409)
            while (queue.size() > 1) {
                const uint32_t a = queue.front();
                queue.pop_front();
                const uint32_t b = queue.front();
                queue.pop_front();
                queue.push_back(combine(IVX_UNION, a, b, k));
            }
            if (!queue.empty()) {
                out.has_sdf = true;
                out.sdf = queue.front();
            }
            return true;
        }
        case IVX_META_RAY_TRANSLATION_TO_SURFACE: return ray_translation(nd, outs, out);
        case IVX_META_CLOSEST_TRANSLATION_TO_SURFACE: return closest_translation(nd, outs, out);
        case IVX_META_ROTATION_TO_GRADIENT: return rotation_to_gradient(nd, outs, out);
        default: return fail("unknown meta node kind " + std::to_string(nd.kind));
    }
}

// MetaClosestTranslationToSurface::resolve + compute_translation_to_closest_point_on_surface (meta.rs:1620-1688, 2411-2479)
bool Compiler::closest_translation(const ivx_meta_node& nd, const std::vector<Output>& outs, Output& out) {
    const Output& subj = outs[nd.child[1]];
    if (subj.kind != OUT_INSTANCES)
        return fail(std::string("ClosestTranslationToSurface node expects Instances as input, got ") + subj.label());
    const Output& sdf = outs[nd.child[0]];
    if (sdf.kind != OUT_SDF) return fail(std::string("ClosestTranslationToSurface node expects SingleSDF as input 1, got ") + sdf.label());
    out.kind = OUT_INSTANCES;
    if (!sdf.has_sdf || subj.instances.empty()) {
        out.instances = subj.instances;
        return true;
    }
    Probe pr;
    if (!make_probe(sdf.sdf, "ClosestTranslationToSurface", pr)) return false;
    const size_t m = subj.instances.size();
    std::vector<V3> start(m), pos(m);
    for (size_t i = 0; i < m; ++i) pos[i] = start[i] = pr.surf.inverse_transform_point(subj.instances[i].transform.transform_point(v3(0, 0, 0)));
    std::vector<uint8_t> alive(m, 1), active(m, 1);
    for (int it = 0; it < 5; ++it) {  // Newton-Raphson, max_iterations = 5, max_distance_from_surface = 0.1
        std::vector<size_t> idx;
        std::vector<V3> p;
        for (size_t i = 0; i < m; ++i)
            if (active[i]) {
                idx.push_back(i);
                p.push_back(pos[i]);
            }
        if (idx.empty()) break;
        std::vector<f32> sd;
        std::vector<V3> grad;
        if (!sample_with_gradient(pr, p, sd, grad)) return false;
        for (size_t q = 0; q < idx.size(); ++q) {
            const size_t i = idx[q];
            const f32 n2 = dot(grad[q], grad[q]);
            if (std::fabs(n2) <= 1e-8f) {
                alive[i] = active[i] = 0;
                continue;
            }
            pos[i] = pos[i] + (-sd[q] / n2) * grad[q];
            if (std::fabs(sd[q]) <= 0.1f) active[i] = 0;
        }
    }
    for (size_t i = 0; i < m; ++i)
        if (alive[i]) {
            Instance r = subj.instances[i];
            r.transform = r.transform.translated(pr.surf.transform_vector(pos[i] - start[i]));
            out.instances.push_back(r);
        }
    return true;
}

// MetaRotationToGradient::resolve + compute_rotation_to_gradient (meta.rs:1798-1862, 2481-2540)
bool Compiler::rotation_to_gradient(const ivx_meta_node& nd, const std::vector<Output>& outs, Output& out) {
    const Output& subj = outs[nd.child[1]];
    if (subj.kind != OUT_INSTANCES) return fail(std::string("RotationToGradient node expects Instances as input, got ") + subj.label());
    const Output& sdf = outs[nd.child[0]];
    if (sdf.kind != OUT_SDF) return fail(std::string("RotationToGradient node expects SingleSDF as input 1, got ") + sdf.label());
    out.kind = OUT_INSTANCES;
    if (!sdf.has_sdf || subj.instances.empty()) {
        out.instances = subj.instances;
        return true;
    }
    Probe pr;
    if (!make_probe(sdf.sdf, "RotationToGradient", pr)) return false;
    const size_t m = subj.instances.size();
    std::vector<V3> centre(m);
    for (size_t i = 0; i < m; ++i) centre[i] = pr.surf.inverse_transform_point(subj.instances[i].transform.transform_point(v3(0, 0, 0)));
    std::vector<f32> sd;
    std::vector<V3> grad;
    if (!sample_with_gradient(pr, centre, sd, grad)) return false;
    const f32 tiny = 1e-8f * 1e-8f;
    for (size_t i = 0; i < m; ++i) {
        const V3 y_axis = subj.instances[i].transform.transform_vector(v3(0, 1, 0));
        const V3 gp = pr.surf.transform_vector(grad[i]);
        const f32 ny = dot(y_axis, y_axis), ng = dot(gp, gp);
        if (!(ny > tiny && ng > tiny)) continue;
        const Quat q = quat_from_rotation_arc(y_axis / std::sqrt(ny), gp / std::sqrt(ng));
        Instance r = subj.instances[i];
        r.transform = r.transform.rotated(q);
        out.instances.push_back(r);
    }
    return true;
}

// MetaRayTranslationToSurface::resolve + compute_spherecast_translation_to_surface (meta.rs:1690-1796, 2534-2748), all
// instances in lock step
bool Compiler::ray_translation(const ivx_meta_node& nd, const std::vector<Output>& outs, Output& out) {
    const Output& subj = outs[nd.child[1]];
    if (subj.kind != OUT_INSTANCES) return fail(std::string("RayTranslationToSurface node expects Instances as input, got ") + subj.label());
    const Output& sdf = outs[nd.child[0]];
    if (sdf.kind != OUT_SDF) return fail(std::string("RayTranslationToSurface node expects SingleSDF as input 1, got ") + sdf.label());
    out.kind = OUT_INSTANCES;
    if (!sdf.has_sdf) {
        out.instances = subj.instances;
        return true;
    }
    Probe pr;
    if (!make_probe(sdf.sdf, "RayTranslationToSurface", pr)) return false;
    const size_t m = subj.instances.size();
    const bool anchor_shape = nd.anchor == 1;
    std::vector<V3> origins(m), dirs(m), pos(m);
    std::vector<f32> radii(m), t_start(m, 0.0f), t_end(m, 0.0f), dist(m), sd(m, 0.0f);
    std::vector<uint8_t> alive(m), active(m), crossed(m, 0);
    for (size_t i = 0; i < m; ++i) {
        const Instance& ins = subj.instances[i];
        V3 c = v3(0, 0, 0);
        f32 r = 0.0f;
        if (anchor_shape && ins.shape >= 0) {
            if (ins.shape == IVX_META_SPHERES) {
                r = ins.sp[0];
            } else if (ins.shape == IVX_META_CAPSULES) {
                c = v3(0, 0.5f * ins.sp[0], 0);
                r = ins.sp[1];
            } else {
                r = 0.5f * std::fmin(std::fmin(ins.sp[0], ins.sp[1]), ins.sp[2]);
                c = v3(0, 0.5f * ins.sp[1] - r, 0);
            }
        }
        const Sim& tf = ins.transform;
        const V3 cp = tf.transform_point(c);
        const f32 rp = tf.s * r;
        const V3 dp = tf.transform_vector(v3(0, 1, 0));
        origins[i] = pr.surf.inverse_transform_point(cp);
        radii[i] = (1.0f / pr.surf.s) * rp;
        const V3 ds = pr.surf.inverse_transform_vector(dp);
        const f32 n2 = dot(ds, ds);
        const bool ok = n2 > 1e-8f * 1e-8f;
        dirs[i] = ok ? ds / std::sqrt(n2) : v3(0, 1, 0);
        alive[i] = ok ? 1 : 0;
    }
    // ray / domain intersection (axis_aligned_box.rs:420-455)
    for (size_t i = 0; i < m; ++i) {
        if (!alive[i]) continue;
        f32 tmin = 0.0f, tmax = INFINITY;
        bool hit = true;
        for (int d = 0; d < 3; ++d) {
            const f32 dir = comp(dirs[i], d), org = comp(origins[i], d);
            if (dir != 0.0f) {
                const f32 rc = 1.0f / dir;
                const f32 t1 = (pr.dom_lo[d] - org) * rc, t2 = (pr.dom_hi[d] - org) * rc;
                const f32 te = t1 < t2 ? t1 : t2, tx = t1 < t2 ? t2 : t1;
                tmin = std::fmax(tmin, te);
                tmax = std::fmin(tmax, tx);
                if (tmax < tmin) {
                    hit = false;
                    break;
                }
            } else if (org < pr.dom_lo[d] || org > pr.dom_hi[d]) {
                hit = false;
                break;
            }
        }
        if (!hit || tmax < 0.0f) {
            alive[i] = 0;
            continue;
        }
        t_start[i] = std::fmax(tmin, 0.0f) - radii[i];
        t_end[i] = tmax;
    }
    // compute_smallest_signed_distance_on_sphere for the masked instances
    auto smallest_sd = [&](const std::vector<uint8_t>& mask, std::vector<f32>& val, std::vector<uint8_t>& ok) -> bool {
        val.assign(m, 0.0f);
        ok = mask;
        std::vector<size_t> idx;
        for (size_t i = 0; i < m; ++i)
            if (mask[i]) idx.push_back(i);
        if (idx.empty()) return true;
        std::vector<V3> probe(idx.size());
        for (size_t q = 0; q < idx.size(); ++q) probe[q] = pos[idx[q]];
        std::vector<size_t> with_r;
        std::vector<V3> blk;
        for (size_t q = 0; q < idx.size(); ++q)
            if (std::fabs(radii[idx[q]]) > F32_EPS) {
                with_r.push_back(q);
                const V3 p = pos[idx[q]];
                blk.push_back(v3(p.x - 0.5f, p.y - 0.5f, p.z - 0.5f));  // 2x2x2 block around the position
            }
        if (!with_r.empty()) {
            std::vector<f32> d;
            if (!pr.eval(blk, 2, d, err)) return false;
            for (size_t w = 0; w < with_r.size(); ++w) {
                const f32* q8 = &d[8 * w];
                const f32 d000 = q8[0], d001 = q8[1], d010 = q8[2], d011 = q8[3], d100 = q8[4], d101 = q8[5], d110 = q8[6], d111 = q8[7];
                const V3 grad = 0.25f * v3((((d100 + d110) + d101) + d111) - (((d000 + d010) + d001) + d011),
                                           (((d010 + d110) + d011) + d111) - (((d000 + d100) + d001) + d101),
                                           (((d001 + d101) + d011) + d111) - (((d000 + d100) + d010) + d110));
                const f32 n2 = dot(grad, grad);
                const size_t q = with_r[w], i = idx[q];
                if (n2 > 1e-8f * 1e-8f) {
                    const V3 gdir = grad / std::sqrt(n2);
                    probe[q] = pos[i] - radii[i] * gdir;
                } else {
                    const V3 gdir = grad / std::sqrt(1.0f);
                    probe[q] = pos[i] - radii[i] * gdir;
                    ok[i] = 0;
                }
            }
        }
        std::vector<f32> v;
        if (!pr.eval(probe, 1, v, err)) return false;
        for (size_t q = 0; q < idx.size(); ++q) val[idx[q]] = v[q];
        return true;
    };
    for (size_t i = 0; i < m; ++i) {
        dist[i] = t_start[i];
        pos[i] = origins[i] + dist[i] * dirs[i];
    }
    std::vector<f32> val;
    std::vector<uint8_t> ok;
    if (!smallest_sd(alive, val, ok)) return false;
    for (size_t i = 0; i < m; ++i) {
        sd[i] = val[i];
        alive[i] = alive[i] && ok[i];
        if (sd[i] < 0.0f) alive[i] = 0;  // already penetrating: a miss (meta.rs:2646-2650)
        active[i] = alive[i] && std::fabs(sd[i]) > 0.1f;
    }
    int step = 0;
    auto any_active = [&]() {
        for (size_t i = 0; i < m; ++i)
            if (active[i]) return true;
        return false;
    };
    while (any_active()) {
        step++;
        if (step >= 128) {
            for (size_t i = 0; i < m; ++i)
                if (active[i] && !crossed[i]) alive[i] = 0;  // gave up without crossing: a miss
            break;
        }
        for (size_t i = 0; i < m; ++i) {
            if (!active[i]) continue;
            dist[i] = dist[i] + sd[i] * 0.5f;
            if (std::signbit(sd[i])) crossed[i] = 1;
            if (dist[i] > t_end[i] || dist[i] < t_start[i]) {
                alive[i] = 0;
                active[i] = 0;
                continue;
            }
            pos[i] = origins[i] + dist[i] * dirs[i];
        }
        if (!smallest_sd(active, val, ok)) return false;
        for (size_t i = 0; i < m; ++i) {
            if (!active[i]) continue;
            if (!ok[i]) {
                alive[i] = 0;
                active[i] = 0;
                continue;
            }
            sd[i] = val[i];
            if (!(std::fabs(sd[i]) > 0.1f)) active[i] = 0;
        }
    }
    for (size_t i = 0; i < m; ++i) {
        if (!alive[i]) continue;
        const V3 tr_surface = pos[i] - origins[i];
        Instance r = subj.instances[i];
        r.transform = r.transform.translated(pr.surf.transform_vector(tr_surface));
        out.instances.push_back(r);
    }
    return true;
}

}  // namespace

extern "C" int ivx_meta_compile(ivx_ctx* ctx, const ivx_meta_node* nodes, uint32_t n_nodes, float scale_factor, uint64_t seed,
                                ivx_sdf_node* out_nodes, uint32_t capacity, uint32_t* out_count, uint32_t* out_root,
                                int* out_empty, char* err, size_t err_capacity) {
    if ((n_nodes && !nodes) || !out_count || !out_root || !out_empty) return IVX_ERR_INVALID_ARGUMENT;
    *out_count = 0;
    *out_root = 0;
    *out_empty = 1;
    if (err && err_capacity) err[0] = 0;
    for (uint32_t i = 0; i < n_nodes; ++i)
        if (nodes[i].kind > IVX_META_SDF_GROUP_UNION) {
            if (err && err_capacity) std::snprintf(err, err_capacity, "unknown meta node kind %u", nodes[i].kind);
            return IVX_ERR_GRAPH;
        }
    Compiler c{ctx, nodes, n_nodes, scale_factor, seed, {}, {}, IVX_ERR_GRAPH};
    uint32_t root = 0;
    bool empty = true;
    if (!c.build(root, empty)) {
        if (err && err_capacity) std::snprintf(err, err_capacity, "%s", c.err.c_str());
        return c.err_code;
    }
    *out_empty = empty ? 1 : 0;
    *out_count = empty ? 0u : (uint32_t)c.graph.size();
    *out_root = root;
    if (empty) return IVX_OK;
    if (c.graph.size() > capacity) return IVX_ERR_CAPACITY;
    if (!out_nodes) return IVX_ERR_INVALID_ARGUMENT;
    std::memcpy(out_nodes, c.graph.data(), c.graph.size() * sizeof(ivx_sdf_node));
    return IVX_OK;
}
