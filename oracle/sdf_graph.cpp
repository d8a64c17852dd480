// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
//
// Restates V/generation/sdf/atomic.rs (V = engine/crates/impact_voxel/src):
//   SDFGenerator::new_in                       atomic.rs:228-493
//   determine_transforms_and_margins           atomic.rs:495-596
//   compute_signed_distances_for_block         atomic.rs:633-875
//   ..._preserving_gradients                   atomic.rs:877-998
//   primitives                                 atomic.rs:1151-1292
//   MultifractalNoiseSDFModifier               atomic.rs:1363-1573
//   update_signed_distances_for_block          atomic.rs:1601-1658
//   26-point block predicates                  atomic.rs:1661-1797
//   combine kernels                            atomic.rs:1800-1848
// and V/generation/sdf.rs:46-102 (hard / smooth CSG operators).
#include <cassert>
#include <cmath>

#include "oracle.hpp"

namespace orc {

// ---- sdf.rs:74-102 ---------------------------------------------------------
static inline float smooth_union(float d1, float d2, float k, float qik) {
    float h = std::fmax(k - std::fabs(d1 - d2), 0.0f);
    return std::fmin(d1, d2) - (h * h) * qik;
}
static inline float op_union(float d1, float d2, float k, float qik) {
    return k == 0.0f ? std::fmin(d1, d2) : smooth_union(d1, d2, k, qik);
}
static inline float op_subtraction(float d1, float d2, float k, float qik) {
    return k == 0.0f ? std::fmax(d1, -d2) : -smooth_union(-d1, d2, k, qik);
}
static inline float op_intersection(float d1, float d2, float k, float qik) {
    return k == 0.0f ? std::fmax(d1, d2) : -smooth_union(-d1, -d2, k, qik);
}
static inline float op_combine(uint32_t kind, float d1, float d2, float k, float qik) {
    switch (kind) {
        case K_UNION: return op_union(d1, d2, k, qik);
        case K_SUBTRACTION: return op_subtraction(d1, d2, k, qik);
        default: return op_intersection(d1, d2, k, qik);
    }
}

// ---- atomic.rs:1590-1598 ---------------------------------------------------
static inline float soft_combine_domain_padding(float smoothness, uint32_t leaf_count) {
    float local_padding = 0.25f * smoothness;
    return local_padding * std::log2((float)leaf_count);
}

// ---- atomic.rs:1852-1858, 1364-1390 ----------------------------------------
static float noise_scale_for(uint32_t octaves, float persistence, float amplitude) {
    float inherent;
    if (std::fabs(persistence - 1.0f) > 1e-6f) {
        inherent = (1.0f - powi(persistence, (int)octaves)) / (1.0f - persistence);
    } else {
        inherent = (float)octaves;
    }
    // abs_diff_ne!(inherent, 0.0) with default epsilon f32::EPSILON
    if (std::fabs(inherent - 0.0f) > 1.1920929e-7f) return amplitude / inherent;
    return 0.0f;
}

static Aabb leaf_domain(const SdfNode& n) {
    switch (n.kind) {
        case K_SPHERE: {
            V3 h = v3s(n.p[0]);
            return Aabb{-h, h};
        }
        case K_CAPSULE: {
            V3 h = v3s(n.p[1]);
            h.y += 0.5f * n.p[0];
            return Aabb{-h, h};
        }
        default: {
            V3 h = 0.5f * v3(n.p[0], n.p[1], n.p[2]);
            return Aabb{-h, h};
        }
    }
}

static ProgNode make_prog_node(const SdfNode& n) {
    ProgNode pn{};
    pn.kind = n.kind;
    pn.octaves = n.octaves;
    pn.seed = n.seed;
    switch (n.kind) {
        case K_SPHERE: pn.p[0] = n.p[0]; break;
        case K_CAPSULE:
            pn.p[0] = 0.5f * n.p[0];
            pn.p[1] = n.p[1];
            break;
        case K_BOX: {
            V3 h = 0.5f * v3(n.p[0], n.p[1], n.p[2]);
            pn.p[0] = h.x;
            pn.p[1] = h.y;
            pn.p[2] = h.z;
            break;
        }
        case K_TRANSLATION:
            pn.p[0] = n.p[0];
            pn.p[1] = n.p[1];
            pn.p[2] = n.p[2];
            break;
        case K_ROTATION:
            for (int i = 0; i < 4; ++i) pn.p[i] = n.p[i];
            break;
        case K_SCALING: pn.p[0] = n.p[0]; break;
        case K_NOISE:
            for (int i = 0; i < 4; ++i) pn.p[i] = n.p[i];
            pn.p[4] = noise_scale_for(n.octaves, n.p[2], n.p[3]);
            break;
        default:
            pn.p[0] = n.p[0];
            pn.p[1] = 0.25f / n.p[0];
            break;
    }
    return pn;
}

static void determine_transforms_and_margins(std::vector<ProgNode>& nodes) {
    size_t n = nodes.size();
    std::vector<M4> tstack(n + 1);
    std::vector<float> mstack(n + 1, 0.0f);
    size_t top = 0;
    tstack[0] = m4_identity();
    mstack[0] = SD_MAX_F32;
    for (size_t r = n; r-- > 0;) {
        ProgNode& node = nodes[r];
        M4 transform = tstack[top];
        float margin = mstack[top];
        std::memcpy(node.transform, transform.c, sizeof(float) * 16);
        node.margin = margin;
        Aabb d{v3(node.dom_lo[0], node.dom_lo[1], node.dom_lo[2]),
               v3(node.dom_hi[0], node.dom_hi[1], node.dom_hi[2])};
        d = aabb_expanded(d, margin);
        node.dom_lo[0] = d.lo.x; node.dom_lo[1] = d.lo.y; node.dom_lo[2] = d.lo.z;
        node.dom_hi[0] = d.hi.x; node.dom_hi[1] = d.hi.y; node.dom_hi[2] = d.hi.z;
        switch (node.kind) {
            case K_SPHERE:
            case K_CAPSULE:
            case K_BOX: top = top > 0 ? top - 1 : 0; break;
            case K_TRANSLATION: {
                // translate_transform(&(-translation)): w_axis += (-t, 0)
                tstack[top].c[3][0] += -node.p[0];
                tstack[top].c[3][1] += -node.p[1];
                tstack[top].c[3][2] += -node.p[2];
                tstack[top].c[3][3] += 0.0f;
                break;
            }
            case K_ROTATION: {
                Quat q{node.p[0], node.p[1], node.p[2], node.p[3]};
                tstack[top] = m4_mul(m4_from_quat(quat_conj(q)), transform);
                break;
            }
            case K_SCALING: {
                float s = 1.0f / node.p[0];
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 3; ++i) tstack[top].c[j][i] = s * tstack[top].c[j][i];
                mstack[top] = margin / node.p[0];
                break;
            }
            case K_NOISE: mstack[top] = margin + node.p[3]; break;
            default: {
                tstack[top + 1] = transform;
                float mc = margin + 2.5f * soft_combine_domain_padding(node.p[0], node.leaf_count);
                mstack[top] = mc;
                mstack[top + 1] = mc;
                top += 1;
                break;
            }
        }
    }
    assert(top == 0);
}

std::string build_generator(const SdfNode* nodes, uint32_t n, uint32_t root, Generator& out) {
    out = Generator{};
    if (n == 0) return "";
    enum State : uint8_t { UNVISITED, VISITING, DETERMINED };
    std::vector<Aabb> domains(n, Aabb{{0, 0, 0}, {0, 0, 0}});
    std::vector<uint32_t> leaf_counts(n, 0);
    std::vector<float> padding(n, 0.0f);
    std::vector<uint8_t> states(n, UNVISITED);
    struct Op {
        bool process;
        uint32_t id;
    };
    std::vector<Op> ops;
    ops.push_back({false, root});
    int64_t stack_top = 0, max_stack_top = 0;

    while (!ops.empty()) {
        Op op = ops.back();
        ops.pop_back();
        uint32_t idx = op.id;
        if (!op.process) {
            if (idx >= n) return "Missing SDF node " + std::to_string(idx);
            if (states[idx] == VISITING) return "Detected cycle in SDF generator node graph";
            if (states[idx] == UNVISITED) states[idx] = VISITING;
            ops.push_back({true, idx});
            const SdfNode& node = nodes[idx];
            if (node.kind > K_INTERSECTION) return "Invalid SDF node kind";
            if (node.kind >= K_UNION) {
                ops.push_back({false, node.child[1]});
                ops.push_back({false, node.child[0]});
            } else if (node.kind >= K_TRANSLATION) {
                ops.push_back({false, node.child[0]});
            }
        } else {
            const SdfNode& node = nodes[idx];
            if (states[idx] != DETERMINED) {
                states[idx] = DETERMINED;
                uint32_t c1 = node.child[0], c2 = node.child[1];
                switch (node.kind) {
                    case K_SPHERE:
                    case K_CAPSULE:
                    case K_BOX:
                        domains[idx] = leaf_domain(node);
                        leaf_counts[idx] = 1;
                        break;
                    case K_TRANSLATION:
                        domains[idx] =
                            aabb_translated(domains[c1], v3(node.p[0], node.p[1], node.p[2]));
                        leaf_counts[idx] = leaf_counts[c1];
                        padding[idx] = padding[c1];
                        break;
                    case K_ROTATION:
                        domains[idx] = aabb_of_rotated_obb(
                            domains[c1], Quat{node.p[0], node.p[1], node.p[2], node.p[3]});
                        leaf_counts[idx] = leaf_counts[c1];
                        padding[idx] = padding[c1];
                        break;
                    case K_SCALING:
                        domains[idx] = aabb_scaled(domains[c1], node.p[0]);
                        leaf_counts[idx] = leaf_counts[c1];
                        padding[idx] = padding[c1];
                        break;
                    case K_NOISE:
                        domains[idx] = aabb_expanded(domains[c1], node.p[3]);
                        leaf_counts[idx] = leaf_counts[c1];
                        padding[idx] = padding[c1];
                        break;
                    case K_UNION:
                        domains[idx] = aabb_from_pair(domains[c1], domains[c2]);
                        leaf_counts[idx] = leaf_counts[c1] + leaf_counts[c2];
                        padding[idx] = soft_combine_domain_padding(node.p[0], leaf_counts[idx]);
                        break;
                    case K_SUBTRACTION:
                        domains[idx] = domains[c1];
                        leaf_counts[idx] = leaf_counts[c1] + leaf_counts[c2];
                        padding[idx] = soft_combine_domain_padding(node.p[0], leaf_counts[idx]);
                        break;
                    default: {
                        Aabb ov;
                        if (!aabb_overlap(domains[c1], domains[c2], ov))
                            ov = Aabb{{0, 0, 0}, {0, 0, 0}};
                        domains[idx] = ov;
                        leaf_counts[idx] = leaf_counts[c1] + leaf_counts[c2];
                        padding[idx] = soft_combine_domain_padding(node.p[0], leaf_counts[idx]);
                        break;
                    }
                }
            }
            Aabb padded = aabb_expanded(domains[idx], padding[idx]);
            ProgNode pn = make_prog_node(node);
            M4 id = m4_identity();
            std::memcpy(pn.transform, id.c, sizeof(float) * 16);
            pn.dom_lo[0] = padded.lo.x; pn.dom_lo[1] = padded.lo.y; pn.dom_lo[2] = padded.lo.z;
            pn.dom_hi[0] = padded.hi.x; pn.dom_hi[1] = padded.hi.y; pn.dom_hi[2] = padded.hi.z;
            pn.margin = 0.0f;
            pn.leaf_count = leaf_counts[idx];
            out.nodes.push_back(pn);
            if (node.kind <= K_BOX) {
                stack_top += 1;
                if (stack_top > max_stack_top) max_stack_top = stack_top;
            } else if (node.kind >= K_UNION) {
                stack_top -= 1;
            }
        }
    }
    determine_transforms_and_margins(out.nodes);
    out.stack_size = (uint32_t)max_stack_top;
    out.domain = aabb_expanded(domains[root], padding[root]);
    return "";
}

// ---- primitives (atomic.rs:1183-1291) --------------------------------------
static inline float sd_sphere(const ProgNode& n, V3 p) { return norm(p) - n.p[0]; }
static inline float sd_capsule(const ProgNode& n, V3 p) {
    float h = n.p[0];
    // f32::clamp(-h, h)
    float c = p.y;
    if (c < -h) c = -h;
    if (c > h) c = h;
    p.y -= c;
    return norm(p) - n.p[1];
}
static inline float sd_box(const ProgNode& n, V3 p) {
    V3 q = vabs(p) - v3(n.p[0], n.p[1], n.p[2]);
    return norm(vmax(q, v3s(0.0f))) + std::fmin(max_component(q), 0.0f);
}
static inline float sd_leaf(const ProgNode& n, V3 p) {
    switch (n.kind) {
        case K_SPHERE: return sd_sphere(n, p);
        case K_CAPSULE: return sd_capsule(n, p);
        default: return sd_box(n, p);
    }
}

static inline M4 node_transform(const ProgNode& n) {
    M4 m;
    std::memcpy(m.c, n.transform, sizeof(float) * 16);
    return m;
}
static inline Aabb node_domain(const ProgNode& n) {
    return Aabb{v3(n.dom_lo[0], n.dom_lo[1], n.dom_lo[2]), v3(n.dom_hi[0], n.dom_hi[1], n.dom_hi[2])};
}

// atomic.rs:1171-1180, 1224-1234, 1279-1285 with margin := -domain_margin
static Aabb leaf_interior_bounds(const ProgNode& n) {
    float m = -n.margin;
    const float FRAC_1_SQRT_3 = 0.577350269189625764509148780501957456f;
    switch (n.kind) {
        case K_SPHERE: {
            V3 h = v3s(n.p[0] * FRAC_1_SQRT_3 + m);
            return Aabb{-h, h};
        }
        case K_CAPSULE: {
            V3 h = v3s(n.p[1] * FRAC_1_SQRT_3 + m);
            h.y += n.p[0];
            return Aabb{-h, h};
        }
        default: {
            V3 h = v3(n.p[0], n.p[1], n.p[2]) + v3s(m);
            return Aabb{-h, h};
        }
    }
}

// update_signed_distances_for_block (atomic.rs:1601-1658)
static void eval_leaf_block(const ProgNode& n, const M4& m, V3 block_origin, int size, float* out) {
    V3 origin = transform_point(m, block_origin);
    V3 dx = col3(m, 0), dy = col3(m, 1), dz = col3(m, 2);
    int idx = 0;
    for (int i = 0; i < size; ++i) {
        V3 opx = origin + (float)i * dx;
        for (int j = 0; j < size; ++j) {
            V3 pos = opx + (float)j * dy;
            for (int k = 0; k < size; ++k) {
                out[idx] = sd_leaf(n, pos);
                pos = pos + dz;
                idx += 1;
            }
        }
    }
}

// The 26 (index, position) block test samples (atomic.rs:1683-1797).
static void block_test_samples(int size, V3 o, V3 dx, V3 dy, V3 dz, int idx_out[26], V3 pos_out[26]) {
    auto flat = [size](int i, int j, int k) { return i * size * size + j * size + k; };
    const int L = size - 1;
    const int H = size / 2;
    float s = (float)(size - 1);
    float h = s * 0.5f;
    int n = 0;
    auto put = [&](int idx, V3 p) {
        idx_out[n] = idx;
        pos_out[n] = p;
        n++;
    };
    // corners
    put(flat(0, 0, 0), o);
    put(flat(L, 0, 0), o + s * dx);
    put(flat(0, L, 0), o + s * dy);
    put(flat(0, 0, L), o + s * dz);
    put(flat(L, L, 0), o + s * (dx + dy));
    put(flat(L, 0, L), o + s * (dx + dz));
    put(flat(0, L, L), o + s * (dy + dz));
    put(flat(L, L, L), o + s * ((dx + dy) + dz));
    // x edges
    put(flat(0, 0, 0), o + h * dx);
    put(flat(0, L, 0), (o + h * dx) + s * dy);
    put(flat(0, 0, L), (o + h * dx) + s * dz);
    put(flat(0, L, L), (o + h * dx) + s * (dy + dz));
    // y edges
    put(flat(0, 0, 0), o + h * dy);
    put(flat(L, 0, 0), (o + h * dy) + s * dx);
    put(flat(0, 0, L), (o + h * dy) + s * dz);
    put(flat(L, 0, L), (o + h * dy) + s * (dx + dz));
    // z edges
    put(flat(0, 0, 0), o + h * dz);
    put(flat(L, 0, 0), (o + h * dz) + s * dx);
    put(flat(0, L, 0), (o + h * dz) + s * dy);
    put(flat(L, L, 0), (o + h * dz) + s * (dx + dy));
    // faces
    put(flat(0, H, H), (o + h * dy) + h * dz);
    put(flat(L, H, H), ((o + s * dx) + h * dy) + h * dz);
    put(flat(H, 0, H), (o + h * dx) + h * dz);
    put(flat(H, L, H), ((o + s * dy) + h * dx) + h * dz);
    put(flat(H, H, 0), (o + h * dx) + h * dy);
    put(flat(H, H, L), ((o + s * dz) + h * dx) + h * dy);
}

struct NoiseFrame {
    V3 origin_for_noise, dxn, dyn, dzn;
    float unscaled_frequency;
    bool rotated;
};
static NoiseFrame noise_frame(const ProgNode& n, const M4& m, V3 block_origin) {
    V3 origin = transform_point(m, block_origin);
    V3 dx = col3(m, 0), dy = col3(m, 1), dz = col3(m, 2);
    float inverse_scale = norm(dx);
    float scale = 1.0f / inverse_scale;
    NoiseFrame f;
    f.unscaled_frequency = n.p[0] * inverse_scale;
    f.origin_for_noise = scale * origin;
    f.dxn = scale * dx;
    f.dyn = scale * dy;
    f.dzn = scale * dz;
    f.rotated = std::fabs(dx.x * inverse_scale - 1.0f) > 1e-6f ||
                std::fabs(dy.y * inverse_scale - 1.0f) > 1e-6f;
    return f;
}
// One `fbm_3d_offset(pos.z,1,pos.y,1,pos.x,1)` sample: simdnoise x := our z.
static inline float noise_at(const ProgNode& n, float freq, V3 pos) {
    return fbm3(pos.z * freq, pos.y * freq, pos.x * freq, n.p[1], n.p[2], n.octaves,
                (int32_t)n.seed);
}

// modify_signed_distances_for_block (atomic.rs:1423-1507)
static void apply_noise_block(const ProgNode& n, const M4& m, V3 block_origin, int size, float* d) {
    NoiseFrame f = noise_frame(n, m, block_origin);
    float noise_scale = n.p[4];
    if (f.rotated) {
        int idx = 0;
        for (int i = 0; i < size; ++i) {
            V3 opx = f.origin_for_noise + (float)i * f.dxn;
            for (int j = 0; j < size; ++j) {
                V3 pos = opx + (float)j * f.dyn;
                for (int k = 0; k < size; ++k) {
                    d[idx] += noise_at(n, f.unscaled_frequency, pos) * noise_scale;
                    pos = pos + f.dzn;
                    idx += 1;
                }
            }
        }
        return;
    }
    // Block call `fbm_3d_offset(o.z, S, o.y, S, o.x, S)`: simdnoise walks its
    // x (our k) as `x_arr[l] = start + l` for one SIMD vector then `+= width`,
    // and its y / z (our j / i) by repeated `+= 1.0` (simdnoise 3.1.x
    // noise_helpers get_3d_noise). Vector width restated as 8 (AVX2).
    const int VW = 8;
    int idx = 0;
    float zc = f.origin_for_noise.x;
    for (int i = 0; i < size; ++i) {
        float yc = f.origin_for_noise.y;
        for (int j = 0; j < size; ++j) {
            for (int k = 0; k < size; ++k) {
                float xc;
                if (size < VW) {
                    xc = f.origin_for_noise.z + (float)k;
                } else {
                    xc = f.origin_for_noise.z + (float)(k % VW);
                    for (int v = 0; v < k / VW; ++v) xc = xc + (float)VW;
                }
                float nv = fbm3(xc * f.unscaled_frequency, yc * f.unscaled_frequency,
                                zc * f.unscaled_frequency, n.p[1], n.p[2], n.octaves,
                                (int32_t)n.seed);
                d[idx] += nv * noise_scale;
                idx += 1;
            }
            yc = yc + 1.0f;
        }
        zc = zc + 1.0f;
    }
}

// all_modified_signed_distances_at_block_test_positions_pass_predicate (atomic.rs:1510-1572)
static bool noise_all_samples_ge_margin(const ProgNode& n, const M4& m, V3 block_origin, int size,
                                        const float* d) {
    NoiseFrame f = noise_frame(n, m, block_origin);
    int idx[26];
    V3 pos[26];
    block_test_samples(size, f.origin_for_noise, f.dxn, f.dyn, f.dzn, idx, pos);
    for (int s = 0; s < 26; ++s) {
        float v = d[idx[s]] + noise_at(n, f.unscaled_frequency, pos[s]) * n.p[4];
        if (!(v >= n.margin)) return false;
    }
    return true;
}

static bool combine_all_samples_ge_margin(const ProgNode& n, int size, const float* d1,
                                          const float* d2) {
    int idx[26];
    V3 pos[26];
    block_test_samples(size, v3s(0), v3s(0), v3s(0), v3s(0), idx, pos);
    for (int s = 0; s < 26; ++s) {
        float v = op_combine(n.kind, d1[idx[s]], d2[idx[s]], n.p[0], n.p[1]);
        if (!(v >= n.margin)) return false;
    }
    return true;
}

static void apply_combine(const ProgNode& n, int count, float* d1, const float* d2) {
    float k = n.p[0], qik = n.p[1];
    for (int i = 0; i < count; ++i) d1[i] = op_combine(n.kind, d1[i], d2[i], k, qik);
}

void eval_chunk(const Generator& g, V3 lo, float* stack, uint8_t* decisions) {
    const int S = CHUNK_SIZE, C = CHUNK_VOXELS;
    if (g.nodes.empty()) {
        for (int i = 0; i < C; ++i) stack[i] = SD_MAX_F32;
        return;
    }
    Aabb block{lo, lo + v3s((float)S)};
    size_t top = 0;
    for (size_t ni = 0; ni < g.nodes.size(); ++ni) {
        const ProgNode& n = g.nodes[ni];
        uint8_t dec = 0;
        switch (n.kind) {
            case K_SPHERE:
            case K_CAPSULE:
            case K_BOX: {
                M4 m = node_transform(n);
                Aabb bn = aabb_of_transformed(block, m);
                float* out = stack + top * C;
                if (aabb_box_lies_outside(node_domain(n), bn)) {
                    for (int i = 0; i < C; ++i) out[i] = n.margin;
                    dec = 1;
                } else if (aabb_contains_box(leaf_interior_bounds(n), bn)) {
                    for (int i = 0; i < C; ++i) out[i] = -n.margin;
                    dec = 2;
                } else {
                    eval_leaf_block(n, m, lo, S, out);
                }
                top += 1;
                break;
            }
            case K_TRANSLATION:
            case K_ROTATION: break;
            case K_SCALING: {
                float* d = stack + (top - 1) * C;
                for (int i = 0; i < C; ++i) d[i] *= n.p[0];
                break;
            }
            case K_NOISE: {
                M4 m = node_transform(n);
                Aabb bn = aabb_of_transformed(block, m);
                float* d = stack + (top - 1) * C;
                if (!aabb_box_lies_outside(node_domain(n), bn) ||
                    !noise_all_samples_ge_margin(n, m, lo, S, d)) {
                    apply_noise_block(n, m, lo, S, d);
                } else {
                    dec = 1;
                }
                break;
            }
            default: {
                top -= 1;
                M4 m = node_transform(n);
                Aabb bn = aabb_of_transformed(block, m);
                float* d1 = stack + (top - 1) * C;
                const float* d2 = stack + top * C;
                if (!aabb_box_lies_outside(node_domain(n), bn) ||
                    !combine_all_samples_ge_margin(n, S, d1, d2)) {
                    apply_combine(n, C, d1, d2);
                } else {
                    dec = 1;
                }
                break;
            }
        }
        if (decisions) decisions[ni] = dec;
    }
    assert(top == 1);
}

void eval_block_preserving_gradients(const Generator& g, V3 origin, int size, float* stack) {
    const int C = size * size * size;
    if (g.nodes.empty()) {
        for (int i = 0; i < C; ++i) stack[i] = SD_MAX_F32;
        return;
    }
    size_t top = 0;
    for (const ProgNode& n : g.nodes) {
        switch (n.kind) {
            case K_SPHERE:
            case K_CAPSULE:
            case K_BOX:
                eval_leaf_block(n, node_transform(n), origin, size, stack + top * C);
                top += 1;
                break;
            case K_TRANSLATION:
            case K_ROTATION: break;
            case K_SCALING: {
                float* d = stack + (top - 1) * C;
                for (int i = 0; i < C; ++i) d[i] *= n.p[0];
                break;
            }
            case K_NOISE:
                apply_noise_block(n, node_transform(n), origin, size, stack + (top - 1) * C);
                break;
            default:
                top -= 1;
                apply_combine(n, C, stack + (top - 1) * C, stack + top * C);
                break;
        }
    }
    assert(top == 1);
}

}  // namespace orc
