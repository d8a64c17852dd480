// Host-side graph compile: atomic SDF node list → linear post-order program.
//
// Replaces SDFGraph::build_in → SDFGenerator::new_in and
// determine_transforms_and_margins
// (engine/crates/impact_voxel/src/generation/sdf/atomic.rs:1031-1037, 228-493,
// 495-596). The output is the ProcessedSDFNode list the device kernels
// interpret; it must be identical, float for float, to what the reference
// computes, so every expression keeps the reference's f32 operation order
// (glam SSE2 paths, no FMA contraction: build with -ffp-contract=off).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "ivx_internal.h"

namespace ivx {
namespace {

struct F3 {
    float v[3];
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
};
struct Box3 {
    F3 lo, hi;
};

inline F3 f3(float a, float b, float c) { return F3{{a, b, c}}; }
inline F3 splat(float a) { return F3{{a, a, a}}; }
inline F3 add(F3 a, F3 b) { return f3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline F3 sub(F3 a, F3 b) { return f3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline F3 scale(float s, F3 a) { return f3(s * a[0], s * a[1], s * a[2]); }
inline F3 cmin(F3 a, F3 b) { return f3(std::fmin(a[0], b[0]), std::fmin(a[1], b[1]), std::fmin(a[2], b[2])); }
inline F3 cmax(F3 a, F3 b) { return f3(std::fmax(a[0], b[0]), std::fmax(a[1], b[1]), std::fmax(a[2], b[2])); }
inline float dot3(F3 a, F3 b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline bool neg_bit(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return (u >> 31) != 0;
}
inline Box3 grow(const Box3& b, float m) { return Box3{sub(b.lo, splat(m)), add(b.hi, splat(m))}; }

// glam quat_to_axes (Mat3A::from_quat / Mat4::from_quat)
void quat_axes(const float q[4], F3& ax, F3& ay, F3& az) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
    float wx = w * x2, wy = w * y2, wz = w * z2;
    ax = f3(1.0f - (yy + zz), xy + wz, xz - wy);
    ay = f3(xy - wz, 1.0f - (xx + zz), yz + wx);
    az = f3(xz + wy, yz - wx, 1.0f - (xx + yy));
}

// AABB of OrientedBox::from_axis_aligned_box(b).rotated(q) (atomic.rs:355-361,
// impact_geometry oriented_box.rs:62-68,189-214): rotate the centre with
// Quat::mul_vec3a, span the 8 corners with the rotated axes.
Box3 rotated_box_bounds(const Box3& b, const float q[4]) {
    F3 c = scale(0.5f, add(b.lo, b.hi));
    F3 h = scale(0.5f, sub(b.hi, b.lo));
    F3 qv = f3(q[0], q[1], q[2]);
    float w = q[3];
    float b2 = dot3(qv, qv);
    F3 t1 = scale(w * w - b2, c);
    F3 t2 = scale(dot3(c, qv) * 2.0f, qv);
    F3 cr = f3(qv[1] * c[2] - qv[2] * c[1], qv[2] * c[0] - qv[0] * c[2], qv[0] * c[1] - qv[1] * c[0]);
    F3 t3 = scale(w * 2.0f, cr);
    // the reference multiplies vector * scalar; scalar * vector is the same f32 product
    F3 rc = add(add(t1, t2), t3);
    F3 ax, ay, az;
    quat_axes(q, ax, ay, az);
    F3 hw = scale(h[0], ax), hh = scale(h[1], ay), hd = scale(h[2], az);
    Box3 out;
    bool first = true;
    for (int sx = 0; sx < 2; ++sx)
        for (int sy = 0; sy < 2; ++sy)
            for (int sz = 0; sz < 2; ++sz) {
                F3 p = sx ? add(rc, hw) : sub(rc, hw);
                p = sy ? add(p, hh) : sub(p, hh);
                p = sz ? add(p, hd) : sub(p, hd);
                if (first) {
                    out.lo = out.hi = p;
                    first = false;
                } else {
                    out.lo = cmin(out.lo, p);
                    out.hi = cmax(out.hi, p);
                }
            }
    return out;
}

// compiler-rt __powisf2 == Rust f32::powi
float powi_f32(float a, int b) {
    const bool recip = b < 0;
    float r = 1.0f;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}

// soft_combine_domain_padding (atomic.rs:1590-1598)
float combine_padding(float smoothness, uint32_t leaves) {
    float local = 0.25f * smoothness;
    return local * std::log2((float)leaves);
}

struct NodeFacts {
    Box3 domain{{{0, 0, 0}}, {{0, 0, 0}}};
    uint32_t leaves = 0;
    float padding = 0.0f;
    uint8_t state = 0;  // 0 unvisited, 1 on the DFS path, 2 resolved
};

void resolve_facts(const ivx_sdf_node* nodes, uint32_t id, std::vector<NodeFacts>& facts) {
    const ivx_sdf_node& n = nodes[id];
    NodeFacts& f = facts[id];
    const NodeFacts* a = n.kind >= IVX_TRANSLATION ? &facts[n.child[0]] : nullptr;
    const NodeFacts* b = n.kind >= IVX_UNION ? &facts[n.child[1]] : nullptr;
    switch (n.kind) {
        case IVX_SPHERE: {
            F3 h = splat(n.p[0]);
            f.domain = Box3{f3(-h[0], -h[1], -h[2]), h};
            f.leaves = 1;
            break;
        }
        case IVX_CAPSULE: {
            F3 h = splat(n.p[1]);
            h[1] += 0.5f * n.p[0];
            f.domain = Box3{f3(-h[0], -h[1], -h[2]), h};
            f.leaves = 1;
            break;
        }
        case IVX_BOX: {
            F3 h = scale(0.5f, f3(n.p[0], n.p[1], n.p[2]));
            f.domain = Box3{f3(-h[0], -h[1], -h[2]), h};
            f.leaves = 1;
            break;
        }
        case IVX_TRANSLATION: {
            F3 t = f3(n.p[0], n.p[1], n.p[2]);
            f.domain = Box3{add(a->domain.lo, t), add(a->domain.hi, t)};
            f.leaves = a->leaves;
            f.padding = a->padding;
            break;
        }
        case IVX_ROTATION:
            f.domain = rotated_box_bounds(a->domain, n.p);
            f.leaves = a->leaves;
            f.padding = a->padding;
            break;
        case IVX_SCALING:
            f.domain = Box3{scale(n.p[0], a->domain.lo), scale(n.p[0], a->domain.hi)};
            f.leaves = a->leaves;
            f.padding = a->padding;
            break;
        case IVX_MULTIFRACTAL_NOISE:
            f.domain = grow(a->domain, n.p[3]);
            f.leaves = a->leaves;
            f.padding = a->padding;
            break;
        case IVX_UNION:
            f.domain = Box3{cmin(a->domain.lo, b->domain.lo), cmax(a->domain.hi, b->domain.hi)};
            f.leaves = a->leaves + b->leaves;
            f.padding = combine_padding(n.p[0], f.leaves);
            break;
        case IVX_SUBTRACTION:
            f.domain = a->domain;
            f.leaves = a->leaves + b->leaves;
            f.padding = combine_padding(n.p[0], f.leaves);
            break;
        default: {  // intersection
            F3 lo = cmax(a->domain.lo, b->domain.lo);
            F3 hi = cmin(a->domain.hi, b->domain.hi);
            F3 e = sub(hi, lo);
            if (neg_bit(e[0]) || neg_bit(e[1]) || neg_bit(e[2]))
                f.domain = Box3{splat(0.0f), splat(0.0f)};
            else
                f.domain = Box3{lo, hi};
            f.leaves = a->leaves + b->leaves;
            f.padding = combine_padding(n.p[0], f.leaves);
            break;
        }
    }
}

ivx_node lower_node(const ivx_sdf_node& n, const NodeFacts& f) {
    ivx_node o;
    std::memset(&o, 0, sizeof(o));
    o.kind = n.kind;
    o.octaves = n.octaves;
    o.seed = n.seed;
    o.leaf_count = f.leaves;
    switch (n.kind) {
        case IVX_SPHERE: o.p[0] = n.p[0]; break;
        case IVX_CAPSULE:
            o.p[0] = 0.5f * n.p[0];
            o.p[1] = n.p[1];
            break;
        case IVX_BOX:
            for (int i = 0; i < 3; ++i) o.p[i] = 0.5f * n.p[i];
            break;
        case IVX_TRANSLATION:
            for (int i = 0; i < 3; ++i) o.p[i] = n.p[i];
            break;
        case IVX_ROTATION:
            for (int i = 0; i < 4; ++i) o.p[i] = n.p[i];
            break;
        case IVX_SCALING: o.p[0] = n.p[0]; break;
        case IVX_MULTIFRACTAL_NOISE: {
            for (int i = 0; i < 4; ++i) o.p[i] = n.p[i];
            // MultifractalNoiseSDFModifier::new (atomic.rs:1364-1390, 1852-1858)
            float persistence = n.p[2], amplitude = n.p[3];
            float inherent = std::fabs(persistence - 1.0f) > 1e-6f
                                 ? (1.0f - powi_f32(persistence, (int)n.octaves)) / (1.0f - persistence)
                                 : (float)n.octaves;
            o.p[4] = std::fabs(inherent) > 1.1920929e-7f ? amplitude / inherent : 0.0f;
            break;
        }
        default:
            o.p[0] = n.p[0];
            o.p[1] = 0.25f / n.p[0];  // Smoothness::new (sdf.rs:16-21)
            break;
    }
    for (int i = 0; i < 4; ++i) o.transform_to_node_space[5 * i] = 1.0f;
    Box3 padded = grow(f.domain, f.padding);
    for (int i = 0; i < 3; ++i) {
        o.domain_lo[i] = padded.lo[i];
        o.domain_hi[i] = padded.hi[i];
    }
    return o;
}

// determine_transforms_and_margins (atomic.rs:495-596): walk root → leaves,
// carrying the root→node transform and the margin each node must honour.
void assign_transforms_and_margins(std::vector<ivx_node>& prog) {
    struct Frame {
        float m[16];
        float margin;
    };
    std::vector<Frame> frames(prog.size() + 1);
    size_t top = 0;
    std::memset(frames[0].m, 0, sizeof(frames[0].m));
    for (int i = 0; i < 4; ++i) frames[0].m[5 * i] = 1.0f;
    frames[0].margin = 0.02f * 127.0f;  // VoxelSignedDistance::MAX_F32
    for (size_t r = prog.size(); r-- > 0;) {
        ivx_node& n = prog[r];
        Frame cur = frames[top];
        std::memcpy(n.transform_to_node_space, cur.m, sizeof(cur.m));
        n.margin = cur.margin;
        for (int i = 0; i < 3; ++i) {
            n.domain_lo[i] = n.domain_lo[i] - cur.margin;
            n.domain_hi[i] = n.domain_hi[i] + cur.margin;
        }
        float* M = frames[top].m;
        switch (n.kind) {
            case IVX_SPHERE:
            case IVX_CAPSULE:
            case IVX_BOX:
                if (top > 0) top -= 1;
                break;
            case IVX_TRANSLATION:
                for (int i = 0; i < 3; ++i) M[12 + i] += -n.p[i];
                M[15] += 0.0f;
                break;
            case IVX_ROTATION: {
                float qi[4] = {-n.p[0], -n.p[1], -n.p[2], n.p[3]};
                F3 ax, ay, az;
                quat_axes(qi, ax, ay, az);
                float R[16] = {ax[0], ax[1], ax[2], 0, ay[0], ay[1], ay[2], 0, az[0], az[1], az[2], 0, 0, 0, 0, 1};
                float out[16];
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 4; ++i) {
                        float v = R[i] * cur.m[4 * j];
                        v = v + R[4 + i] * cur.m[4 * j + 1];
                        v = v + R[8 + i] * cur.m[4 * j + 2];
                        v = v + R[12 + i] * cur.m[4 * j + 3];
                        out[4 * j + i] = v;
                    }
                std::memcpy(M, out, sizeof(out));
                break;
            }
            case IVX_SCALING: {
                float inv = 1.0f / n.p[0];
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 3; ++i) M[4 * j + i] = inv * M[4 * j + i];
                frames[top].margin = cur.margin / n.p[0];
                break;
            }
            case IVX_MULTIFRACTAL_NOISE: frames[top].margin = cur.margin + n.p[3]; break;
            default: {
                float child_margin = cur.margin + 2.5f * combine_padding(n.p[0], n.leaf_count);
                frames[top].margin = child_margin;
                frames[top + 1] = frames[top];
                std::memcpy(frames[top + 1].m, cur.m, sizeof(cur.m));
                top += 1;
                break;
            }
        }
    }
}

}  // namespace

std::string compile_program(const ivx_sdf_node* nodes, uint32_t n, uint32_t root, HostProgram& out) {
    out = HostProgram{};
    if (n == 0) return "";
    std::vector<NodeFacts> facts(n);
    // Iterative post-order DFS; a node reached again through another parent is
    // emitted again (the DAG is unrolled into a tree), child 1 before child 2.
    struct Item {
        uint32_t id;
        bool emit;
    };
    std::vector<Item> todo;
    todo.push_back({root, false});
    int64_t depth = 0, max_depth = 0;
    while (!todo.empty()) {
        Item it = todo.back();
        todo.pop_back();
        if (!it.emit) {
            if (it.id >= n) return "Missing SDF node " + std::to_string(it.id);
            NodeFacts& f = facts[it.id];
            if (f.state == 1) return "Detected cycle in SDF generator node graph";
            if (f.state == 0) f.state = 1;
            const ivx_sdf_node& node = nodes[it.id];
            if (node.kind > IVX_INTERSECTION) return "Invalid SDF node kind " + std::to_string(node.kind);
            todo.push_back({it.id, true});
            if (node.kind >= IVX_UNION) todo.push_back({node.child[1], false});
            if (node.kind >= IVX_TRANSLATION) todo.push_back({node.child[0], false});
        } else {
            NodeFacts& f = facts[it.id];
            if (f.state != 2) {
                f.state = 2;
                resolve_facts(nodes, it.id, facts);
            }
            out.nodes.push_back(lower_node(nodes[it.id], f));
            uint32_t kind = nodes[it.id].kind;
            if (kind <= IVX_BOX) {
                depth += 1;
                if (depth > max_depth) max_depth = depth;
            } else if (kind >= IVX_UNION) {
                depth -= 1;
            }
        }
    }
    assign_transforms_and_margins(out.nodes);
    out.stack_depth = (uint32_t)max_depth;
    const NodeFacts& rf = facts[root];
    Box3 rd = grow(rf.domain, rf.padding);
    for (int i = 0; i < 3; ++i) {
        out.domain_lo[i] = rd.lo[i];
        out.domain_hi[i] = rd.hi[i];
    }
    return "";
}

}  // namespace ivx

extern "C" int ivx_program_compile_host(const ivx_sdf_node* nodes, uint32_t n_nodes, uint32_t root_node_id,
                                        ivx_node* out_nodes, uint32_t capacity, uint32_t* out_count,
                                        ivx_program_info* out_info, char* err, size_t err_capacity) {
    if ((n_nodes && !nodes) || !out_count) return IVX_ERR_INVALID_ARGUMENT;
    ivx::HostProgram hp;
    std::string e = ivx::compile_program(nodes, n_nodes, root_node_id, hp);
    if (!e.empty()) {
        if (err && err_capacity) {
            std::strncpy(err, e.c_str(), err_capacity - 1);
            err[err_capacity - 1] = 0;
        }
        return IVX_ERR_GRAPH;
    }
    *out_count = (uint32_t)hp.nodes.size();
    if (out_info) {
        out_info->node_count = (uint32_t)hp.nodes.size();
        out_info->stack_depth = hp.stack_depth;
        for (int d = 0; d < 3; ++d) {
            out_info->domain_lo[d] = hp.domain_lo[d];
            out_info->domain_hi[d] = hp.domain_hi[d];
        }
    }
    if (hp.nodes.size() > capacity) return IVX_ERR_CAPACITY;
    if (out_nodes && !hp.nodes.empty()) std::memcpy(out_nodes, hp.nodes.data(), hp.nodes.size() * sizeof(ivx_node));
    return IVX_OK;
}
