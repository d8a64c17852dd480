// Voxel modification kernels (sm_100a): sphere absorption over the touched
// chunk range, per-chunk internal state refresh and mesh invalidation.
//
// Replaces
//   VoxelObject::modify_voxels_within_sphere        (object/intersection.rs:283-394)
//   VoxelObject::modify_voxels_within_capsule       (object/intersection.rs:417-537)
//   Capsule::trim_segment_outside_aab, CapsulePointContainmentTester (impact_geometry/src/capsule.rs:144-250)
//   apply_capsule_absorption's closure              (interaction/absorption.rs:869-888)
//   apply_sphere_absorption's closure               (interaction/absorption.rs:823-843)
//   VoxelAbsorbingSphere::compute_new_signed_distance (absorption.rs:170-179)
//   Voxel::set_signed_distance                      (lib.rs:451-461)
//   update_all_internal_state_and_determine_sparseness (object.rs:2761-2874)
//   handle_chunk_voxels_modified                    (object/intersection.rs:539-598)
// The cross-chunk part of the update (intersection.rs:391-393) reuses the
// boundary kernels in derive.cu with a per-chunk face mask.
#include "common.cuh"
#include "kernels.h"

namespace ivx {

// `x as usize` of a non-negative float, saturating (32-bit indices are plenty: grids are < 2^16 voxels wide)
__device__ __forceinline__ uint32_t sat_u32(float v) {
    return !(v > 0.0f) ? 0u : (v >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)v);
}

// The voxel range of chunk `cc` the modification visits, [t0, t1) per axis; false if the chunk is skipped before
// anything happens to it.
//   sphere:  the object-level touched range clipped to the chunk (intersection.rs:336-345)
//   capsule: the capsule is first trimmed to the chunk's box expanded by the radius
//            (Capsule::trim_segment_outside_aab → AxisAlignedBox::find_contained_subsegment); the range is the
//            trimmed capsule's AABB clipped to the chunk (voxel_ranges_touching_aab), intersection.rs:445-461
__device__ __forceinline__ bool touched_range_in_chunk(const AbsorbShape& s, const AbsorbRange& r, const uint32_t cc[3],
                                                       uint32_t t0[3], uint32_t t1[3]) {
    if (s.capsule != 1) {  // sphere, or the voxel ranges of a mutual absorption (modify_voxels_within_ranges)
        for (int d = 0; d < 3; ++d) {
            t0[d] = max(cc[d] * 16u, r.v0[d]);
            t1[d] = min(cc[d] * 16u + 16u, r.v1[d]);
        }
        return true;
    }
    const float R = s.influence_radius;
    float t_min = 0.0f, t_max = 1.0f;
    for (int d = 0; d < 3; ++d) {
        const float lo = (float)(cc[d] * 16u) - R, hi = (float)((cc[d] + 1u) * 16u) + R;
        const float v = s.seg[d], o = s.center[d];
        if (fabsf(v) > 1e-8f) {
            const float recip = 1.0f / v;
            const float a = (lo - o) * recip, b = (hi - o) * recip;
            const float t_entry = a < b ? a : b, t_exit = a < b ? b : a;
            t_min = fmaxf(t_min, t_entry);
            t_max = fminf(t_max, t_exit);
        } else if (o < lo || o > hi) {
            return false;
        }
    }
    if (!(t_min <= t_max)) return false;
    bool any = true;
    for (int d = 0; d < 3; ++d) {
        const float a = s.center[d] + s.seg[d] * t_min;     // trimmed segment start
        const float b = a + s.seg[d] * (t_max - t_min);     // ... and end
        const float lo = fminf(a - R, b - R), hi = fmaxf(a + R, b + R);  // Capsule::compute_aabb
        t0[d] = max(cc[d] * 16u, sat_u32(fmaxf(floorf(lo), 0.0f)));
        t1[d] = min(cc[d] * 16u + 16u, sat_u32(ceilf(hi)));
        if (t0[d] >= t1[d]) any = false;
    }
    return any;
}

// ---- mutual absorption (interaction/absorption.rs:891-1094) ---------------------------------------------------
// glam Quat::mul_vec3a, the order oracle_math.hpp's quat_rotate uses: v (w^2 - b.b) + b (2 v.b) + (b x v)(2 w)
__device__ __forceinline__ f3 quat_rotate(float qx, float qy, float qz, float qw, f3 v) {
    const f3 b = mk3(qx, qy, qz);
    const float b2 = dot3(b, b);
    const float s1 = qw * qw - b2, s2 = dot3(v, b) * 2.0f, s3 = qw * 2.0f;
    const f3 c = mk3(b.y * v.z - b.z * v.y, b.z * v.x - b.x * v.z, b.x * v.y - b.y * v.x);
    return mk3((v.x * s1 + b.x * s2) + c.x * s3, (v.y * s1 + b.y * s2) + c.y * s3, (v.z * s1 + b.z * s2) + c.z * s3);
}
// evaluate_sdf_from_corner_samples (object/sdf.rs:579-592); d[x << 2 | y << 1 | z]
__device__ __forceinline__ float corner_samples(const float d[8], f3 o) {
    const f3 ro = mk3(1.0f - o.x, 1.0f - o.y, 1.0f - o.z);
    const float d00 = d[0] * ro.x + d[4] * o.x;
    const float d01 = d[1] * ro.x + d[5] * o.x;
    const float d10 = d[2] * ro.x + d[6] * o.x;
    const float d11 = d[3] * ro.x + d[7] * o.x;
    const float d0 = d00 * ro.y + d10 * o.y;
    const float d1 = d01 * ro.y + d11 * o.y;
    return d0 * ro.z + d1 * o.z;
}
// VoxelObject::voxel(i, j, k).signed_distance().to_f32() (object.rs:1050-1062)
__device__ __forceinline__ float other_voxel_sd(const MutualArgs& m, uint32_t i, uint32_t j, uint32_t k) {
    const DevChunk c = m.o_chunks[((i >> 4) * m.o_nb[1] + (j >> 4)) * m.o_nb[2] + (k >> 4)];
    if (c.kind == 0) return sd_decode(127);
    if (c.kind == 1) return sd_decode((int)c.u_sd);
    return sd_decode((int)(int8_t)m.o_voxels[(size_t)c.slot * SLOT_BYTES + PLANE_SD + vidx((int)(i & 15u), (int)(j & 15u), (int)(k & 15u))]);
}
// a floored coordinate as the lower index of an interpolation cell: false when the cell [out, out + 1] is not inside
// [lo, hi_exclusive). `sign_bit`: sample_voxel_object_sdf tests has_negative_component (the sign bit, so -0.0 is
// rejected); the snapshot lookup casts to isize and compares (so -0.0 is index 0)
__device__ __forceinline__ bool lower_index(float f, uint32_t lo, uint32_t hi_exclusive, uint32_t& out, bool sign_bit) {
    if ((sign_bit ? sign_neg(f) : f < 0.0f) || !(f < 4294967040.0f)) return false;
    out = (uint32_t)f;
    return out >= lo && (uint64_t)out + 1u < (uint64_t)hi_exclusive;
}
// sample_voxel_object_sdf (object/sdf.rs:636-675) of object B at a position in B's normalized voxel space
__device__ __forceinline__ float sample_other_sdf(const MutualArgs& m, f3 p) {
    const f3 lc = mk3(p.x - 0.5f, p.y - 0.5f, p.z - 0.5f);
    const f3 li = mk3(floorf(lc.x), floorf(lc.y), floorf(lc.z));
    const f3 fo = lc - li;
    uint32_t i, j, k;
    if (!lower_index(li.x, 0u, m.o_nb[0] * 16u, i, true) || !lower_index(li.y, 0u, m.o_nb[1] * 16u, j, true) ||
        !lower_index(li.z, 0u, m.o_nb[2] * 16u, k, true))
        return sd_decode(127);
    const float d[8] = {other_voxel_sd(m, i, j, k),         other_voxel_sd(m, i, j, k + 1),
                        other_voxel_sd(m, i, j + 1, k),     other_voxel_sd(m, i, j + 1, k + 1),
                        other_voxel_sd(m, i + 1, j, k),     other_voxel_sd(m, i + 1, j, k + 1),
                        other_voxel_sd(m, i + 1, j + 1, k), other_voxel_sd(m, i + 1, j + 1, k + 1)};
    return corner_samples(d, fo);
}
__device__ __forceinline__ size_t snapshot_index(const MutualArgs& m, uint32_t i, uint32_t j, uint32_t k) {
    return ((size_t)(i - m.s0[0]) * (m.s1[1] - m.s0[1]) + (j - m.s0[1])) * (m.s1[2] - m.s0[2]) + (k - m.s0[2]);
}
// compute_subtracted_signed_distance (absorption.rs:1082-1094)
__device__ __forceinline__ float subtracted_sd(float sd, float inside_other, float k, float qik) {
    const float inter = fmaxf(sd, inside_other);
    return op_combine(IVX_SUBTRACTION, sd, inter, k, qik);
}
// The closures of apply_mutual_absorption for the voxel (gi, gj, gk) with distance code `code`: false = not visited
// (maximally outside, or — object B — outside A's snapshot); else `nsd` is the new signed distance.
__device__ __forceinline__ bool mutual_closure(const MutualArgs& m, int mode, uint32_t gi, uint32_t gj, uint32_t gk, int code,
                                               float& nsd) {
    if (code == 127) return false;
    const float sd = sd_decode(code);
    const f3 c = mk3(((float)gi + 0.5f) * m.extent, ((float)gj + 0.5f) * m.extent, ((float)gk + 0.5f) * m.extent);
    if (mode == 2) {
        m.snapshot[snapshot_index(m, gi, gj, gk)] = sd;
        // inverse_voxel_extent_b * transform_from_b_to_a.inverse_transform_point(center_in_a)
        const f3 r = quat_rotate(-m.q[0], -m.q[1], -m.q[2], m.q[3], mk3(c.x - m.t[0], c.y - m.t[1], c.z - m.t[2]));
        const f3 pb = mk3(m.inv_extent_other * r.x, m.inv_extent_other * r.y, m.inv_extent_other * r.z);
        nsd = subtracted_sd(sd, sample_other_sdf(m, pb) * m.dist_scale, m.smoothness, m.qik);
        return true;
    }
    // inverse_voxel_extent_a * transform_from_b_to_a.transform_point(center_in_b)
    const f3 r = quat_rotate(m.q[0], m.q[1], m.q[2], m.q[3], c);
    const f3 pa = mk3(m.inv_extent_other * (r.x + m.t[0]), m.inv_extent_other * (r.y + m.t[1]), m.inv_extent_other * (r.z + m.t[2]));
    const f3 lc = mk3(pa.x - 0.5f, pa.y - 0.5f, pa.z - 0.5f);
    const f3 li = mk3(floorf(lc.x), floorf(lc.y), floorf(lc.z));
    const f3 fo = lc - li;
    uint32_t i, j, k;
    if (!lower_index(li.x, m.s0[0], m.s1[0], i, false) || !lower_index(li.y, m.s0[1], m.s1[1], j, false) ||
        !lower_index(li.z, m.s0[2], m.s1[2], k, false))
        return false;
    const float* s = m.snapshot;
    const float d[8] = {s[snapshot_index(m, i, j, k)],         s[snapshot_index(m, i, j, k + 1)],
                        s[snapshot_index(m, i, j + 1, k)],     s[snapshot_index(m, i, j + 1, k + 1)],
                        s[snapshot_index(m, i + 1, j, k)],     s[snapshot_index(m, i + 1, j, k + 1)],
                        s[snapshot_index(m, i + 1, j + 1, k)], s[snapshot_index(m, i + 1, j + 1, k + 1)]};
    nsd = subtracted_sd(sd, corner_samples(d, fo) * m.dist_scale, m.smoothness, m.qik);
    return true;
}

__global__ void k_absorb_plan(const DevChunk* __restrict__ chunks, uint3 nb, AbsorbRange r, AbsorbShape shape,
                              uint32_t* __restrict__ need_slot, uint32_t n_range) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_range) return;
    const uint32_t ek = r.c1[2] - r.c0[2], ej = r.c1[1] - r.c0[1];
    const uint32_t k = r.c0[2] + t % ek, j = r.c0[1] + (t / ek) % ej, i = r.c0[0] + t / (ek * ej);
    const DevChunk c = chunks[(i * nb.y + j) * nb.z + k];
    const uint32_t cc[3] = {i, j, k};
    uint32_t t0[3], t1[3];
    // only chunks the modification reaches are converted to non-uniform
    const bool reached = touched_range_in_chunk(shape, r, cc, t0, t1);
    need_slot[t] = (reached && c.kind == 1 && c.slot == 0xFFFFFFFFu) ? 1u : 0u;
}

#ifndef IVX_ABSORB_CTAS
#define IVX_ABSORB_CTAS 3  // 80 registers, no spills: the CTAs of a usual absorption are resident at once (one wave)
#endif
__global__ void __launch_bounds__(256, IVX_ABSORB_CTAS) k_absorb_apply(AbsorbArgs a) {
    __shared__ __align__(16) int8_t s_sd[4096];
    __shared__ __align__(16) uint8_t s_fl[4096];
    __shared__ uint32_t s_cnt[8];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    const uint3 nb = a.nb;
    const AbsorbRange r = a.range;
    const uint32_t ek = r.c1[2] - r.c0[2], ej = r.c1[1] - r.c0[1];

    for (uint32_t t = blockIdx.x; t < a.n_range; t += gridDim.x) {
        const uint32_t ck = r.c0[2] + t % ek, cj = r.c0[1] + (t / ek) % ej, ci = r.c0[0] + t / (ek * ej);
        const uint32_t cidx = (ci * nb.y + cj) * nb.z + ck;
        DevChunk me = a.chunks[cidx];
        const uint32_t cc[3] = {ci, cj, ck};
        uint32_t v0[3], v1[3], t0[3], t1[3];
        for (int d = 0; d < 3; ++d) {
            v0[d] = cc[d] * 16u;
            v1[d] = v0[d] + 16u;
        }
        if (!touched_range_in_chunk(a.shape, r, cc, t0, t1)) continue;  // block-uniform
        if (me.kind == 0) continue;
        unsigned char* slot;
        if (me.kind == 1) {
            // convert_to_non_uniform_if_uniform: every uniform chunk of the touched range (intersection.rs:318-331)
            if (me.slot == 0xFFFFFFFFu) me.slot = a.first_new_slot + a.new_slot_ord[t];
            me.kind = 2;
            if (tid == 0 && a.label_stale) a.label_stale[cidx] = 1;  // no region labels yet
            for (int q = 0; q < 6; ++q) me.face[q] = 1;
            me.flags = 0x3F;
            slot = a.voxels + (size_t)me.slot * SLOT_BYTES;
            const uint32_t tw = (uint32_t)me.u_type * 0x01010101u;
            *reinterpret_cast<uint4*>(slot + PLANE_TYPE + tid * 16) = make_uint4(tw, tw, tw, tw);
            *reinterpret_cast<uint4*>(&s_sd[tid * 16]) = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
            *reinterpret_cast<uint4*>(&s_fl[tid * 16]) = make_uint4(0xFCFCFCFCu, 0xFCFCFCFCu, 0xFCFCFCFCu, 0xFCFCFCFCu);
        } else {
            slot = a.voxels + (size_t)me.slot * SLOT_BYTES;
            *reinterpret_cast<uint4*>(&s_sd[tid * 16]) = *reinterpret_cast<const uint4*>(slot + PLANE_SD + tid * 16);
            *reinterpret_cast<uint4*>(&s_fl[tid * 16]) = *reinterpret_cast<const uint4*>(slot + PLANE_FLAGS + tid * 16);
        }
        if (tid < 8) s_cnt[tid] = 0;
        __syncthreads();

        const uint32_t gi = v0[0] + ti, gj = v0[1] + tj;
        uint32_t n_touched = 0, n_emptied = 0, removed_bits = 0;
        if (gi >= t0[0] && gi < t1[0] && gj >= t0[1] && gj < t1[1]) {
            const float px = (float)gi + 0.5f, py = (float)gj + 0.5f;
            for (uint32_t gk = t0[2]; gk < t1[2]; ++gk) {
                const float pz = (float)gk + 0.5f;
                float d2 = 0.0f, nsd_mutual = 0.0f;
                bool inside;
                if (a.shape.capsule >= 2) {
                    inside = mutual_closure(a.mutual, a.shape.capsule, gi, gj, gk, (int)s_sd[vidx(ti, tj, (int)(gk & 15u))], nsd_mutual);
                } else if (a.shape.capsule) {
                    // shortest_squared_distance_from_point_to_segment_if_contained (capsule.rs:225-250): boundary included
                    const f3 c0 = mk3(a.shape.center[0], a.shape.center[1], a.shape.center[2]);
                    const f3 sv = mk3(a.shape.seg[0], a.shape.seg[1], a.shape.seg[2]);
                    const f3 sp = mk3(px, py, pz) - c0;
                    float t = dot3(sp, mk3(a.shape.seg_over_len2[0], a.shape.seg_over_len2[1], a.shape.seg_over_len2[2]));
                    if (t < 0.0f) t = 0.0f;
                    if (t > 1.0f) t = 1.0f;
                    const f3 closest = mk3(c0.x + sv.x * t, c0.y + sv.y * t, c0.z + sv.z * t);
                    const f3 df = mk3(px, py, pz) - closest;
                    d2 = dot3(df, df);
                    inside = d2 <= a.shape.influence_radius_sq;
                } else {
                    // Point3::squared_distance_between(centre, voxel centre); strictly inside
                    const f3 df = mk3(a.shape.center[0] - px, a.shape.center[1] - py, a.shape.center[2] - pz);
                    d2 = dot3(df, df);
                    inside = d2 < a.shape.influence_radius_sq;
                }
                if (inside) {
                    const int idx = vidx(ti, tj, (int)(gk & 15u));
                    const bool was_empty = (s_fl[idx] & 1) != 0;
                    const float sphere_sd = sqrtf(d2) - a.shape.radius;
                    const float nsd = a.shape.capsule >= 2 ? nsd_mutual : fmaxf(sd_decode((int)s_sd[idx]), -sphere_sd);
                    const int code = sd_encode(nsd);
                    s_sd[idx] = (int8_t)code;
                    if (code >= 0) {
                        s_fl[idx] |= 1;
                        if (!was_empty) {
                            n_emptied++;
                            removed_bits |= 1u << (gk & 15u);  // the closure's remove_voxel callback (absorption.rs:836-840)
                        }
                    }
                    n_touched++;
                }
            }
        }
        const int touched = __syncthreads_or(n_touched != 0);
        if (touched) {
            // update_all_internal_state_and_determine_sparseness as a pure function of
            // (previous flags, emptiness): bits of empty voxels that the reference's
            // sweep never writes keep their previous (possibly stale) value.
            uint8_t nf[16];
            uint32_t empty_mask = 0, void_mask = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int idx = vidx(ti, tj, k);
                uint8_t f = s_fl[idx];
                const bool e = (f & 1) != 0;
                if (e) {
                    empty_mask |= 1u << k;
                    if ((int)s_sd[idx] > 100) void_mask |= 1u << k;
                }
                const int up[3] = {ti < 15 ? vidx(ti + 1, tj, k) : -1, tj < 15 ? vidx(ti, tj + 1, k) : -1,
                                   k < 15 ? vidx(ti, tj, k + 1) : -1};
                const int dn[3] = {ti > 0 ? vidx(ti - 1, tj, k) : -1, tj > 0 ? vidx(ti, tj - 1, k) : -1,
                                   k > 0 ? vidx(ti, tj, k - 1) : -1};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const uint8_t ub = (uint8_t)(1u << (5 + d)), db = (uint8_t)(1u << (2 + d));
                    if (up[d] >= 0 && !e) {
                        if (s_fl[up[d]] & 1) f &= (uint8_t)~ub; else f |= ub;
                    }
                    if (dn[d] >= 0) {
                        if (s_fl[dn[d]] & 1) f &= (uint8_t)~db;
                        else if (!e) f |= db;
                    }
                }
                nf[k] = f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) s_fl[vidx(ti, tj, k)] = nf[k];
            const uint32_t ne = __popc(empty_mask);
            if (ti == 0) atomicAdd(&s_cnt[0], ne);
            if (ti == 15) atomicAdd(&s_cnt[1], ne);
            if (tj == 0) atomicAdd(&s_cnt[2], ne);
            if (tj == 15) atomicAdd(&s_cnt[3], ne);
            if (empty_mask & 1u) atomicAdd(&s_cnt[4], 1u);
            if (empty_mask & 0x8000u) atomicAdd(&s_cnt[5], 1u);
            atomicAdd(&s_cnt[6], n_touched);
            atomicAdd(&s_cnt[7], n_emptied);
            const int only_empty = __syncthreads_and(empty_mask == 0xFFFFu);
            const int is_void = __syncthreads_and(void_mask == 0xFFFFu);
            for (int q = 0; q < 6; ++q) me.face[q] = s_cnt[q] == 256u ? 0 : (s_cnt[q] == 0u ? 1 : 2);
            if (only_empty) me.flags |= 1u << 6; else me.flags &= (uint8_t)~(1u << 6);
            if (a.removed_cols) {
                a.removed_cols[(size_t)t * 256 + tid] = (uint16_t)removed_bits;
                if (tid == 0) {
                    a.removed_info[2 * t] = s_cnt[7];
                    a.removed_info[2 * t + 1] = me.slot;  // still valid for reading the types if the chunk is dropped below
                }
            }
            if (is_void) {
                // the chunk is dropped; its slot is orphaned like in the reference (intersection.rs:559-562)
                DevChunk v{};
                v.kind = 0;
                v.slot = 0xFFFFFFFFu;
                me = v;
            }
            if (tid == 0) {
                atomicAdd(&a.stats[0], 1u);
                atomicAdd(&a.stats[1], s_cnt[6]);
                atomicAdd(&a.stats[2], s_cnt[7]);
                if (is_void) atomicAdd(&a.stats[3], 1u);
                // region labels are a function of which voxels are empty
                if (a.label_stale && s_cnt[7] != 0u) a.label_stale[cidx] = 1;
                // invalidated meshes: this chunk and face neighbours whose border band was touched
                a.dirty[cidx] = 1;
                for (int d = 0; d < 3; ++d) {
                    if (cc[d] > 0 && t0[d] - v0[d] < 2u) {
                        uint32_t n[3] = {ci, cj, ck};
                        n[d] -= 1;
                        a.dirty[(n[0] * nb.y + n[1]) * nb.z + n[2]] = 1;
                    }
                    const uint32_t nbd = d == 0 ? nb.x : (d == 1 ? nb.y : nb.z);
                    if (cc[d] + 1 < nbd && v1[d] - t1[d] < 2u) {
                        uint32_t n[3] = {ci, cj, ck};
                        n[d] += 1;
                        a.dirty[(n[0] * nb.y + n[1]) * nb.z + n[2]] = 1;
                    }
                }
            }
        }
        __syncthreads();
        if (me.kind == 2) {
            *reinterpret_cast<uint4*>(slot + PLANE_SD + tid * 16) = *reinterpret_cast<const uint4*>(&s_sd[tid * 16]);
            *reinterpret_cast<uint4*>(slot + PLANE_FLAGS + tid * 16) = *reinterpret_cast<const uint4*>(&s_fl[tid * 16]);
        }
        if (tid == 0) a.chunks[cidx] = me;
        __syncthreads();
    }
}

// face mask for the boundary refresh after a modification: update_mutual_face_adjacencies
// is called for (c, c+x), (c, c+y), (c, c+z) with c in the range `b` (intersection.rs:391-393)
__global__ void k_absorb_face_mask(uint3 nb, AbsorbRange b, uint8_t* __restrict__ face_mask, uint32_t n) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const uint32_t k = c % nb.z, j = (c / nb.z) % nb.y, i = c / (nb.z * nb.y);
    const uint32_t cc[3] = {i, j, k};
    auto in_range = [&](uint32_t x, uint32_t y, uint32_t z) {
        return x >= b.c0[0] && x < b.c1[0] && y >= b.c0[1] && y < b.c1[1] && z >= b.c0[2] && z < b.c1[2];
    };
    uint8_t m = 0;
    if (in_range(i, j, k)) m |= (1u << 1) | (1u << 3) | (1u << 5);  // upper faces
    for (int d = 0; d < 3; ++d) {
        if (cc[d] == 0) continue;
        uint32_t n3[3] = {i, j, k};
        n3[d] -= 1;
        if (in_range(n3[0], n3[1], n3[2])) m |= (uint8_t)(1u << (2 * d));  // lower face, pair owned by the lower chunk
    }
    face_mask[c] = m;
}

// slots for uniform chunks converted by the boundary refresh that have none reserved
__global__ void k_need_slot_for_convert(const DevChunk* __restrict__ chunks, const uint32_t* __restrict__ convert_flag,
                                        uint32_t n, uint32_t* __restrict__ need, uint8_t* __restrict__ label_stale) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    need[c] = (convert_flag[c] && chunks[c].slot == 0xFFFFFFFFu) ? 1u : 0u;
    if (label_stale && convert_flag[c]) label_stale[c] = 1;  // a chunk that becomes NonUniform has no region labels yet
}
// `first_extra`: slots handed out earlier in the same call whose number is only known on the device
__global__ void k_assign_slots(const DevChunk* __restrict__ chunks, const uint32_t* __restrict__ need,
                               const uint32_t* __restrict__ ord, uint32_t first, const uint32_t* __restrict__ first_extra, uint32_t n,
                               uint32_t* __restrict__ slot_of) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    slot_of[c] = need[c] ? first + (first_extra ? *first_extra : 0u) + ord[c] : chunks[c].slot;
}
// number of non-zero bytes (invalidated chunks)
__global__ void k_count_nonzero_u8(const uint8_t* __restrict__ a, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = __ballot_sync(0xffffffffu, c < n && a[c] != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, (uint32_t)__popc(b));
}

// update_occupied_chunk_ranges (object.rs:1156-1182): chunk-level bounds of the chunks that hold a non-empty voxel
__global__ void k_occupied_chunk_ranges(const DevChunk* __restrict__ chunks, uint32_t n, uint3 nb, uint32_t* __restrict__ cmm,
                                        const uint32_t* __restrict__ gate) {
    if (gate && *gate == 0) return;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
    bool any = false;
    if (c < n) {
        const DevChunk ch = chunks[c];
        if (!(ch.kind == 0 || (ch.kind == 2 && (ch.flags & (1u << 6))))) {  // contains_only_empty_voxels
            any = true;
            lo[2] = hi[2] = c % nb.z;
            lo[1] = hi[1] = (c / nb.z) % nb.y;
            lo[0] = hi[0] = c / (nb.z * nb.y);
        }
    }
    if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
        hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&cmm[d], lo[d]);
            atomicMax(&cmm[3 + d], hi[d]);
        }
    }
}

// bounding range of non-empty voxels over the whole object (object.rs:1149-1280). Like the reference, voxels are
// looked at only in the outermost occupied chunk planes (`cmm`, from k_occupied_chunk_ranges): no other chunk can
// hold the first or the last non-empty voxel of an axis. A CTA tests 256 consecutive chunks at once (one per thread);
// its warps then take the few that lie in such a plane, one chunk per warp at a time.
__global__ void __launch_bounds__(256) k_occupied_ranges(const DevChunk* __restrict__ chunks, uint32_t n, uint3 nb,
                                                         uint32_t first_i, const unsigned char* __restrict__ voxels,
                                                         const uint32_t* __restrict__ cmm, uint32_t* __restrict__ occ,
                                                         const uint32_t* __restrict__ gate) {
    __shared__ uint32_t s_list[256];
    __shared__ uint32_t s_count;
    if (gate && *gate == 0) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_count = 0;
    __syncthreads();
    {
        const uint32_t c = blockIdx.x * 256u + tid;
        if (c < n) {
            const DevChunk ch = chunks[c];
            bool take = !(ch.kind == 0 || (ch.kind == 2 && (ch.flags & (1u << 6))));
            if (take && cmm) {
                const uint32_t k = c % nb.z, j = (c / nb.z) % nb.y, i = c / (nb.z * nb.y);
                take = i == cmm[0] || i == cmm[3] || j == cmm[1] || j == cmm[4] || k == cmm[2] || k == cmm[5];
            }
            if (take) s_list[atomicAdd(&s_count, 1u)] = c;
        }
    }
    __syncthreads();
    const uint32_t count = s_count;
    for (uint32_t q = warp; q < count; q += 8u) {
        const uint32_t c = s_list[q];
        const DevChunk ch = chunks[c];
        const uint32_t k = c % nb.z, j = (c / nb.z) % nb.y, i = c / (nb.z * nb.y);
        const uint32_t org[3] = {(i + first_i) * 16u, j * 16u, k * 16u};
        uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
        if (ch.kind == 1) {
            lo[0] = lo[1] = lo[2] = 0u;
            hi[0] = hi[1] = hi[2] = 15u;
        } else {
            // a lane takes the voxel rows lane, lane + 32, ... (row = i * 16 + j: 16 voxels along k, 16 sign bytes)
            const unsigned char* sd = voxels + (size_t)ch.slot * SLOT_BYTES + PLANE_SD;
#pragma unroll
            for (uint32_t t = 0; t < 8u; ++t) {
                const uint32_t row = lane + 32u * t;
                const uint4 w = *reinterpret_cast<const uint4*>(sd + row * 16u);
                const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
                uint32_t nonempty = 0;
#pragma unroll
                for (int b = 0; b < 16; ++b)
                    if ((ws[b >> 2] >> (8 * (b & 3) + 7)) & 1u) nonempty |= 1u << b;
                if (nonempty) {
                    lo[0] = min(lo[0], row >> 4);
                    hi[0] = max(hi[0], row >> 4);
                    lo[1] = min(lo[1], row & 15u);
                    hi[1] = max(hi[1], row & 15u);
                    lo[2] = min(lo[2], (uint32_t)(__ffs(nonempty) - 1));
                    hi[2] = max(hi[2], (uint32_t)(31 - __clz(nonempty)));
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
                hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
            }
        }
        if (lane == 0 && lo[0] != 0xFFFFFFFFu)
            for (int d = 0; d < 3; ++d) {
                atomicMin(&occ[d], org[d] + lo[d]);
                atomicMax(&occ[3 + d], org[d] + hi[d]);
            }
    }
}

cudaError_t launch_absorb_plan(const DevChunk* chunks, const uint32_t nb[3], const AbsorbRange& r, const AbsorbShape& shape,
                               uint32_t* need_slot, uint32_t n_range, cudaStream_t st) {
    if (n_range == 0) return cudaSuccess;
    k_absorb_plan<<<(n_range + 255) / 256, 256, 0, st>>>(chunks, make_uint3(nb[0], nb[1], nb[2]), r, shape, need_slot, n_range);
    return cudaGetLastError();
}
cudaError_t launch_absorb_apply(const AbsorbArgs& a, uint32_t grid, cudaStream_t st) {
    if (a.n_range == 0) return cudaSuccess;
    k_absorb_apply<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_absorb_face_mask(const uint32_t nb[3], const AbsorbRange& b, uint8_t* face_mask, uint32_t n,
                                    cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_absorb_face_mask<<<(n + 255) / 256, 256, 0, st>>>(make_uint3(nb[0], nb[1], nb[2]), b, face_mask, n);
    return cudaGetLastError();
}
cudaError_t launch_need_slot_for_convert(const DevChunk* chunks, const uint32_t* convert_flag, uint32_t n, uint32_t* need,
                                         uint8_t* label_stale, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_need_slot_for_convert<<<(n + 255) / 256, 256, 0, st>>>(chunks, convert_flag, n, need, label_stale);
    return cudaGetLastError();
}
cudaError_t launch_assign_slots(const DevChunk* chunks, const uint32_t* need, const uint32_t* ord, uint32_t first,
                                const uint32_t* first_extra, uint32_t n, uint32_t* slot_of, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_assign_slots<<<(n + 255) / 256, 256, 0, st>>>(chunks, need, ord, first, first_extra, n, slot_of);
    return cudaGetLastError();
}
cudaError_t launch_count_nonzero_u8(const uint8_t* a, uint32_t n, uint32_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_count_nonzero_u8<<<(n + 255) / 256, 256, 0, st>>>(a, n, out);
    return cudaGetLastError();
}
cudaError_t launch_occupied_ranges(const DevChunk* chunks, uint32_t n, const uint32_t nb[3], uint32_t first_i,
                                   const unsigned char* voxels, uint32_t* occ, uint32_t* chunk_minmax_scratch,
                                   const uint32_t* gate, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const uint3 nb3 = make_uint3(nb[0], nb[1], nb[2]);
    if (chunk_minmax_scratch) {
        const uint32_t init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
        cudaError_t e = cudaMemcpyAsync(chunk_minmax_scratch, init, sizeof(init), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        k_occupied_chunk_ranges<<<(n + 255) / 256, 256, 0, st>>>(chunks, n, nb3, chunk_minmax_scratch, gate);
    }
    k_occupied_ranges<<<(n + 255) / 256, 256, 0, st>>>(chunks, n, nb3, first_i, voxels, chunk_minmax_scratch, occ, gate);
    return cudaGetLastError();
}

}  // namespace ivx
