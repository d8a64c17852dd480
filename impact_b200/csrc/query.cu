// Surface-voxel queries (sm_100a): the voxels collision detection and the interaction code ask a voxel object for.
//
// Replaces
//   VoxelObject::for_each_surface_voxel_in_voxel_ranges              (object/intersection.rs:97-151)
//   VoxelObject::for_each_surface_voxel / _maybe_intersecting_sphere / _capsule
//                                                                    (object/intersection.rs:51-95)
//   Voxel::placement, VoxelFlags::placement                          (lib.rs:330-343, 432-438)
// The reference calls a closure per surface voxel, chunk by chunk (i → j → k) and voxel by voxel (i → j → k) inside the
// included range; here the same sequence is written to an array: a counting pass per chunk, an exclusive scan over the
// chunks of the range (the range index is the visiting order), and an emit pass with a block scan over the 256 columns —
// ballot / prefix-sum stream compaction that keeps the order.
#include "api_internal.cuh"

namespace ivx {

// bits of the 16 voxels of column (ti, tj) that are surface voxels inside [t0, t1): non-empty with fewer than six
// blocked faces
__device__ __forceinline__ uint32_t surface_mask(uint4 fl, uint32_t k0, uint32_t k1) {
    const uint32_t w[4] = {fl.x, fl.y, fl.z, fl.w};
    uint32_t m = 0;
#pragma unroll
    for (uint32_t k = 0; k < 16; ++k) {
        const uint32_t f = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        if (!(f & 1u) && __popc(f & 0xFCu) < 6 && k >= k0 && k < k1) m |= 1u << k;
    }
    return m;
}

// for_each_sphere_voxel_object_contact (collidable.rs:1097-1127): the sphere in the space `transform_to_object_space`
// starts from, every surface voxel as a sphere of radius -signed_distance * extent around its centre carried back into
// that space, determine_sphere_sphere_contact_geometry (impact_physics/src/collision/collidable/sphere.rs:105-136)
// for_each_voxel_object_plane_contact (collidable.rs:1176-1209): corner voxels only, determine_sphere_plane_contact_geometry
// (impact_physics sphere.rs:138-156); for_each_capsule_voxel_object_contact (collidable.rs:1257-1288):
// determine_capsule_sphere_contact_geometry (impact_physics capsule.rs:212-270).
struct SphereContactArgs {
    float q[4], t[3];   // transform_to_object_space
    float center[3], radius;  // sphere: centre, radius | plane: unit normal, displacement | capsule: segment start, radius
    float seg[3];             // capsule: segment vector
    float extent;
    int shape;                // 0 sphere, 1 plane, 2 capsule
};
__device__ __forceinline__ f3 rotate_by(float qx, float qy, float qz, float qw, f3 v) {
    const f3 b = mk3(qx, qy, qz);
    const float b2 = dot3(b, b);
    const float s1 = qw * qw - b2, s2 = dot3(v, b) * 2.0f, s3 = qw * 2.0f;
    const f3 c = mk3(b.y * v.z - b.z * v.y, b.z * v.x - b.x * v.z, b.x * v.y - b.y * v.x);
    return mk3((v.x * s1 + b.x * s2) + c.x * s3, (v.y * s1 + b.y * s2) + c.y * s3, (v.z * s1 + b.z * s2) + c.z * s3);
}
__device__ __forceinline__ bool sphere_contact(const SphereContactArgs& a, uint32_t i, uint32_t j, uint32_t k, int code,
                                               ivx_voxel_contact* out) {
    const float e = a.extent;
    const f3 c_voxel = mk3(((float)i + 0.5f) * e, ((float)j + 0.5f) * e, ((float)k + 0.5f) * e);
    const f3 vc = rotate_by(-a.q[0], -a.q[1], -a.q[2], a.q[3], mk3(c_voxel.x - a.t[0], c_voxel.y - a.t[1], c_voxel.z - a.t[2]));
    const float vr = -sd_decode(code) * e;
    if (a.shape == 1) {
        const f3 n = mk3(a.center[0], a.center[1], a.center[2]);
        const float sd = dot3(n, vc) - a.radius;  // Plane::compute_signed_distance
        const float depth = vr - sd;
        if (depth < 0.0f) return false;
        if (out) {
            out->indices[0] = i;
            out->indices[1] = j;
            out->indices[2] = k;
            out->position[0] = vc.x - sd * n.x;
            out->position[1] = vc.y - sd * n.y;
            out->position[2] = vc.z - sd * n.z;
            out->surface_normal[0] = n.x;
            out->surface_normal[1] = n.y;
            out->surface_normal[2] = n.z;
            out->penetration_depth = depth;
        }
        return true;
    }
    if (a.shape == 2) {
        const f3 s0 = mk3(a.center[0], a.center[1], a.center[2]), sv = mk3(a.seg[0], a.seg[1], a.seg[2]);
        // parameter_of_closest_point_on_line_segment_to_point (impact_geometry line.rs:26-45)
        const float len2 = dot3(sv, sv);
        float tpar = 0.0f;
        if (!(len2 <= 1e-8f)) {
            const f3 sp = mk3(vc.x - s0.x, vc.y - s0.y, vc.z - s0.z);
            tpar = fminf(fmaxf(dot3(sv, sp) / len2, 0.0f), 1.0f);
        }
        const f3 closest = mk3(s0.x + tpar * sv.x, s0.y + tpar * sv.y, s0.z + tpar * sv.z);
        const f3 sdisp = mk3(vc.x - closest.x, vc.y - closest.y, vc.z - closest.z);
        const float sd2 = dot3(sdisp, sdisp);
        const float max_sd = vr + a.radius;
        if (sd2 > max_sd * max_sd) return false;
        if (out) {
            const float sdist = sqrtf(sd2);
            f3 cn;
            float depth;
            if (sdist > 1e-8f) {
                cn = mk3(sdisp.x / sdist, sdisp.y / sdist, sdisp.z / sdist);
                depth = fmaxf(0.0f, max_sd - sdist);
            } else {
                // the voxel's centre lies on the segment: any vector normal to it (glam Vec3A::any_orthogonal_vector)
                const f3 o = fabsf(sv.x) > fabsf(sv.y) ? mk3(-sv.z, 0.0f, sv.x) : mk3(0.0f, sv.z, -sv.y);
                const float on = norm3(o);
                cn = on > 1e-8f ? mk3(o.x / on, o.y / on, o.z / on) : mk3(0.0f, 0.0f, 1.0f);
                depth = fmaxf(0.0f, max_sd);
            }
            const f3 n = mk3(-cn.x, -cn.y, -cn.z);
            out->indices[0] = i;
            out->indices[1] = j;
            out->indices[2] = k;
            out->position[0] = vc.x + vr * n.x;
            out->position[1] = vc.y + vr * n.y;
            out->position[2] = vc.z + vr * n.z;
            out->surface_normal[0] = n.x;
            out->surface_normal[1] = n.y;
            out->surface_normal[2] = n.z;
            out->penetration_depth = depth;
        }
        return true;
    }
    const f3 disp = mk3(a.center[0] - vc.x, a.center[1] - vc.y, a.center[2] - vc.z);
    const float d2 = dot3(disp, disp);
    const float max_d = a.radius + vr;
    if (d2 > max_d * max_d) return false;
    if (out) {
        const float dist = sqrtf(d2);
        const f3 n = dist > 1e-8f ? mk3(disp.x / dist, disp.y / dist, disp.z / dist) : mk3(0.0f, 0.0f, 1.0f);
        out->indices[0] = i;
        out->indices[1] = j;
        out->indices[2] = k;
        out->position[0] = vc.x + vr * n.x;
        out->position[1] = vc.y + vr * n.y;
        out->position[2] = vc.z + vr * n.z;
        out->surface_normal[0] = n.x;
        out->surface_normal[1] = n.y;
        out->surface_normal[2] = n.z;
        out->penetration_depth = fmaxf(0.0f, max_d - dist);
    }
    return true;
}

template <bool EMIT, bool CONTACT>
__global__ void __launch_bounds__(256) k_surface_voxels(const DevChunk* __restrict__ chunks, uint3 nb, AbsorbRange r, uint32_t n_range,
                                                        const unsigned char* __restrict__ voxels, uint32_t* __restrict__ count,
                                                        const uint32_t* __restrict__ first, ivx_surface_voxel* __restrict__ out,
                                                        SphereContactArgs ca, ivx_voxel_contact* __restrict__ out_contacts) {
    __shared__ uint32_t s_warp[8];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15, lane = tid & 31, warp = tid >> 5;
    const uint32_t ek = r.c1[2] - r.c0[2], ej = r.c1[1] - r.c0[1];
    for (uint32_t t = blockIdx.x; t < n_range; t += gridDim.x) {
        const uint32_t ck = r.c0[2] + t % ek, cj = r.c0[1] + (t / ek) % ej, ci = r.c0[0] + t / (ek * ej);
        const DevChunk c = chunks[(ci * nb.y + cj) * nb.z + ck];
        if (c.kind != 2) {  // only non-uniform chunks can have surface voxels
            if (!EMIT && tid == 0) count[t] = 0;
            continue;
        }
        if (EMIT && count[t] == 0) continue;
        const uint32_t gi = ci * 16u + ti, gj = cj * 16u + tj;
        const bool in_ij = gi >= r.v0[0] && gi < r.v1[0] && gj >= r.v0[1] && gj < r.v1[1];
        const uint32_t k_lo = max(ck * 16u, r.v0[2]) - ck * 16u;
        const uint32_t k_hi = min(ck * 16u + 16u, r.v1[2]) - ck * 16u;
        const unsigned char* slot = voxels + (size_t)c.slot * SLOT_BYTES;
        uint32_t m = 0;
        uint4 fl = make_uint4(0, 0, 0, 0), sdw = make_uint4(0, 0, 0, 0);
        if (in_ij && k_lo < k_hi) {
            fl = *reinterpret_cast<const uint4*>(slot + PLANE_FLAGS + tid * 16);
            m = surface_mask(fl, k_lo, k_hi);
            if (CONTACT && m) {
                // keep the surface voxels whose sphere touches the query sphere
                sdw = *reinterpret_cast<const uint4*>(slot + PLANE_SD + tid * 16);
                const uint32_t wsd[4] = {sdw.x, sdw.y, sdw.z, sdw.w};
                const uint32_t wfl[4] = {fl.x, fl.y, fl.z, fl.w};
                for (uint32_t b = m; b; b &= b - 1) {
                    const uint32_t k = (uint32_t)__ffs(b) - 1u;
                    const int code = (int)(int8_t)(wsd[k >> 2] >> (8 * (k & 3)));
                    // against a plane only the corner voxels (at most three blocked faces) make contacts
                    const bool corner = __popc((wfl[k >> 2] >> (8 * (k & 3))) & 0xFCu) < 4;
                    if ((ca.shape == 1 && !corner) || !sphere_contact(ca, gi, gj, ck * 16u + k, code, nullptr)) m &= ~(1u << k);
                }
            }
        }
        const uint32_t mine = __popc(m);
        uint32_t x = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        uint32_t before = x - mine, total = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) before += s_warp[w];
            total += s_warp[w];
        }
        if (!EMIT) {
            if (tid == 0) count[t] = total;
        } else if (CONTACT) {
            if (mine) {
                const uint32_t wsd[4] = {sdw.x, sdw.y, sdw.z, sdw.w};
                ivx_voxel_contact* o = out_contacts + first[t] + before;
                for (uint32_t b = m; b; b &= b - 1) {
                    const uint32_t k = (uint32_t)__ffs(b) - 1u;
                    sphere_contact(ca, gi, gj, ck * 16u + k, (int)(int8_t)(wsd[k >> 2] >> (8 * (k & 3))), o++);
                }
            }
        } else if (mine) {
            const uint4 sd = *reinterpret_cast<const uint4*>(slot + PLANE_SD + tid * 16);
            const uint4 ty = *reinterpret_cast<const uint4*>(slot + PLANE_TYPE + tid * 16);
            const uint32_t ws[4] = {sd.x, sd.y, sd.z, sd.w}, wt[4] = {ty.x, ty.y, ty.z, ty.w}, wf[4] = {fl.x, fl.y, fl.z, fl.w};
            ivx_surface_voxel* o = out + first[t] + before;
            for (uint32_t b = m; b; b &= b - 1) {
                const uint32_t k = (uint32_t)__ffs(b) - 1u;
                const uint32_t f = (wf[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                ivx_surface_voxel v;
                v.indices[0] = gi;
                v.indices[1] = gj;
                v.indices[2] = ck * 16u + k;
                v.voxel.voxel_type = (uint8_t)(wt[k >> 2] >> (8 * (k & 3)));
                v.voxel.signed_distance = (int8_t)(ws[k >> 2] >> (8 * (k & 3)));
                v.voxel.flags = (uint8_t)f;
                const int blocked = __popc(f & 0xFCu);
                v.placement = blocked == 5 ? 0 : (blocked == 4 ? 1 : 2);  // Face, Edge, Corner
                *o++ = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace ivx

extern "C" {

int ivx_object_surface_voxels_in_ranges(ivx_ctx* ctx, const ivx_object* obj, const uint32_t ranges[6], ivx_surface_voxel* out,
                                        size_t capacity, uint64_t* out_count) {
    if (!ctx || !obj || !ranges || !out_count || (!out && capacity)) return IVX_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(ivx_surface_voxel) == 16, "16-byte records");
    cudaSetDevice(ctx->device);
    *out_count = 0;
    if (obj->first_i != 0 || obj->nb[0] != obj->chunk_counts[0])
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "surface voxel queries on a slab-partitioned object are not supported");
    AbsorbRange r{};
    for (int d = 0; d < 3; ++d) {
        if (ranges[2 * d + 1] > obj->chunk_counts[d] * 16u) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "voxel range beyond the grid");
        r.v0[d] = ranges[2 * d];
        r.v1[d] = ranges[2 * d + 1];
        if (r.v0[d] >= r.v1[d]) return IVX_OK;  // any(Range::is_empty)
        r.c0[d] = r.v0[d] / 16;
        r.c1[d] = (r.v1[d] + 15) / 16;
    }
    const uint32_t n_range = (r.c1[0] - r.c0[0]) * (r.c1[1] - r.c0[1]) * (r.c1[2] - r.c0[2]);
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* count = tmp.get<uint32_t>(n_range);
    uint32_t* first = tmp.get<uint32_t>(n_range);
    uint32_t* total = ctx->d_scratch + 52;
    if (!count || !first) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "surface voxels: out of device memory");
    const uint3 nb = make_uint3(obj->nb[0], obj->nb[1], obj->nb[2]);
    const uint32_t grid = ivx_persistent_grid(ctx, n_range, 8);
    KL(ctx, (k_surface_voxels<false, false><<<grid, 256, 0, st>>>(obj->d_chunks, nb, r, n_range, obj->d_voxels, count, nullptr,
                                                                  nullptr, SphereContactArgs{}, nullptr),
             cudaGetLastError()));
    KL(ctx, launch_exclusive_scan(count, first, n_range, total, st));
    uint32_t n = 0;
    if (int rc = ivx_read_words(ctx, total, 1, &n)) return rc;
    *out_count = n;
    if (n == 0) return IVX_OK;
    if ((size_t)n > capacity) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "surface voxels: %u found, room for %zu", n, capacity);
    ivx_surface_voxel* d_out = tmp.get<ivx_surface_voxel>(n);
    if (!d_out) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "surface voxels: out of device memory");
    KL(ctx, (k_surface_voxels<true, false><<<grid, 256, 0, st>>>(obj->d_chunks, nb, r, n_range, obj->d_voxels, count, first, d_out,
                                                                 SphereContactArgs{}, nullptr),
             cudaGetLastError()));
    CU(ctx, cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(ivx_surface_voxel), cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return IVX_OK;
}

// voxel_ranges_in_object_touching_aab (object/intersection.rs:693-700) of a sphere's / capsule's box in normalized
// voxel space, then the query above
static void ranges_touching_box(const ivx_object* obj, const float lo[3], const float hi[3], uint32_t out[6]) {
    for (int d = 0; d < 3; ++d) {
        const float fl = std::fmax(std::floor(lo[d]), 0.0f), ce = std::ceil(hi[d]);
        const uint32_t s = fl >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)fl;
        const uint32_t e = !(ce > 0.0f) ? 0u : (ce >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)ce);
        out[2 * d] = std::max(obj->occ_voxels[d], s);
        out[2 * d + 1] = std::min(obj->occ_voxels[3 + d], e);
    }
}

// the two passes + scan over the chunks of `ranges` with the contact closure `ca`
static int contacts_in_ranges(ivx_ctx* ctx, const ivx_object* obj, const ivx::SphereContactArgs& ca, const uint32_t ranges[6],
                              ivx_voxel_contact* out, size_t capacity, uint64_t* out_count) {
    using namespace ivx;
    AbsorbRange r{};
    for (int d = 0; d < 3; ++d) {
        r.v0[d] = ranges[2 * d];
        r.v1[d] = ranges[2 * d + 1];
        if (r.v0[d] >= r.v1[d]) return IVX_OK;
        r.c0[d] = r.v0[d] / 16;
        r.c1[d] = (r.v1[d] + 15) / 16;
    }
    const uint32_t n_range = (r.c1[0] - r.c0[0]) * (r.c1[1] - r.c0[1]) * (r.c1[2] - r.c0[2]);
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* count = tmp.get<uint32_t>(n_range);
    uint32_t* first = tmp.get<uint32_t>(n_range);
    uint32_t* total = ctx->d_scratch + 52;
    if (!count || !first) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "contacts: out of device memory");
    const uint3 nb = make_uint3(obj->nb[0], obj->nb[1], obj->nb[2]);
    const uint32_t grid = ivx_persistent_grid(ctx, n_range, 8);
    KL(ctx, (k_surface_voxels<false, true><<<grid, 256, 0, st>>>(obj->d_chunks, nb, r, n_range, obj->d_voxels, count, nullptr, nullptr,
                                                                 ca, nullptr),
             cudaGetLastError()));
    KL(ctx, launch_exclusive_scan(count, first, n_range, total, st));
    uint32_t n = 0;
    if (int rc = ivx_read_words(ctx, total, 1, &n)) return rc;
    *out_count = n;
    if (n == 0) return IVX_OK;
    if ((size_t)n > capacity) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "contacts: %u found, room for %zu", n, capacity);
    ivx_voxel_contact* d_out = tmp.get<ivx_voxel_contact>(n);
    if (!d_out) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "contacts: out of device memory");
    KL(ctx, (k_surface_voxels<true, true><<<grid, 256, 0, st>>>(obj->d_chunks, nb, r, n_range, obj->d_voxels, count, first, nullptr,
                                                                ca, d_out),
             cudaGetLastError()));
    CU(ctx, cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(ivx_voxel_contact), cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return IVX_OK;
}

// glam Quat::mul_vec3a operation order, as in k_surface_voxels' rotate_by
static void host_rotate(const float q[4], const float v[3], float out[3]) {
    const float bx = q[0], by = q[1], bz = q[2], w = q[3];
    const float b2 = (bx * bx + by * by) + bz * bz, vb = (v[0] * bx + v[1] * by) + v[2] * bz;
    const float s1 = w * w - b2, s2 = vb * 2.0f, s3 = w * 2.0f;
    const float cx = by * v[2] - bz * v[1], cy = bz * v[0] - bx * v[2], cz = bx * v[1] - by * v[0];
    out[0] = (v[0] * s1 + bx * s2) + cx * s3;
    out[1] = (v[1] * s1 + by * s2) + cy * s3;
    out[2] = (v[2] * s1 + bz * s2) + cz * s3;
}

static int contact_args(ivx_ctx* ctx, const ivx_object* obj, const ivx_isometry* T, ivx::SphereContactArgs& ca) {
    if (obj->first_i != 0 || obj->nb[0] != obj->chunk_counts[0])
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "contact queries on a slab-partitioned object are not supported");
    for (int q = 0; q < 4; ++q) ca.q[q] = T->rotation[q];
    for (int d = 0; d < 3; ++d) ca.t[d] = T->translation[d];
    ca.extent = obj->voxel_extent;
    return IVX_OK;
}

int ivx_object_sphere_contacts(ivx_ctx* ctx, const ivx_object* obj, const ivx_isometry* transform_to_object_space,
                               const float center[3], float radius, ivx_voxel_contact* out, size_t capacity, uint64_t* out_count) {
    if (!ctx || !obj || !transform_to_object_space || !center || !out_count || (!out && capacity) || !(radius >= 0.0f))
        return IVX_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(ivx_voxel_contact) == 40, "ten words");
    cudaSetDevice(ctx->device);
    *out_count = 0;
    ivx::SphereContactArgs ca{};
    if (int rc = contact_args(ctx, obj, transform_to_object_space, ca)) return rc;
    for (int d = 0; d < 3; ++d) ca.center[d] = center[d];
    ca.radius = radius;
    ca.shape = 0;
    // sphere.iso_transformed(transform).scaled(inverse_voxel_extent).compute_aabb() clipped to the occupied ranges
    const float inv_e = 1.0f / obj->voxel_extent;
    float c_obj[3];
    host_rotate(ca.q, center, c_obj);
    const float rn = inv_e * radius;
    float lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        const float cn = inv_e * (c_obj[d] + ca.t[d]);
        lo[d] = cn - rn;
        hi[d] = cn + rn;
    }
    uint32_t ranges[6];
    ranges_touching_box(obj, lo, hi, ranges);
    return contacts_in_ranges(ctx, obj, ca, ranges, out, capacity, out_count);
}

int ivx_object_plane_contacts(ivx_ctx* ctx, const ivx_object* obj, const ivx_isometry* transform_to_object_space,
                              const float unit_normal[3], float displacement, ivx_voxel_contact* out, size_t capacity,
                              uint64_t* out_count) {
    if (!ctx || !obj || !transform_to_object_space || !unit_normal || !out_count || (!out && capacity)) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    *out_count = 0;
    ivx::SphereContactArgs ca{};
    if (int rc = contact_args(ctx, obj, transform_to_object_space, ca)) return rc;
    for (int d = 0; d < 3; ++d) ca.center[d] = unit_normal[d];
    ca.radius = displacement;
    ca.shape = 1;
    // plane.iso_transformed(transform) (plane.rs:197-203): the point normal * displacement and the normal are carried
    // over, the displacement is their dot product; then .scaled(inverse_voxel_extent) and the box of the occupied ranges
    // projected onto the plane's negative halfspace (intersection.rs:30-40, 751-761)
    const float p[3] = {unit_normal[0] * displacement, unit_normal[1] * displacement, unit_normal[2] * displacement};
    float tp[3], tn[3];
    host_rotate(ca.q, p, tp);
    host_rotate(ca.q, unit_normal, tn);
    for (int d = 0; d < 3; ++d) tp[d] = tp[d] + ca.t[d];
    const float td = (tn[0] * tp[0] + tn[1] * tp[1]) + tn[2] * tp[2];
    uint32_t occ[6], ranges[6];
    for (int d = 0; d < 3; ++d) {
        occ[2 * d] = obj->occ_voxels[d];
        occ[2 * d + 1] = obj->occ_voxels[3 + d];
    }
    if (int rc = ivx_voxel_ranges_within_plane(occ, tn, td * (1.0f / obj->voxel_extent), ranges)) return rc;
    return contacts_in_ranges(ctx, obj, ca, ranges, out, capacity, out_count);
}

int ivx_object_capsule_contacts(ivx_ctx* ctx, const ivx_object* obj, const ivx_isometry* transform_to_object_space,
                                const float segment_start[3], const float segment_vector[3], float radius,
                                ivx_voxel_contact* out, size_t capacity, uint64_t* out_count) {
    if (!ctx || !obj || !transform_to_object_space || !segment_start || !segment_vector || !out_count || (!out && capacity) ||
        !(radius >= 0.0f))
        return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    *out_count = 0;
    ivx::SphereContactArgs ca{};
    if (int rc = contact_args(ctx, obj, transform_to_object_space, ca)) return rc;
    for (int d = 0; d < 3; ++d) {
        ca.center[d] = segment_start[d];
        ca.seg[d] = segment_vector[d];
    }
    ca.radius = radius;
    ca.shape = 2;
    // capsule.iso_transformed(transform) (capsule.rs:122-128) .scaled(inverse_voxel_extent) .compute_aabb() (:132-137)
    const float inv_e = 1.0f / obj->voxel_extent;
    float s0[3], sv[3];
    host_rotate(ca.q, segment_start, s0);
    host_rotate(ca.q, segment_vector, sv);
    const float rn = inv_e * radius;
    float lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        const float a = inv_e * (s0[d] + ca.t[d]), v = inv_e * sv[d], b = a + v;
        lo[d] = std::fmin(a - rn, b - rn);
        hi[d] = std::fmax(a + rn, b + rn);
    }
    uint32_t ranges[6];
    ranges_touching_box(obj, lo, hi, ranges);
    return contacts_in_ranges(ctx, obj, ca, ranges, out, capacity, out_count);
}

int ivx_voxel_ranges_within_plane(const uint32_t occupied[6], const float unit_normal[3], float displacement, uint32_t out_ranges[6]) {
    if (!occupied || !unit_normal || !out_ranges) return IVX_ERR_INVALID_ARGUMENT;
    // normalized_aabb_from_voxel_ranges(occupied).projected_onto_negative_halfspace(plane) (axis_aligned_box.rs:460-488)
    const float c[2][3] = {{(float)occupied[0], (float)occupied[2], (float)occupied[4]},
                           {(float)occupied[1], (float)occupied[3], (float)occupied[5]}};
    float lo[3] = {c[0][0], c[0][1], c[0][2]}, hi[3] = {c[1][0], c[1][1], c[1][2]};
    const int perm[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
    for (const auto& p : perm) {
        const int i = p[0], j = p[1], k = p[2];
        if (std::fabs(unit_normal[k]) > 1e-8f) {
            const float a = unit_normal[i] * c[0][i] + unit_normal[j] * c[0][j];
            const float b = unit_normal[i] * c[0][i] + unit_normal[j] * c[1][j];
            const float cc = unit_normal[i] * c[1][i] + unit_normal[j] * c[0][j];
            const float d = unit_normal[i] * c[1][i] + unit_normal[j] * c[1][j];
            const float extremal = (displacement - std::fmin(std::fmin(std::fmin(a, b), cc), d)) / unit_normal[k];
            if (!std::signbit(unit_normal[k])) {
                lo[k] = std::fmin(lo[k], extremal);
                hi[k] = std::fmin(hi[k], extremal);
            } else {
                lo[k] = std::fmax(lo[k], extremal);
                hi[k] = std::fmax(hi[k], extremal);
            }
        }
    }
    // voxel_ranges_touching_aab (object/intersection.rs:766-782)
    for (int d = 0; d < 3; ++d) {
        const float fl = std::fmax(std::floor(lo[d]), 0.0f), ce = std::ceil(hi[d]);
        const uint32_t s = fl >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)fl;
        const uint32_t e = !(ce > 0.0f) ? 0u : (ce >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)ce);
        out_ranges[2 * d] = std::max(occupied[2 * d], s);
        out_ranges[2 * d + 1] = std::min(occupied[2 * d + 1], e);
    }
    return IVX_OK;
}

int ivx_object_surface_voxels_within_plane(ivx_ctx* ctx, const ivx_object* obj, const float unit_normal[3], float displacement,
                                           ivx_surface_voxel* out, size_t capacity, uint64_t* out_count) {
    if (!ctx || !obj || !unit_normal) return IVX_ERR_INVALID_ARGUMENT;
    uint32_t occ[6], ranges[6];
    for (int d = 0; d < 3; ++d) {
        occ[2 * d] = obj->occ_voxels[d];
        occ[2 * d + 1] = obj->occ_voxels[3 + d];
    }
    if (int rc = ivx_voxel_ranges_within_plane(occ, unit_normal, displacement, ranges)) return rc;
    return ivx_object_surface_voxels_in_ranges(ctx, obj, ranges, out, capacity, out_count);
}

int ivx_object_surface_voxels_touching_sphere(ivx_ctx* ctx, const ivx_object* obj, const float center[3], float radius,
                                              ivx_surface_voxel* out, size_t capacity, uint64_t* out_count) {
    if (!ctx || !obj || !center || !(radius >= 0.0f)) return IVX_ERR_INVALID_ARGUMENT;
    float lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        lo[d] = center[d] - radius;
        hi[d] = center[d] + radius;
    }
    uint32_t ranges[6];
    ranges_touching_box(obj, lo, hi, ranges);
    return ivx_object_surface_voxels_in_ranges(ctx, obj, ranges, out, capacity, out_count);
}

int ivx_object_surface_voxels_touching_capsule(ivx_ctx* ctx, const ivx_object* obj, const float segment_start[3],
                                               const float segment_vector[3], float radius, ivx_surface_voxel* out,
                                               size_t capacity, uint64_t* out_count) {
    if (!ctx || !obj || !segment_start || !segment_vector || !(radius >= 0.0f)) return IVX_ERR_INVALID_ARGUMENT;
    float lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {  // Capsule::compute_aabb (capsule.rs:132-137)
        const float a = segment_start[d], b = segment_start[d] + segment_vector[d];
        lo[d] = std::fmin(a - radius, b - radius);
        hi[d] = std::fmax(a + radius, b + radius);
    }
    uint32_t ranges[6];
    ranges_touching_box(obj, lo, hi, ranges);
    return ivx_object_surface_voxels_in_ranges(ctx, obj, ranges, out, capacity, out_count);
}

}  // extern "C"
