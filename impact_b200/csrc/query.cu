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

template <bool EMIT>
__global__ void __launch_bounds__(256) k_surface_voxels(const DevChunk* __restrict__ chunks, uint3 nb, AbsorbRange r, uint32_t n_range,
                                                        const unsigned char* __restrict__ voxels, uint32_t* __restrict__ count,
                                                        const uint32_t* __restrict__ first, ivx_surface_voxel* __restrict__ out) {
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
        uint4 fl = make_uint4(0, 0, 0, 0);
        if (in_ij && k_lo < k_hi) {
            fl = *reinterpret_cast<const uint4*>(slot + PLANE_FLAGS + tid * 16);
            m = surface_mask(fl, k_lo, k_hi);
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
    KL(ctx, (k_surface_voxels<false><<<grid, 256, 0, st>>>(obj->d_chunks, nb, r, n_range, obj->d_voxels, count, nullptr, nullptr),
             cudaGetLastError()));
    KL(ctx, launch_exclusive_scan(count, first, n_range, total, st));
    uint32_t n = 0;
    if (int rc = ivx_read_words(ctx, total, 1, &n)) return rc;
    *out_count = n;
    if (n == 0) return IVX_OK;
    if ((size_t)n > capacity) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "surface voxels: %u found, room for %zu", n, capacity);
    ivx_surface_voxel* d_out = tmp.get<ivx_surface_voxel>(n);
    if (!d_out) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "surface voxels: out of device memory");
    KL(ctx, (k_surface_voxels<true><<<grid, 256, 0, st>>>(obj->d_chunks, nb, r, n_range, obj->d_voxels, count, first, d_out),
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
