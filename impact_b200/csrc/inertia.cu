// Inertial moments of a voxel object (sm_100a): mass, first moments, moments and products of inertia integrated
// over all non-empty voxels, with respect to the origin of the voxel grid.
//
// Replaces
//   VoxelObjectInertialPropertyManager::initialized_from       (object/inertia.rs:125-137)
//   compute_inertial_property_moments_for_object               (object/inertia.rs:754-789)
//   compute_moments_for_non_uniform_chunk / _uniform_chunk     (object/inertia.rs:629-752)
//   VoxelObjectInertialPropertyUpdater::remove_voxel → compute_moments_for_voxel, for the voxels an absorption
//   empties (ivx_apply_removed_voxels, called by ivx_object_absorb_*_inertial)  (object/inertia.rs:374-395, 591-625)
//
// The reference's result is a chain of f32 additions in a fixed order (voxels i → j → k inside a chunk, chunks
// i → j → k over the occupied range), and f32 addition does not reassociate. To return the same bits, the order
// is kept and the parallelism is taken across chains instead of inside them:
//   k_moments_flags + scan every non-void chunk gets a row, rows in linear chunk order (void chunks add nothing to the
//                          reference's sums: they get no row);
//   k_moments_classify     one thread per chunk: closed form for uniform chunks, non-uniform chunks appended to a
//                          work list;
//   k_moments_non_uniform  one thread per non-uniform chunk walks its 4096 voxels in the reference's order with
//                          its ten accumulators in registers (the 2 B/voxel it reads are the kernel's HBM
//                          traffic; ~10^5 independent chains on a 1024^3 object fill the machine);
//   k_moments_sum          the rows are added in order by ten lanes of one warp, one lane per component, from
//                          shared-memory tiles the rest of the block stages ahead of them. The tail of the last tile
//                          is padded with +0.0, which leaves a partial sum that is not -0.0 unchanged — and the
//                          partial sums start at +0.0 (or the caller's values) and only become -0.0 by adding two -0.0;
//                          a caller-provided -0.0 start is kept by not touching the padding at all (the loop stops at
//                          the row count).
#include "api_internal.cuh"

namespace ivx {

struct Densities {
    float v[256];  // voxel_type_densities, zero beyond n
};

struct MomentsArgs {
    const DevChunk* chunks;
    const unsigned char* voxels;
    uint3 nb;           // locally stored chunk planes
    uint32_t first_i;   // global chunk-i of local plane 0
    uint32_t c_begin, c_end;  // local linear chunk range that is summed (the owned planes)
    float e;            // voxel extent
    uint32_t n_densities;
    float* part;        // 10 floats per non-void chunk of [c_begin, c_end), in linear chunk order
    const uint32_t* row;  // per chunk of [c_begin, c_end): its row in `part` (exclusive scan of the non-void flags)
    uint32_t* list;     // non-uniform chunks (local linear index)
    uint32_t* counters; // [0] list length, [1] error: a non-empty voxel type without a density, [2] rows
};

__global__ void k_moments_flags(const DevChunk* __restrict__ chunks, uint32_t c_begin, uint32_t n, uint32_t* __restrict__ flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) flag[t] = chunks[c_begin + t].kind != 0 ? 1u : 0u;
}

// compute_moments_for_uniform_chunk (inertia.rs:710-752)
__device__ __forceinline__ void uniform_chunk_moments(float e, float density, const uint32_t cc[3], float out[10]) {
    const float ce = 16.0f * e;
    float h2[3], h3[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float lo = (float)cc[d] * ce;
        const float hi = lo + ce;
        const float lo2 = lo * lo, hi2 = hi * hi;
        const float lo3 = lo2 * lo, hi3 = hi2 * hi;
        h2[d] = hi2 - lo2;
        h3[d] = hi3 - lo3;
    }
    const float ce2 = ce * ce, ce3 = ce2 * ce;
    const float fm = (0.5f * ce2) * density;
    const float fi = ((1.0f / 3.0f) * ce2) * density;
    const float fp = (0.25f * ce) * density;
    out[0] = ce3 * density;
    out[1] = fm * h2[0];
    out[2] = fm * h2[1];
    out[3] = fm * h2[2];
    out[4] = fi * (h3[1] + h3[2]);
    out[5] = fi * (h3[0] + h3[2]);
    out[6] = fi * (h3[0] + h3[1]);
    out[7] = fp * (h2[0] * h2[1]);
    out[8] = fp * (h2[1] * h2[2]);
    out[9] = fp * (h2[2] * h2[0]);
}

__global__ void __launch_bounds__(256) k_moments_classify(MomentsArgs a, const __grid_constant__ Densities dens) {
    const uint32_t c = a.c_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = c < a.c_end;
    DevChunk ch{};
    if (in) ch = a.chunks[c];
    const bool nu = in && ch.kind == 2;
    // warp-aggregated append (the list order does not reach the result: every term lands in its chunk's row)
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, nu);
    uint32_t base = 0;
    const int lane = threadIdx.x & 31;
    if (m) {
        if (lane == __ffs(m) - 1) base = atomicAdd(&a.counters[0], (uint32_t)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    }
    if (!in) return;
    if (nu) {
        a.list[base + __popc(m & ((1u << lane) - 1u))] = c;
        return;
    }
    if (ch.kind != 1) return;
    float out[10];
    const uint32_t k = c % a.nb.z, j = (c / a.nb.z) % a.nb.y, i = c / (a.nb.z * a.nb.y) + a.first_i;
    const uint32_t cc[3] = {i, j, k};
    if (ch.u_type >= a.n_densities) atomicOr(&a.counters[1], 1u);
    uniform_chunk_moments(a.e, dens.v[ch.u_type], cc, out);
    float* p = a.part + (size_t)a.row[c - a.c_begin] * 10;
#pragma unroll
    for (int q = 0; q < 10; ++q) p[q] = out[q];
}

// compute_moments_for_non_uniform_chunk (inertia.rs:629-706): one thread, one chunk, the reference's loop
__global__ void __launch_bounds__(64) k_moments_non_uniform(MomentsArgs a, const __grid_constant__ Densities dens) {
    __shared__ float s_dens[256];
    for (int q = threadIdx.x; q < 256; q += 64) s_dens[q] = dens.v[q];
    __syncthreads();
    const uint32_t t = blockIdx.x * 64 + threadIdx.x;
    if (t >= a.counters[0]) return;
    const uint32_t c = a.list[t];
    const DevChunk ch = a.chunks[c];
    const uint32_t ck = c % a.nb.z, cj = (c / a.nb.z) % a.nb.y, ci = c / (a.nb.z * a.nb.y) + a.first_i;
    const float e = a.e;
    // position of the lower corner of the chunk's first voxel
    const float x0 = (float)(ci * 16u) * e, y0 = (float)(cj * 16u) * e, z0 = (float)(ck * 16u) * e;
    // the z walk restarts at z0 for every (i, j) column, so its sixteen terms are computed once
    float h2z[16], h3z[16];
    {
        float zl = z0, zh = zl + e;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float zl2 = zl * zl, zh2 = zh * zh;
            const float zl3 = zl2 * zl, zh3 = zh2 * zh;
            h2z[k] = zh2 - zl2;
            h3z[k] = zh3 - zl3;
            zl = zh;
            zh += e;
        }
    }
    const unsigned char* slot = a.voxels + (size_t)ch.slot * SLOT_BYTES;
    const uint4* pf = reinterpret_cast<const uint4*>(slot + PLANE_FLAGS);
    const uint4* pt = reinterpret_cast<const uint4*>(slot + PLANE_TYPE);
    float mass = 0.0f, m0 = 0.0f, m1 = 0.0f, m2 = 0.0f, i0 = 0.0f, i1 = 0.0f, i2 = 0.0f, p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    uint32_t bad = 0;
    uint4 f = __ldg(pf), ty = __ldg(pt);
    float xl = x0, xh = xl + e;
    for (int i = 0; i < 16; ++i) {
        const float xl2 = xl * xl, xh2 = xh * xh;
        const float xl3 = xl2 * xl, xh3 = xh2 * xh;
        const float h2x = xh2 - xl2, h3x = xh3 - xl3;
        float yl = y0, yh = yl + e;
        for (int j = 0; j < 16; ++j) {
            // the next column's 32 bytes are requested before this column's arithmetic
            const int r = i * 16 + j;
            uint4 fn = f, tn = ty;
            if (r + 1 < 256) {
                fn = __ldg(pf + r + 1);
                tn = __ldg(pt + r + 1);
            }
            const float yl2 = yl * yl, yh2 = yh * yh;
            const float yl3 = yl2 * yl, yh3 = yh2 * yh;
            const float h2y = yh2 - yl2, h3y = yh3 - yl3;
            const float h3xy = h3x + h3y, h2xy = h2x * h2y;
            const uint32_t fw[4] = {f.x, f.y, f.z, f.w}, tw[4] = {ty.x, ty.y, ty.z, ty.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t fb = (fw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                if (!(fb & 1u)) {
                    const uint32_t tb = (tw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                    if (tb >= a.n_densities) bad = 1;
                    const float d = s_dens[tb];
                    mass += d;
                    m0 += d * h2x;
                    m1 += d * h2y;
                    m2 += d * h2z[k];
                    i0 += d * (h3y + h3z[k]);
                    i1 += d * (h3x + h3z[k]);
                    i2 += d * h3xy;
                    p0 += d * h2xy;
                    p1 += d * (h2y * h2z[k]);
                    p2 += d * (h2z[k] * h2x);
                }
            }
            f = fn;
            ty = tn;
            yl = yh;
            yh += e;
        }
        xl = xh;
        xh += e;
    }
    if (bad) atomicOr(&a.counters[1], 1u);
    const float e2 = e * e, e3 = e2 * e;
    const float fm = 0.5f * e2, fi = (1.0f / 3.0f) * e2, fp = 0.25f * e;
    float* p = a.part + (size_t)a.row[c - a.c_begin] * 10;
    p[0] = mass * e3;
    p[1] = m0 * fm;
    p[2] = m1 * fm;
    p[3] = m2 * fm;
    p[4] = i0 * fi;
    p[5] = i1 * fi;
    p[6] = i2 * fi;
    p[7] = p0 * fp;
    p[8] = p1 * fp;
    p[9] = p2 * fp;
}

// compute_inertial_property_moments_for_object's outer loop (inertia.rs:769-787): `*mass += chunk_mass` … in linear
// chunk order. One block: warps 1.. stage the next tile of rows in shared memory while lanes 0-9 of warp 0 each extend
// one component's chain over the current tile.
constexpr int SUM_TILE = 512;  // rows per tile: 20 KiB, two tiles in flight
__global__ void __launch_bounds__(256) k_moments_sum(const float* __restrict__ part, const uint32_t* __restrict__ n_rows_ptr,
                                                     const float* initial, float* __restrict__ out) {
    __shared__ __align__(16) float s_tile[2][SUM_TILE * 10];
    const int tid = threadIdx.x;
    const uint32_t n_rows = *n_rows_ptr;
    const uint32_t n_vec = (n_rows * 10u + 3u) / 4u;  // the rows as float4 (the buffer is padded to a multiple of 16 bytes)
    const uint32_t n_tiles = (n_rows + SUM_TILE - 1) / SUM_TILE;
    const float4* part4 = reinterpret_cast<const float4*>(part);
    auto stage = [&](uint32_t tile, int first_thread, int n_threads) {
        // all of a thread's loads are in flight before the first store: a tile costs one memory latency, not six
        constexpr uint32_t TILE_VEC = SUM_TILE * 10u / 4u;
        const uint32_t v0 = tile * TILE_VEC;
        float4* dst = reinterpret_cast<float4*>(s_tile[tile & 1]);
        float4 buf[6];  // 6 x 224 threads >= 1280 vectors
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const uint32_t v = (uint32_t)(tid - first_thread) + (uint32_t)(u * n_threads);
            buf[u] = (v < TILE_VEC && v0 + v < n_vec) ? part4[v0 + v] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const uint32_t v = (uint32_t)(tid - first_thread) + (uint32_t)(u * n_threads);
            if (v < TILE_VEC) dst[v] = buf[u];
        }
    };
    if (n_tiles) stage(0, 0, 256);
    __syncthreads();
    float acc = (tid < 10) ? initial[tid] : 0.0f;
    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (tid >= 32) {
            if (t + 1 < n_tiles) stage(t + 1, 32, 224);
        } else if (tid < 10) {
            const float* s = s_tile[t & 1] + tid;
            const int rows = (int)min((uint32_t)SUM_TILE, n_rows - t * SUM_TILE);
            if (rows == SUM_TILE) {
                // the chain is one dependent FADD per row; the next 32 rows are fetched from shared memory while the
                // current 32 are being added, so the chain never waits for a load
                float cur[32], nxt[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) cur[u] = s[u * 10];
#pragma unroll
                for (int q0 = 0; q0 < SUM_TILE; q0 += 32) {
                    if (q0 + 32 < SUM_TILE) {
#pragma unroll
                        for (int u = 0; u < 32; ++u) nxt[u] = s[(q0 + 32 + u) * 10];
                    }
#pragma unroll
                    for (int u = 0; u < 32; ++u) acc += cur[u];
#pragma unroll
                    for (int u = 0; u < 32; ++u) cur[u] = nxt[u];
                }
            } else {
                for (int q = 0; q < rows; ++q) acc += s[q * 10];
            }
        }
        __syncthreads();
    }
    if (tid < 10) out[tid] = acc;
}

// the rows spread back over all chunks (void chunks: zero) for callers that ask for the per-chunk terms
__global__ void k_moments_scatter(const DevChunk* __restrict__ chunks, uint32_t c_begin, uint32_t n, const uint32_t* __restrict__ row,
                                  const float* __restrict__ part, float* __restrict__ dense) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 10u) return;
    const uint32_t c = t / 10u, q = t % 10u;
    dense[t] = chunks[c_begin + c].kind != 0 ? part[(size_t)row[c] * 10 + q] : 0.0f;
}

// ---- incremental update: the voxels an absorption emptied ------------------------------------------------------
// compute_moments_for_voxel (inertia.rs:591-625), negated: `parent.mass -= voxel_mass` is `parent.mass + (-voxel_mass)`
__device__ __forceinline__ void negated_voxel_moments(float e, float e2, float e3, float density, const uint32_t ijk[3],
                                                      float* __restrict__ out) {
    float h2[3], h3[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float lo = e * (float)ijk[d];
        const float hi = lo + e;
        const float lo2 = lo * lo, hi2 = hi * hi;
        const float lo3 = lo2 * lo, hi3 = hi2 * hi;
        h2[d] = hi2 - lo2;
        h3[d] = hi3 - lo3;
    }
    const float fm = (0.5f * e2) * density;
    const float fi = ((1.0f / 3.0f) * e2) * density;
    const float fp = (0.25f * e) * density;
    out[0] = -(e3 * density);
    out[1] = -(fm * h2[0]);
    out[2] = -(fm * h2[1]);
    out[3] = -(fm * h2[2]);
    out[4] = -(fi * (h3[1] + h3[2]));
    out[5] = -(fi * (h3[0] + h3[2]));
    out[6] = -(fi * (h3[0] + h3[1]));
    out[7] = -(fp * (h2[0] * h2[1]));
    out[8] = -(fp * (h2[1] * h2[2]));
    out[9] = -(fp * (h2[2] * h2[0]));
}

__global__ void k_removed_counts(const uint32_t* __restrict__ info, uint32_t n_range, uint32_t* __restrict__ count) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_range) count[t] = info[2 * t];
}

// One block per chunk of the touched range: the emptied voxels' negated terms, one row each, in the order the
// reference's closure was called (chunks of the range i → j → k = the range index, voxels i → j → k inside a chunk).
__global__ void __launch_bounds__(256) k_removed_terms(AbsorbRange r, const uint32_t* __restrict__ info,
                                                       const uint16_t* __restrict__ cols, const uint32_t* __restrict__ first_row,
                                                       const unsigned char* __restrict__ voxels, float e, uint32_t n_densities,
                                                       const __grid_constant__ Densities dens, float* __restrict__ rows,
                                                       uint32_t* __restrict__ error) {
    __shared__ uint32_t s_warp[8];
    const uint32_t t = blockIdx.x;
    if (info[2 * t] == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bits = cols[(size_t)t * 256 + tid];
    const uint32_t mine = __popc(bits);
    // exclusive scan of the column counts over the block (column order = voxel order)
    uint32_t x = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t before = x - mine;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (!mine) return;
    const uint32_t ek = r.c1[2] - r.c0[2], ej = r.c1[1] - r.c0[1];
    const uint32_t ck = r.c0[2] + t % ek, cj = r.c0[1] + (t / ek) % ej, ci = r.c0[0] + t / (ek * ej);
    const unsigned char* types = voxels + (size_t)info[2 * t + 1] * SLOT_BYTES + PLANE_TYPE + tid * 16;
    const float e2 = e * e, e3 = e2 * e;
    float* out = rows + ((size_t)first_row[t] + before) * 10;
    for (uint32_t b = bits; b; b &= b - 1) {
        const uint32_t k = (uint32_t)__ffs(b) - 1u;
        const uint32_t ty = types[k];
        if (ty >= n_densities) atomicOr(error, 1u);
        const uint32_t ijk[3] = {ci * 16u + (uint32_t)(tid >> 4), cj * 16u + (uint32_t)(tid & 15), ck * 16u + k};
        negated_voxel_moments(e, e2, e3, dens.v[ty], ijk, out);
        out += 10;
    }
}

}  // namespace ivx

int ivx_apply_removed_voxels(ivx_ctx* ctx, const ivx_object* obj, const AbsorbRange& r, uint32_t n_range,
                             const uint32_t* removed_info, const uint16_t* removed_cols, const InertialUpdate& upd) {
    if (n_range == 0) return IVX_OK;
    Densities dens{};
    for (uint32_t q = 0; q < upd.n_densities; ++q) dens.v[q] = upd.densities[q];
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* count = tmp.get<uint32_t>(n_range);
    uint32_t* first_row = tmp.get<uint32_t>(n_range);
    float* d_io = tmp.get<float>(32);  // [0..10) the sums before, [16..26) after, [26] rows, [27] error
    uint32_t* counters = ctx->d_scratch + 48;  // [0] rows [1] error
    if (!count || !first_row || !d_io) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "inertial update: out of device memory");
    CU(ctx, cudaMemsetAsync(counters, 0, 8, st));
    CU(ctx, cudaMemcpyAsync(d_io, upd.inout, sizeof(ivx_inertial_moments), cudaMemcpyHostToDevice, st));
    KL(ctx, (k_removed_counts<<<(n_range + 255) / 256, 256, 0, st>>>(removed_info, n_range, count), cudaGetLastError()));
    KL(ctx, launch_exclusive_scan(count, first_row, n_range, counters, st));
    uint32_t n_rows = 0;
    if (int rc = ivx_read_words(ctx, counters, 1, &n_rows)) return rc;
    if (n_rows == 0) return IVX_OK;
    float* rows = tmp.get<float>((size_t)n_rows * 10 + 4);
    if (!rows) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "inertial update: out of device memory");
    KL(ctx, (k_removed_terms<<<n_range, 256, 0, st>>>(r, removed_info, removed_cols, first_row, obj->d_voxels, obj->voxel_extent,
                                                      upd.n_densities, dens, rows, counters + 1),
             cudaGetLastError()));
    KLP(ctx, 10, (k_moments_sum<<<1, 256, 0, st>>>(rows, counters, d_io, d_io + 16), cudaGetLastError()));
    CU(ctx, cudaMemcpyAsync(d_io + 26, counters, 8, cudaMemcpyDeviceToDevice, st));
    uint32_t w[12];
    if (int rc = ivx_read_words(ctx, reinterpret_cast<const uint32_t*>(d_io + 16), 12, w)) return rc;
    if (w[11])
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT,
                 "inertial update: an emptied voxel has a type without a density (%u densities given)", upd.n_densities);
    std::memcpy(upd.inout, w, sizeof(ivx_inertial_moments));
    return IVX_OK;
}

extern "C" {

int ivx_object_inertial_moments(ivx_ctx* ctx, const ivx_object* obj, const float* voxel_type_densities,
                                uint32_t n_densities, const ivx_inertial_moments* initial, ivx_inertial_moments* out,
                                float* per_chunk_terms, size_t per_chunk_capacity) {
    if (!ctx || !obj || !out || (!voxel_type_densities && n_densities)) return IVX_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(ivx_inertial_moments) == 40, "ten packed floats");
    cudaSetDevice(ctx->device);
    if (n_densities > 256) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "inertial moments: at most 256 voxel types");
    if (obj->derive_pending)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT,
                 "inertial moments: the slab's derived state is pending (uniform chunks may still convert)");
    const uint32_t plane = obj->nb[1] * obj->nb[2];
    const uint32_t c_begin = (obj->own_begin - obj->first_i) * plane, c_end = (obj->own_end - obj->first_i) * plane;
    const uint32_t n = obj->n_chunks ? c_end - c_begin : 0;
    if (per_chunk_terms && per_chunk_capacity < (size_t)n)
        IVX_FAIL(ctx, IVX_ERR_CAPACITY, "inertial moments: per_chunk_terms holds %zu chunks, %u needed", per_chunk_capacity, n);
    ivx_inertial_moments start{};
    if (initial) start = *initial;
    if (n == 0) {
        *out = start;
        return IVX_OK;
    }
    Densities dens{};
    for (uint32_t q = 0; q < n_densities; ++q) dens.v[q] = voxel_type_densities[q];
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    MomentsArgs a{};
    a.chunks = obj->d_chunks;
    a.voxels = obj->d_voxels;
    a.nb = make_uint3(obj->nb[0], obj->nb[1], obj->nb[2]);
    a.first_i = obj->first_i;
    a.c_begin = c_begin;
    a.c_end = c_end;
    a.e = obj->voxel_extent;
    a.n_densities = n_densities;
    a.part = tmp.get<float>((size_t)n * 10 + 4);
    a.list = tmp.get<uint32_t>(n);
    uint32_t* flag = tmp.get<uint32_t>(n);
    uint32_t* row = tmp.get<uint32_t>(n);
    a.row = row;
    float* d_io = tmp.get<float>(32);  // [0..10) initial, [16..26) result
    a.counters = ctx->d_scratch + 48;
    if (!a.part || !a.list || !flag || !row || !d_io) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "inertial moments: out of device memory");
    CU(ctx, cudaMemsetAsync(a.counters, 0, 12, st));
    CU(ctx, cudaMemcpyAsync(d_io, &start, sizeof(start), cudaMemcpyHostToDevice, st));
    auto prepare = [&]() -> cudaError_t {
        k_moments_flags<<<(n + 255) / 256, 256, 0, st>>>(a.chunks, c_begin, n, flag);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = launch_exclusive_scan(flag, row, n, a.counters + 2, st);
        if (e != cudaSuccess) return e;
        k_moments_classify<<<(n + 255) / 256, 256, 0, st>>>(a, dens);
        return cudaGetLastError();
    };
    ctx->launches += 2;
    KLP(ctx, 8, prepare());
    // sized for the case that every chunk is non-uniform; threads beyond the list length leave at once
    KLP(ctx, 9, (k_moments_non_uniform<<<(n + 63) / 64, 64, 0, st>>>(a, dens), cudaGetLastError()));
    KLP(ctx, 10, (k_moments_sum<<<1, 256, 0, st>>>(a.part, a.counters + 2, d_io, d_io + 16), cudaGetLastError()));
    if (per_chunk_terms) {
        float* dense = tmp.get<float>((size_t)n * 10);
        if (!dense) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "inertial moments: out of device memory");
        KL(ctx, (k_moments_scatter<<<(n * 10 + 255) / 256, 256, 0, st>>>(a.chunks, c_begin, n, row, a.part, dense), cudaGetLastError()));
        CU(ctx, cudaMemcpyAsync(per_chunk_terms, dense, (size_t)n * 40, cudaMemcpyDeviceToHost, st));
    }
    uint32_t w[12];
    CU(ctx, cudaMemcpyAsync(d_io + 26, a.counters, 8, cudaMemcpyDeviceToDevice, st));
    if (int rc = ivx_read_words(ctx, reinterpret_cast<const uint32_t*>(d_io + 16), 12, w)) return rc;
    if (w[11])
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT,
                 "inertial moments: a non-empty voxel has a type without a density (%u densities given)", n_densities);
    std::memcpy(out, w, sizeof(*out));
    return IVX_OK;
}

}  // extern "C"
