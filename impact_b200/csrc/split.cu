// Connected-region ("split") detection (sm_100a + host):
//
//   k_init_regions        per chunk: kind and, for uniform chunks, their single region
//   k_local_regions       per NonUniform chunk: the reference's chunk-local connected-component labelling
//                         (split_detection.rs:662-893) with ITS label numbering
//   k_region_connections  per chunk and upper face: the distinct (region, adjacent region) pairs across the
//                         face (what the connection updaters of split_detection.rs:1046-1326, 1424-1463 record)
//   resolve (host)        the chunk-level union-find of resolve_connected_regions_between_all_chunks
//                         (split_detection.rs:323-488), count_regions / find_two_disconnected_regions
//                         (:193-301) and the smallest-region choice of extraction.rs:121-281
//
// Why the labelling is emulated sequentially. Which local region ends up representing a global region — and so
// which two regions find_two_disconnected_regions reports and which one an extraction splits off — depends on
// the local label numbering, which in turn depends on which voxel the reference's union-find leaves as the
// root of each local region ("the current voxel's root absorbs the roots of its upper neighbours", visited in
// linear voxel order). That is a property of the visiting order, not of the component structure, so a parallel
// generic parallel labelling cannot reproduce it. One warp per chunk therefore replays the reference's sequence
// on shared-memory state (8 KiB parents + 4 KiB flags + 4 KiB labels, ~13 chunks in flight per SM), parallel
// only where the order provably does not matter: all merges of one k-run go under the same root, the
// boundary traversal is taken 32 positions at a time with the earliest lane winning a contested root, and
// label numbers are ballot prefix counts. The global pass works on ~10^5 regions and stays on the host, as
// in the reference.
#include <chrono>

#include "api_internal.cuh"

namespace ivx {

constexpr uint32_t LABEL_EMPTY = 255u;

// Labels are a pure function of a chunk's own voxel flags, so only chunks modified since the last resolve
// (stale != 0) are re-labelled; the others keep their labels and counts.
__global__ void k_init_regions(const DevChunk* __restrict__ chunks, uint32_t n, uint32_t* __restrict__ regions,
                               uint8_t* __restrict__ stale, uint32_t* __restrict__ work_flag) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const uint32_t kind = chunks[c].kind;
    const bool relabel = kind == 2u && stale[c] != 0;
    if (kind != 2u) regions[c] = kind == 1u ? ((1u << 16) | (1u << 8) | 1u) : 0u;
    work_flag[c] = relabel ? 1u : 0u;
    stale[c] = 0;
}

__global__ void k_mark_box(uint8_t* __restrict__ arr, uint3 nb, uint3 lo, uint3 hi, uint8_t value) {
    const uint32_t ext_y = hi.y - lo.y, ext_z = hi.z - lo.z;
    const uint32_t total = (hi.x - lo.x) * ext_y * ext_z;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const uint32_t k = lo.z + t % ext_z, j = lo.y + (t / ext_z) % ext_y, i = lo.x + t / (ext_z * ext_y);
    arr[(i * nb.y + j) * nb.z + k] = value;
}

cudaError_t launch_mark_box(uint8_t* arr, const uint32_t nb[3], const uint32_t lo[3], const uint32_t hi[3], uint8_t value,
                            cudaStream_t st) {
    if (!arr) return cudaSuccess;
    uint32_t h[3], l[3];
    for (int d = 0; d < 3; ++d) {
        l[d] = std::min(lo[d], nb[d]);
        h[d] = std::min(hi[d], nb[d]);
        if (l[d] >= h[d]) return cudaSuccess;
    }
    const uint32_t total = (h[0] - l[0]) * (h[1] - l[1]) * (h[2] - l[2]);
    k_mark_box<<<(total + 255) / 256, 256, 0, st>>>(arr, make_uint3(nb[0], nb[1], nb[2]), make_uint3(l[0], l[1], l[2]),
                                                    make_uint3(h[0], h[1], h[2]), value);
    return cudaGetLastError();
}

// `par` is accessed through a volatile pointer: during the union pass the lanes of a run walk and compress paths of the
// same forest concurrently. Every racing store writes an ancestor of the slot's set (path compression) or the run's
// root, so any interleaving leaves a valid forest with the same roots; volatile 16-bit accesses are single relaxed
// memory operations (ld/st.volatile), which makes those races defined behaviour under the PTX memory model instead of
// data races on plain stores.
__device__ __forceinline__ uint32_t find_root_compress(volatile uint16_t* par, uint32_t idx) {
    uint32_t r = idx;
    while (par[r] != r) r = par[r];
    while (par[idx] != r) {
        const uint32_t nx = par[idx];
        par[idx] = (uint16_t)r;
        idx = nx;
    }
    return r;
}

__global__ void __launch_bounds__(32) k_local_regions(const DevChunk* __restrict__ chunks, const uint32_t* __restrict__ work,
                                                      uint32_t n_work, const unsigned char* __restrict__ voxels,
                                                      uint8_t* __restrict__ labels, uint32_t* __restrict__ regions,
                                                      uint32_t* __restrict__ error_flag) {
    __shared__ __align__(16) uint8_t s_flags[4096];
    __shared__ __align__(16) uint8_t s_lab[4096];
    __shared__ volatile uint16_t s_par[4096];
    __shared__ uint32_t s_counts[2];
    const int lane = threadIdx.x;
    for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const uint32_t chunk = work[w];
        const uint32_t slot = chunks[chunk].slot;
        const unsigned char* fp = voxels + (size_t)slot * SLOT_BYTES + PLANE_FLAGS;
        uint32_t nonempty = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const uint4 v = *reinterpret_cast<const uint4*>(fp + (t * 32 + lane) * 16);
            *reinterpret_cast<uint4*>(&s_flags[(t * 32 + lane) * 16]) = v;
            // IS_EMPTY is bit 0 of every flag byte
            nonempty += 16u - (__popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
                               __popc(v.w & 0x01010101u));
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) nonempty += __shfl_xor_sync(0xffffffffu, nonempty, d);
        uint8_t* out = labels + (size_t)slot * 4096;
        if (nonempty == 4096u || nonempty == 0u) {
            // one region filling the chunk (label 0, it touches the boundary) / no region at all
            const uint32_t fill = nonempty ? 0u : 0xFFFFFFFFu;
#pragma unroll
            for (int t = 0; t < 8; ++t) *reinterpret_cast<uint4*>(out + (t * 32 + lane) * 16) = make_uint4(fill, fill, fill, fill);
            if (lane == 0) regions[chunk] = (2u << 16) | (nonempty ? ((1u << 8) | 1u) : 0u);
            __syncwarp();
            continue;
        }
        for (int idx = lane; idx < 4096; idx += 32) s_par[idx] = (uint16_t)idx;
        __syncwarp();
        // ---- union pass (split_detection.rs:701-744) ----
        // The reference visits the voxels in linear order; the visited voxel's root absorbs the roots of its
        // upper x / y / z neighbours. Along a k-run of linked voxels every voxel therefore ends up under the
        // root the run's FIRST voxel had when it was visited, and so do all upper neighbours of the run — in
        // whatever order those merges happen. One run per step: lane 0 finds the run's root, then the lanes of
        // the run's voxels merge their upper neighbours' sets under it concurrently (racing writes all store
        // the same root; path compression only ever stores an ancestor).
        for (uint32_t row = 0; row < 256u; ++row) {
            const uint32_t f = lane < 16 ? (uint32_t)s_flags[row * 16u + lane] : 1u;
            const uint32_t present = __ballot_sync(0xffffffffu, (f & 1u) == 0u) & 0xFFFFu;
            if (present == 0u) continue;
            // voxel k is linked to k + 1 when it is present and carries HAS_ADJACENT_Z_UP
            const uint32_t link = __ballot_sync(0xffffffffu, (f & 1u) == 0u && (f & 0x80u) != 0u && lane < 15) & 0x7FFFu;
            const uint32_t i = row >> 4, j = row & 15u;
            uint32_t todo = present;
            while (todo) {
                const uint32_t k0 = (uint32_t)__ffs(todo) - 1u;
                // the run [k0, k1]: extend while linked
                uint32_t k1 = k0;
                while (k1 < 15u && ((link >> k1) & 1u) && ((present >> (k1 + 1u)) & 1u)) ++k1;
                const uint32_t run = ((2u << k1) - 1u) & ~((1u << k0) - 1u);
                todo &= ~run;
                uint32_t root = 0;
                if (lane == 0) root = find_root_compress(s_par, row * 16u + k0);
                root = __shfl_sync(0xffffffffu, root, 0);
                __syncwarp();
                if ((run >> lane) & 1u) {
                    const uint32_t idx = row * 16u + (uint32_t)lane;
                    if (i < 15u && (f & (1u << 5))) {
                        const uint32_t r = find_root_compress(s_par, idx + 256u);
                        if (r != root) s_par[r] = (uint16_t)root;
                    }
                    if (j < 15u && (f & (1u << 6))) {
                        const uint32_t r = find_root_compress(s_par, idx + 16u);
                        if (r != root) s_par[r] = (uint16_t)root;
                    }
                    if ((link >> lane) & 1u) {
                        const uint32_t r = find_root_compress(s_par, idx + 1u);
                        if (r != root) s_par[r] = (uint16_t)root;
                    }
                }
                __syncwarp();
            }
        }
        // ---- representative voxels of boundary regions, in the reference's face order (:760-796) ----
        // 32 boundary voxels of the traversal per step. A region is labelled where its root voxel is visited
        // if the root lies on the boundary, else at its first visited boundary voxel, which then becomes the
        // root (make_voxel_root); labels count these events in traversal order.
        uint32_t current = 0;
        for (uint32_t base = 0; base < 1352u; base += 32u) {
            const uint32_t pos = base + (uint32_t)lane;
            uint32_t idx = 0xFFFFu;
            if (pos < 512u) {
                idx = ((pos < 256u ? 0u : 15u) << 8) | (pos & 255u);
            } else if (pos < 960u) {
                const uint32_t p = pos < 736u ? pos - 512u : pos - 736u;
                idx = ((1u + (p >> 4)) << 8) | ((pos < 736u ? 0u : 15u) << 4) | (p & 15u);
            } else if (pos < 1352u) {
                const uint32_t p = pos < 1156u ? pos - 960u : pos - 1156u;
                idx = ((1u + p / 14u) << 8) | ((1u + p % 14u) << 4) | (pos < 1156u ? 0u : 15u);
            }
            const bool valid = idx != 0xFFFFu;
            const bool present = valid && (s_flags[idx] & 1u) == 0u;
            uint32_t set_id = 0xFFFFFFFFu;
            if (present) set_id = find_root_compress(s_par, idx);
            __syncwarp();
            bool interior_root = false;
            if (present && set_id != idx) {
                const uint32_t si = set_id >> 8, sj = (set_id >> 4) & 15u, sk = set_id & 15u;
                interior_root = si > 0u && si < 15u && sj > 0u && sj < 15u && sk > 0u && sk < 15u;
            }
            // among the lanes that hit the same interior root, the earliest position takes it over
            const uint32_t peers = __match_any_sync(0xffffffffu, interior_root ? set_id : (0x10000u | (uint32_t)lane));
            bool event = present && set_id == idx;
            if (interior_root && (uint32_t)(__ffs(peers) - 1) == (uint32_t)lane) {
                s_par[set_id] = (uint16_t)idx;  // make_voxel_root
                s_par[idx] = (uint16_t)idx;
                event = true;
            }
            const uint32_t events = __ballot_sync(0xffffffffu, event);
            if (event) s_lab[idx] = (uint8_t)min(current + __popc(events & ((1u << lane) - 1u)), 255u);
            else if (valid && !present) s_lab[idx] = (uint8_t)LABEL_EMPTY;
            current += __popc(events);
            __syncwarp();
        }
        const uint32_t boundary = current;
        // ---- interior-only regions: interior root voxels in linear order (:803-822) ----
        for (uint32_t base = 0; base < 4096u; base += 32u) {
            const uint32_t idx = base + (uint32_t)lane;
            const uint32_t vi = idx >> 8, vj = (idx >> 4) & 15u, vk = idx & 15u;
            const bool interior = vi > 0u && vi < 15u && vj > 0u && vj < 15u && vk > 0u && vk < 15u;
            const bool is_root = interior && s_par[idx] == idx;
            const bool event = is_root && (s_flags[idx] & 1u) == 0u;
            const uint32_t events = __ballot_sync(0xffffffffu, event);
            if (event) s_lab[idx] = (uint8_t)min(current + __popc(events & ((1u << lane) - 1u)), 255u);
            else if (is_root) s_lab[idx] = (uint8_t)LABEL_EMPTY;
            current += __popc(events);
        }
        if (lane == 0) {
            s_counts[0] = min(boundary, 255u);
            s_counts[1] = min(current, 255u);
            // the reference asserts boundary < 255 and total < 255 (:798, :835)
            if (boundary >= 255u || current >= 255u) atomicExch(error_flag, 1u);
        }
        __syncwarp();
        // ---- every other non-empty voxel takes its root's label (:837-855) ----
        for (int idx = lane; idx < 4096; idx += 32) {
            if (s_flags[idx] & 1u) continue;
            uint32_t r = idx;
            while (s_par[r] != r) r = s_par[r];
            if (r != (uint32_t)idx) s_lab[idx] = s_lab[r];
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t)
            *reinterpret_cast<uint4*>(out + (t * 32 + lane) * 16) = *reinterpret_cast<const uint4*>(&s_lab[(t * 32 + lane) * 16]);
        if (lane == 0) regions[chunk] = (2u << 16) | (s_counts[0] << 8) | s_counts[1];
        __syncwarp();
    }
}

// One warp per chunk: for each of its three upper faces with a non-void neighbour, the distinct pairs
// (label on this side, label on the far side) over the 256 face-adjacent voxel pairs where both voxels are
// present. Record: x = lower chunk's linear index, y = dim << 16 | label << 8 | adjacent label.
constexpr int CONN_WARPS = 4;
constexpr int CONN_SEEN = 96;

__global__ void __launch_bounds__(CONN_WARPS * 32) k_region_connections(const DevChunk* __restrict__ chunks, uint32_t n, uint3 nb,
                                                                        const uint8_t* __restrict__ labels,
                                                                        uint2* __restrict__ records, uint32_t capacity,
                                                                        uint32_t* __restrict__ counter,
                                                                        uint32_t* __restrict__ error_flag) {
    __shared__ uint32_t s_seen[CONN_WARPS][CONN_SEEN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t c = blockIdx.x * CONN_WARPS + warp;
    if (c >= n) return;
    const DevChunk lo = chunks[c];
    if (lo.kind == 0u) return;
    const uint32_t ck = c % nb.z, cj = (c / nb.z) % nb.y, ci = c / (nb.z * nb.y);
    uint32_t* seen = s_seen[warp];
    for (int d = 0; d < 3; ++d) {
        uint32_t cu;
        if (d == 0) { if (ci + 1 >= nb.x) continue; cu = c + nb.y * nb.z; }
        else if (d == 1) { if (cj + 1 >= nb.y) continue; cu = c + nb.z; }
        else { if (ck + 1 >= nb.z) continue; cu = c + 1; }
        const DevChunk up = chunks[cu];
        if (up.kind == 0u) continue;
        uint32_t count = 0;
        if (lo.kind == 1u && up.kind == 1u) {
            if (lane == 0) seen[0] = 0u;
            count = 1;
        } else {
            const uint8_t* ll = lo.kind == 2u ? labels + (size_t)lo.slot * 4096 : nullptr;
            const uint8_t* lu = up.kind == 2u ? labels + (size_t)up.slot * 4096 : nullptr;
            const uint32_t stride_a = d == 0 ? 16u : 256u, stride_b = d == 2 ? 16u : 1u;  // the two in-face axes
            const uint32_t face_lo = d == 0 ? 15u * 256u : (d == 1 ? 15u * 16u : 15u);
            for (int round = 0; round < 8; ++round) {
                const uint32_t p = (uint32_t)round * 32u + (uint32_t)lane;  // 0..255
                const uint32_t off = (p >> 4) * stride_a + (p & 15u) * stride_b;
                const uint32_t la = ll ? ll[face_lo + off] : 0u;
                const uint32_t lb = lu ? lu[off] : 0u;
                const uint32_t key = (la == LABEL_EMPTY || lb == LABEL_EMPTY) ? 0xFFFFFFFFu : ((la << 8) | lb);
                const uint32_t peers = __match_any_sync(0xffffffffu, key);
                const bool leader = key != 0xFFFFFFFFu && (__ffs(peers) - 1) == lane;
                uint32_t todo = __ballot_sync(0xffffffffu, leader);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const uint32_t k2 = __shfl_sync(0xffffffffu, key, src);
                    bool found = false;
                    for (uint32_t t = lane; t < count; t += 32) found = found || (seen[t] == k2);
                    if (!__any_sync(0xffffffffu, found)) {
                        if (count < (uint32_t)CONN_SEEN) {
                            if (lane == 0) seen[count] = k2;
                            count++;
                        } else if (lane == 0) {
                            atomicExch(error_flag, 2u);  // more distinct pairs on one face than any sane chunk has
                        }
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
        if (count) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(counter, count);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t t = lane; t < count; t += 32)
                if (base + t < capacity) records[base + t] = make_uint2(c, ((uint32_t)d << 16) | seen[t]);
        }
        __syncwarp();
    }
}

// labels by slot → 4096 per NonUniform chunk in linear chunk order
__global__ void __launch_bounds__(256) k_pack_labels(const DevChunk* __restrict__ chunks, uint32_t n,
                                                     const uint32_t* __restrict__ ordinal, const uint8_t* __restrict__ labels,
                                                     uint8_t* __restrict__ out) {
    for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
        if (chunks[c].kind != 2u) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(labels + (size_t)chunks[c].slot * 4096 + threadIdx.x * 16);
        *reinterpret_cast<uint4*>(out + (size_t)ordinal[c] * 4096 + threadIdx.x * 16) = v;
    }
}

}  // namespace ivx

// ---------------------------------------------------------------------------
namespace {

struct RegionForest {
    std::vector<uint32_t>& parent;  // per region entry: GlobalRegionLabel of its parent
    const std::vector<uint32_t>& first;
    uint32_t entry(uint32_t label) const { return first[label >> 8] + (label & 255u); }
    uint32_t find(uint32_t label) {
        uint32_t r = label;
        while (parent[entry(r)] != r) r = parent[entry(r)];
        while (parent[entry(label)] != r) {
            const uint32_t nx = parent[entry(label)];
            parent[entry(label)] = r;
            label = nx;
        }
        return r;
    }
};

}  // namespace

extern "C" {

int ivx_object_resolve_connected_regions(ivx_ctx* ctx, ivx_object* obj, ivx_split_info* out) {
    if (!ctx || !obj || !out) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    std::memset(out, 0, sizeof(*out));
    obj->split_valid = false;
    if (obj->derive_pending || obj->halo_present[0] || obj->halo_present[1] || obj->nb[0] != obj->chunk_counts[0])
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "connected regions are resolved on whole objects (gather the slabs on one rank first)");
    const uint32_t n = obj->n_chunks;
    if (n == 0) {
        obj->h_chunk_regions.clear();
        obj->h_first_region.clear();
        obj->h_region_roots.clear();
        obj->split_valid = true;
        return IVX_OK;
    }
    if (n > (1u << 24)) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "GlobalRegionLabel holds 24 bits of chunk index");
    cudaStream_t st = ctx->stream;
    Tmp tmp(ctx);
    if (obj->label_slots < obj->slot_capacity || !obj->d_labels) {
        ctx->release(obj->d_labels);
        obj->d_labels = static_cast<uint8_t*>(ctx->alloc(std::max<size_t>(1, (size_t)obj->slot_capacity) * 4096));
        if (!obj->d_labels) {
            obj->label_slots = 0;
            IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "region labels (%u chunks): out of device memory", obj->slot_capacity);
        }
        obj->label_slots = obj->slot_capacity;
        if (obj->d_label_stale) CU(ctx, cudaMemsetAsync(obj->d_label_stale, 1, n, st));  // the new buffer holds no labels
    }
    if (!obj->d_regions) {
        obj->d_regions = static_cast<uint32_t*>(ctx->alloc((size_t)n * 4));
        obj->d_label_stale = static_cast<uint8_t*>(ctx->alloc(n));
        if (!obj->d_regions || !obj->d_label_stale) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "split detection: out of device memory");
        CU(ctx, cudaMemsetAsync(obj->d_label_stale, 1, n, st));  // nothing labelled yet
    }
    uint32_t* regions = obj->d_regions;
    uint32_t* flag = tmp.get<uint32_t>(n);
    uint32_t* scan = tmp.get<uint32_t>(n);
    uint32_t* work = tmp.get<uint32_t>(n);
    if (!regions || !flag || !scan || !work) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "split detection: out of device memory");
    uint32_t* counters = ctx->d_scratch + 32;  // [0] n_work [1] n_records [2] error
    CU(ctx, cudaMemsetAsync(counters, 0, 16, st));

    struct EventPair {  // destroyed on every exit path
        cudaEvent_t a = nullptr, b = nullptr;
        EventPair() {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
        }
        ~EventPair() {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } ev;
    const cudaEvent_t e0 = ev.a, e1 = ev.b;
    cudaEventRecord(e0, st);
    ctx->launches++;
    k_init_regions<<<(n + 255) / 256, 256, 0, st>>>(obj->d_chunks, n, regions, obj->d_label_stale, flag);
    CU(ctx, cudaGetLastError());
    KL(ctx, launch_exclusive_scan(flag, scan, n, counters, st));
    KL(ctx, launch_scatter_active(flag, scan, n, work, st));
    uint32_t words[4];
    if (int rc = ivx_read_words(ctx, counters, 3, words)) return rc;
    const uint32_t n_work = words[0];
    if (n_work) {
        ctx->launches++;
        const uint32_t grid = std::min<uint32_t>(n_work, (uint32_t)ctx->sm_count * 13u);
        k_local_regions<<<grid, 32, 0, st>>>(obj->d_chunks, work, n_work, obj->d_voxels, obj->d_labels, regions, counters + 2);
        CU(ctx, cudaGetLastError());
    }
    // connection records: a few per face; grown and re-run if the first guess is too small
    uint32_t capacity = std::max<uint32_t>(4096u, 4u * (n - 0u));
    uint2* records = nullptr;
    uint32_t n_records = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        records = tmp.get<uint2>(capacity);
        if (!records) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "split detection: out of device memory");
        CU(ctx, cudaMemsetAsync(counters + 1, 0, 4, st));
        ctx->launches++;
        k_region_connections<<<(n + CONN_WARPS - 1) / CONN_WARPS, CONN_WARPS * 32, 0, st>>>(
            obj->d_chunks, n, make_uint3(obj->nb[0], obj->nb[1], obj->nb[2]), obj->d_labels, records, capacity, counters + 1,
            counters + 2);
        CU(ctx, cudaGetLastError());
        if (attempt == 0) cudaEventRecord(e1, st);
        if (int rc = ivx_read_words(ctx, counters, 3, words)) return rc;
        n_records = words[1];
        if (n_records <= capacity) break;
        capacity = n_records;
    }
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&out->device_ms, e0, e1);
    if (words[2] == 1u)
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "a chunk has more than 254 connected regions (the reference asserts, split_detection.rs:798)");
    if (words[2] == 2u)
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "a chunk face has more than %d distinct region connections", CONN_SEEN);

    std::vector<uint32_t>& creg = obj->h_chunk_regions;
    creg.resize(n);
    // host scratch keeps its capacity across calls (a resolve per modification step: no page faults on MBs of vectors)
    static thread_local std::vector<uint2> recs;
    recs.resize(n_records);
    CU(ctx, cudaMemcpyAsync(creg.data(), regions, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (n_records) CU(ctx, cudaMemcpyAsync(recs.data(), records, (size_t)n_records * sizeof(uint2), cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));

    // ---- host: chunk-level union-find in the reference's visiting order ----
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<uint32_t>& first = obj->h_first_region;
    first.resize(n + 1);
    uint32_t total = 0;
    for (uint32_t c = 0; c < n; ++c) {
        first[c] = total;
        total += creg[c] & 255u;
    }
    first[n] = total;
    static thread_local std::vector<uint32_t> parent, deg, start, adj, fill;
    parent.resize(total);
    for (uint32_t c = 0; c < n; ++c)
        for (uint32_t r = 0; r < (creg[c] & 255u); ++r) parent[first[c] + r] = (c << 8) | r;
    // adjacency in CSR form, both directions
    const uint32_t stride[3] = {obj->nb[1] * obj->nb[2], obj->nb[2], 1u};
    deg.assign(total + 1, 0);
    for (const uint2& r : recs) {
        const uint32_t c = r.x, d = r.y >> 16, la = (r.y >> 8) & 255u, lb = r.y & 255u, cu = c + stride[d];
        deg[first[c] + la]++;
        deg[first[cu] + lb]++;
    }
    start.assign(total + 1, 0);
    for (uint32_t i = 0; i < total; ++i) start[i + 1] = start[i] + deg[i];
    adj.resize(start[total]);
    fill.assign(start.begin(), start.end() - 1);
    for (const uint2& r : recs) {
        const uint32_t c = r.x, d = r.y >> 16, la = (r.y >> 8) & 255u, lb = r.y & 255u, cu = c + stride[d];
        adj[fill[first[c] + la]++] = (cu << 8) | lb;
        adj[fill[first[cu] + lb]++] = (c << 8) | la;
    }
    // the reference gives each boundary region 256 / boundary_region_count connection slots (uniform chunks: 256)
    for (uint32_t c = 0; c < n; ++c) {
        const uint32_t kind = creg[c] >> 16, bc = (creg[c] >> 8) & 255u;
        const uint32_t cap = kind == 1u ? 256u : 256u / std::max(1u, bc);
        for (uint32_t r = 0; r < (creg[c] & 255u); ++r)
            if (deg[first[c] + r] > cap)
                IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "chunk %u region %u has %u adjacent regions, the reference keeps %u (split_detection.rs:1519-1546)",
                         c, r, deg[first[c] + r], cap);
    }
    RegionForest F{parent, first};
    uint32_t occ[3][2];
    for (int d = 0; d < 3; ++d) {
        occ[d][0] = obj->occ_voxels[d] / 16;
        occ[d][1] = (obj->occ_voxels[3 + d] + 15) / 16;
    }
    const auto lin = [&](uint32_t i, uint32_t j, uint32_t k) { return (i * obj->nb[1] + j) * obj->nb[2] + k; };
    for (uint32_t i = occ[0][0]; i < occ[0][1]; ++i)
        for (uint32_t j = occ[1][0]; j < occ[1][1]; ++j)
            for (uint32_t k = occ[2][0]; k < occ[2][1]; ++k) {
                const uint32_t c = lin(i, j, k);
                const uint32_t bc = (creg[c] >> 8) & 255u;
                for (uint32_t r = 0; r < bc; ++r) {
                    const uint32_t root = F.find((c << 8) | r);
                    const uint32_t e = first[c] + r;
                    for (uint32_t q = start[e]; q < start[e + 1]; ++q) {
                        const uint32_t oroot = F.find(adj[q]);
                        if (oroot != root) parent[F.entry(oroot)] = root;
                    }
                }
            }
    std::vector<uint32_t>& roots = obj->h_region_roots;
    roots.resize(total);
    for (uint32_t c = 0; c < n; ++c)
        for (uint32_t r = 0; r < (creg[c] & 255u); ++r) roots[first[c] + r] = F.find((c << 8) | r);

    // count_regions / find_two_disconnected_regions
    uint32_t two[2] = {0, 0}, n_regions = 0;
    for (uint32_t i = occ[0][0]; i < occ[0][1]; ++i)
        for (uint32_t j = occ[1][0]; j < occ[1][1]; ++j)
            for (uint32_t k = occ[2][0]; k < occ[2][1]; ++k) {
                const uint32_t c = lin(i, j, k);
                for (uint32_t r = 0; r < (creg[c] & 255u); ++r)
                    if (roots[first[c] + r] == ((c << 8) | r)) {
                        if (n_regions < 2) two[n_regions] = (c << 8) | r;
                        n_regions++;
                    }
            }
    out->n_regions = n_regions;
    out->has_two = n_regions >= 2 ? 1u : 0u;
    out->n_local_regions = total;
    out->n_connections = n_records;
    out->n_relabelled_chunks = n_work;
    if (out->has_two) {
        for (int q = 0; q < 2; ++q) {
            out->candidates[q].label = two[q];
            for (int d = 0; d < 3; ++d) out->candidates[q].chunk_min[d] = 0xFFFFFFFFu;
        }
        for (uint32_t i = occ[0][0]; i < occ[0][1]; ++i)
            for (uint32_t j = occ[1][0]; j < occ[1][1]; ++j)
                for (uint32_t k = occ[2][0]; k < occ[2][1]; ++k) {
                    const uint32_t c = lin(i, j, k);
                    const uint32_t idx3[3] = {i, j, k};
                    bool found[2] = {false, false};
                    for (uint32_t r = 0; r < (creg[c] & 255u); ++r) {
                        const uint32_t root = roots[first[c] + r];
                        int q;
                        if (root == two[0] && !found[0]) q = 0;
                        else if (root == two[1] && !found[1]) q = 1;
                        else continue;
                        found[q] = true;
                        ivx_region_candidate& cd = out->candidates[q];
                        cd.chunk_count++;
                        if ((creg[c] >> 16) == 2u) cd.non_uniform_chunk_count++;
                        for (int d = 0; d < 3; ++d) {
                            cd.chunk_min[d] = std::min(cd.chunk_min[d], idx3[d]);
                            cd.chunk_max[d] = std::max(cd.chunk_max[d], idx3[d]);
                        }
                    }
                }
        const ivx_region_candidate &a = out->candidates[0], &b = out->candidates[1];
        if (a.non_uniform_chunk_count != b.non_uniform_chunk_count) out->smallest = a.non_uniform_chunk_count < b.non_uniform_chunk_count ? 0u : 1u;
        else out->smallest = a.chunk_count < b.chunk_count ? 0u : 1u;
    }
    out->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    obj->split_valid = true;
    return IVX_OK;
}

int ivx_object_split_detection_download(ivx_ctx* ctx, const ivx_object* obj, uint8_t* voxel_labels, size_t label_capacity,
                                        ivx_chunk_regions* per_chunk, size_t chunk_capacity, uint32_t* region_roots,
                                        size_t region_capacity) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (!obj->split_valid) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "call ivx_object_resolve_connected_regions first");
    const uint32_t n = obj->n_chunks;
    if (per_chunk) {
        if (chunk_capacity < n) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u chunk entries", n);
        for (uint32_t c = 0; c < n; ++c) {
            per_chunk[c].region_count = (uint16_t)(obj->h_chunk_regions[c] & 255u);
            per_chunk[c].boundary_region_count = (uint16_t)((obj->h_chunk_regions[c] >> 8) & 255u);
            per_chunk[c].first_region = obj->h_first_region[c];
        }
    }
    if (region_roots) {
        if (region_capacity < obj->h_region_roots.size())
            IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %zu region roots", obj->h_region_roots.size());
        std::memcpy(region_roots, obj->h_region_roots.data(), obj->h_region_roots.size() * 4);
    }
    if (voxel_labels && n) {
        Tmp tmp(ctx);
        cudaStream_t st = ctx->stream;
        uint32_t* flag = tmp.get<uint32_t>(n);
        uint32_t* ord = tmp.get<uint32_t>(n);
        if (!flag || !ord) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "download: out of device memory");
        KL(ctx, launch_nonuniform_flags(obj->d_chunks, n, flag, st));
        KL(ctx, launch_exclusive_scan(flag, ord, n, ctx->d_scratch + 28, st));
        uint32_t nnu;
        if (int rc = ivx_read_words(ctx, ctx->d_scratch + 28, 1, &nnu)) return rc;
        if (label_capacity < (size_t)nnu * 4096) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %zu labels", (size_t)nnu * 4096);
        if (nnu) {
            uint8_t* d_out = tmp.get<uint8_t>((size_t)nnu * 4096);
            if (!d_out) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "download: out of device memory");
            ctx->launches++;
            k_pack_labels<<<ivx_persistent_grid(ctx, n, 8), 256, 0, st>>>(obj->d_chunks, n, ord, obj->d_labels, d_out);
            CU(ctx, cudaGetLastError());
            CU(ctx, cudaMemcpyAsync(voxel_labels, d_out, (size_t)nnu * 4096, cudaMemcpyDeviceToHost, st));
            CU(ctx, cudaStreamSynchronize(st));
        }
    }
    return IVX_OK;
}

}  // extern "C"
