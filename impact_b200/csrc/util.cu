// Small bookkeeping kernels: prefix sums, list compaction, capacity planning.
#include "common.cuh"
#include "kernels.h"

namespace ivx {

// Exclusive scan of n uint32 by one 1024-thread CTA (n is chunk-count sized): 8 consecutive elements per thread
// and step (two 16-byte loads / stores), so a 2 x 10^5 element array takes 25 steps.
constexpr int SCAN_PER_THREAD = 8;
__global__ void __launch_bounds__(1024) k_exclusive_scan(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                         uint32_t n, uint32_t* __restrict__ total) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    for (uint32_t base = 0; base < n; base += 1024u * SCAN_PER_THREAD) {
        const uint32_t i0 = base + (uint32_t)tid * SCAN_PER_THREAD;
        uint32_t v[SCAN_PER_THREAD];
        if (aligned && i0 + SCAN_PER_THREAD <= n) {
            const uint4 a = *reinterpret_cast<const uint4*>(in + i0), b = *reinterpret_cast<const uint4*>(in + i0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int q = 0; q < SCAN_PER_THREAD; ++q) v[q] = i0 + q < n ? in[i0 + q] : 0u;
        }
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < SCAN_PER_THREAD; ++q) {
            const uint32_t t = v[q];
            v[q] = sum;  // exclusive within the thread
            sum += t;
        }
        uint32_t x = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t off = carry + (warp > 0 ? s_warp[warp - 1] : 0u) + x - sum;
        if (aligned && i0 + SCAN_PER_THREAD <= n) {
            *reinterpret_cast<uint4*>(out + i0) = make_uint4(off + v[0], off + v[1], off + v[2], off + v[3]);
            *reinterpret_cast<uint4*>(out + i0 + 4) = make_uint4(off + v[4], off + v[5], off + v[6], off + v[7]);
        } else {
#pragma unroll
            for (int q = 0; q < SCAN_PER_THREAD; ++q)
                if (i0 + q < n) out[i0 + q] = off + v[q];
        }
        __syncthreads();
        if (tid == 1023) s_carry = off + sum;
        __syncthreads();
    }
    if (tid == 0 && total) *total = s_carry;
}

// The same scan by one thread-block CLUSTER of 8 CTAs (sm_90+ / sm_100a): 65 536 elements per step instead of 8 192.
// Every CTA scans its 8 192 elements, stores its sum into the shared memory of ALL CTAs of the cluster (distributed
// shared memory), one cluster barrier, and every CTA adds the sums of the lower ranks and the carry of the previous
// steps. The sums are double buffered by step parity, so one barrier per step is enough. A 2 x 10^5 element array (the
// chunk table of a 1024^3 object) takes 4 steps: 54 us -> ~10 us.
constexpr int SCAN_CLUSTER = 8;
__global__ void __cluster_dims__(SCAN_CLUSTER, 1, 1) __launch_bounds__(1024)
    k_exclusive_scan_cluster(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ total) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_sums[2][SCAN_CLUSTER];  // [step parity][rank]: written by the CTA of that rank, in every CTA
    __shared__ uint32_t s_mine;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    uint32_t carry = 0;
    uint32_t step = 0;
    for (uint32_t base = 0; base < n; base += SCAN_CLUSTER * 1024u * SCAN_PER_THREAD, ++step) {
        const uint32_t i0 = base + rank * 1024u * SCAN_PER_THREAD + (uint32_t)tid * SCAN_PER_THREAD;
        uint32_t v[SCAN_PER_THREAD];
        if (aligned && i0 + SCAN_PER_THREAD <= n) {
            const uint4 a = *reinterpret_cast<const uint4*>(in + i0), b = *reinterpret_cast<const uint4*>(in + i0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int q = 0; q < SCAN_PER_THREAD; ++q) v[q] = i0 + q < n ? in[i0 + q] : 0u;
        }
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < SCAN_PER_THREAD; ++q) {
            const uint32_t t = v[q];
            v[q] = sum;
            sum += t;
        }
        uint32_t x = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            s_warp[lane] = w;
            if (lane == 31) s_mine = w;  // this CTA's sum
        }
        __syncthreads();
        const uint32_t in_block = (warp > 0 ? s_warp[warp - 1] : 0u) + x - sum;
        if (tid < SCAN_CLUSTER) {
            // s_sums[parity][rank] of CTA `tid`: shared::cluster address of the peer's copy
            uint32_t local = (uint32_t)__cvta_generic_to_shared(&s_sums[step & 1u][rank]), remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((uint32_t)tid));
            asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(s_mine) : "memory");
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        uint32_t before = 0, all = 0;
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)SCAN_CLUSTER; ++r) {
            const uint32_t sr = s_sums[step & 1u][r];
            if (r < rank) before += sr;
            all += sr;
        }
        const uint32_t off = carry + before + in_block;
        if (aligned && i0 + SCAN_PER_THREAD <= n) {
            *reinterpret_cast<uint4*>(out + i0) = make_uint4(off + v[0], off + v[1], off + v[2], off + v[3]);
            *reinterpret_cast<uint4*>(out + i0 + 4) = make_uint4(off + v[4], off + v[5], off + v[6], off + v[7]);
        } else {
#pragma unroll
            for (int q = 0; q < SCAN_PER_THREAD; ++q)
                if (i0 + q < n) out[i0 + q] = off + v[q];
        }
        carry += all;
        __syncthreads();  // s_warp / s_mine are rewritten by the next step
    }
    if (rank == 0 && tid == 0 && total) *total = carry;
    // no CTA may exit while a peer can still store into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

cudaError_t launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* total, cudaStream_t st) {
    if (n > 2u * 1024u * SCAN_PER_THREAD)
        k_exclusive_scan_cluster<<<SCAN_CLUSTER, 1024, 0, st>>>(in, out, n, total);
    else
        k_exclusive_scan<<<1, 1024, 0, st>>>(in, out, n, total);
    return cudaGetLastError();
}

// capacity of a child block's instruction list = length of its parent's list
__global__ void k_child_caps(const uint32_t* __restrict__ parent_len, uint32_t n_blocks, uint3 nb, uint3 parent_nb,
                             uint32_t ratio, uint32_t* __restrict__ caps) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    uint32_t bk = b % nb.z, bj = (b / nb.z) % nb.y, bi = b / (nb.z * nb.y);
    uint32_t p = ((bi / ratio) * parent_nb.y + bj / ratio) * parent_nb.z + bk / ratio;
    caps[b] = parent_len[p];
}
cudaError_t launch_child_caps(const uint32_t* parent_len, uint32_t n_blocks, const uint32_t nb[3],
                              const uint32_t parent_nb[3], uint32_t ratio, uint32_t* caps, cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    k_child_caps<<<(n_blocks + 255) / 256, 256, 0, st>>>(parent_len, n_blocks, make_uint3(nb[0], nb[1], nb[2]),
                                                         make_uint3(parent_nb[0], parent_nb[1], parent_nb[2]), ratio, caps);
    return cudaGetLastError();
}

// After the exact fold: which chunks need voxel evaluation (ACTIVE) and which
// pre-classified uniform chunks touch a non-uniform neighbour and may have to
// be converted to non-uniform storage later (object.rs:2118-2160).
__global__ void k_plan_slots(const DevChunk* __restrict__ chunks, uint32_t n, uint3 nb, uint32_t* __restrict__ active_flag,
                             uint32_t* __restrict__ slot_flag) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const uint8_t pre = chunks[c].pre;
    uint32_t act = pre == PRE_ACTIVE ? 1u : 0u;
    uint32_t need = act;
    if (pre == PRE_UNIFORM) {
        int k = c % nb.z, j = (c / nb.z) % nb.y, i = c / (nb.z * nb.y);
        const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
        for (int f = 0; f < 6; ++f) {
            int ni = i + d[f][0], nj = j + d[f][1], nk = k + d[f][2];
            bool uni = false;
            if (ni >= 0 && nj >= 0 && nk >= 0 && ni < (int)nb.x && nj < (int)nb.y && nk < (int)nb.z)
                uni = chunks[(ni * nb.y + nj) * nb.z + nk].pre == PRE_UNIFORM;
            if (!uni) need = 1u;
        }
    }
    active_flag[c] = act;
    slot_flag[c] = need;
}
__global__ void k_scatter_active(const uint32_t* __restrict__ active_flag, const uint32_t* __restrict__ active_scan,
                                 uint32_t n, uint32_t* __restrict__ active_list) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    if (active_flag[c]) active_list[active_scan[c]] = c;
}
cudaError_t launch_plan_slots(const DevChunk* chunks, uint32_t n, const uint32_t nb[3], uint32_t* active_flag,
                              uint32_t* slot_flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_plan_slots<<<(n + 255) / 256, 256, 0, st>>>(chunks, n, make_uint3(nb[0], nb[1], nb[2]), active_flag, slot_flag);
    return cudaGetLastError();
}
cudaError_t launch_scatter_active(const uint32_t* active_flag, const uint32_t* active_scan, uint32_t n,
                                  uint32_t* active_list, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_scatter_active<<<(n + 255) / 256, 256, 0, st>>>(active_flag, active_scan, n, active_list);
    return cudaGetLastError();
}

__global__ void k_fill_u32(uint32_t* p, uint32_t n, uint32_t v) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
cudaError_t launch_fill_u32(uint32_t* p, uint32_t n, uint32_t v, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_fill_u32<<<(n + 255) / 256, 256, 0, st>>>(p, n, v);
    return cudaGetLastError();
}

}  // namespace ivx
