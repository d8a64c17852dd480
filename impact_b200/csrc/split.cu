// Connected-region ("split") detection (sm_100a + host):
//
//   k_init_regions        per chunk: kind and, for uniform chunks, their single region
//   k_local_regions       per NonUniform chunk: the reference's chunk-local connected-component labelling
//                         (split_detection.rs:662-893) with ITS label numbering
//   k_region_connections  per chunk and upper face: the distinct (region, adjacent region) pairs across the
//                         face (what the connection updaters of split_detection.rs:1046-1326, 1424-1463 record)
//   regions.cu            the roots resolve_connected_regions_between_all_chunks (split_detection.rs:323-488) leaves,
//                         count_regions / find_two_disconnected_regions (:193-301) and the counts behind the
//                         smallest-region choice of extraction.rs:121-281 — on the device, one read-back per resolve
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
// label numbers are ballot prefix counts. The global pass over ~10^5 regions has the same kind of order
// dependence; regions.cu reduces it to a replay of ~10^3 events (see there).
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
                                                      const uint32_t* __restrict__ n_work_ptr, const unsigned char* __restrict__ voxels,
                                                      uint8_t* __restrict__ labels, uint32_t* __restrict__ regions,
                                                      uint32_t* __restrict__ error_flag) {
    __shared__ __align__(16) uint8_t s_flags[4096];
    __shared__ __align__(16) uint8_t s_lab[4096];
    __shared__ volatile uint16_t s_par[4096];
    __shared__ uint32_t s_seen[256];
    __shared__ uint32_t s_counts[2];
    const int lane = threadIdx.x;
    const uint32_t n_work = *n_work_ptr;  // written by the scan before this launch: no host round trip for the grid size
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
        __syncwarp();  // the staged flags are read across lanes from here on
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
        // ---- union pass (split_detection.rs:701-744) ----
        // The reference visits the voxels in linear order; the visited voxel's root absorbs the roots of its upper
        // x / y / z neighbours. Which voxel ends up as the root of a region depends on that order — and the labels below
        // depend on the roots — but the order only matters between a few sets (regions.cu has the argument for the
        // chunk-level pass; it is the same rule): a voxel v with a linked LOWER neighbour joins the set of the lowest one,
        // a(v), when that one is visited and stays with it, so chasing a() to a voxel without lower neighbours gives
        // trees that are merged wholesale; only links between different trees can move a root, and they are replayed in
        // visiting order by one lane.
        //   link u -> u + 256: u present and HAS_ADJACENT_X_UP; u -> u + 16: Y_UP; u -> u + 1: Z_UP and u + 1 present
        uint16_t* s_events = reinterpret_cast<uint16_t*>(s_lab);  // (absorbing tree, absorbed tree) pairs; labels come later
        constexpr uint32_t MAX_EVENTS = 1024;
        // a(v) for every voxel (independent of each other), then six rounds of pointer jumping: a() goes down by 256, 16
        // or 1, so a chain has at most 45 links and par[v] <- par[par[v]] reaches its end in six rounds. A round reads
        // entries other lanes may be rewriting; every value it can see is an ancestor on the same chain at least as far
        // up as the previous round guaranteed, so the in-place rounds converge like separate ones. (The accesses are
        // volatile: grouped four at a time by hand so that four loads are in flight per lane.)
        for (uint32_t base = 0; base < 4096u; base += 128u) {
            uint32_t a4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t idx = base + 32u * u + (uint32_t)lane, vi = idx >> 8, vj = (idx >> 4) & 15u, vk = idx & 15u;
                const uint32_t f_x = vi > 0u ? s_flags[idx - 256u] : 1u, f_y = vj > 0u ? s_flags[idx - 16u] : 1u;
                const uint32_t f_z = (vk > 0u && (s_flags[idx] & 1u) == 0u) ? s_flags[idx - 1u] : 1u;
                uint32_t a = idx;
                if ((f_z & 1u) == 0u && (f_z & 0x80u)) a = idx - 1u;
                if ((f_y & 1u) == 0u && (f_y & 0x40u)) a = idx - 16u;
                if ((f_x & 1u) == 0u && (f_x & 0x20u)) a = idx - 256u;  // the lowest linked neighbour wins
                a4[u] = a;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) s_par[base + 32u * u + (uint32_t)lane] = (uint16_t)a4[u];
        }
        __syncwarp();
        for (int round = 0; round < 6; ++round) {
            for (uint32_t base = 0; base < 4096u; base += 128u) {
                uint32_t p4[4], g4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) p4[u] = s_par[base + 32u * u + (uint32_t)lane];
#pragma unroll
                for (int u = 0; u < 4; ++u) g4[u] = s_par[p4[u]];
                __syncwarp();  // (in place: every value read is an ancestor either way; the barriers order the accesses)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (g4[u] != p4[u]) s_par[base + 32u * u + (uint32_t)lane] = (uint16_t)g4[u];
                __syncwarp();
            }
        }
        // links between different trees, in visiting order. Most links repeat a pair of trees that an earlier link has
        // already brought together, and a repeated pair is a no-op in the replay (a dropped FIRST occurrence would not
        // be). Two filters: (1) if v and its upper neighbour w both hang below the voxels one step down the same axis
        // (a(v) = v - s, a(w) = w - s) and v - s is linked to w - s in the same direction, that earlier link has the
        // same two trees — rods of voxels growing side by side from a surface produce one link per rod pair instead of
        // one per voxel pair; (2) a small direct-mapped memory of recent pairs (look first, note after the batch).
        const auto lowest_step = [&](uint32_t idx) -> uint32_t {  // idx - a(idx)
            const uint32_t vi = idx >> 8, vj = (idx >> 4) & 15u, vk = idx & 15u;
            if (vi > 0u) {
                const uint32_t fu = s_flags[idx - 256u];
                if ((fu & 1u) == 0u && (fu & 0x20u)) return 256u;
            }
            if (vj > 0u) {
                const uint32_t fu = s_flags[idx - 16u];
                if ((fu & 1u) == 0u && (fu & 0x40u)) return 16u;
            }
            if (vk > 0u && (s_flags[idx] & 1u) == 0u) {
                const uint32_t fu = s_flags[idx - 1u];
                if ((fu & 1u) == 0u && (fu & 0x80u)) return 1u;
            }
            return 0u;
        };
        for (int t = lane; t < 256; t += 32) s_seen[t] = 0xFFFFFFFFu;
        __syncwarp();
        uint32_t n_events = 0;
        bool overflow = false;
        // 128 voxels per step, four consecutive ones per lane (their loads overlap; the forest is only read here: plain
        // loads): the links of a lane come out in voxel order, the lanes in order behind each other
        const uint16_t* par = const_cast<const uint16_t*>(s_par);
        for (uint32_t base = 0; base < 4096u; base += 128u) {
            uint32_t mine4[4], other[12], keys[12], count = 0;
            uint8_t owner[12];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t idx = base + 4u * (uint32_t)lane + (uint32_t)u, vi = idx >> 8, vj = (idx >> 4) & 15u, vk = idx & 15u;
                const uint32_t f = s_flags[idx];
                mine4[u] = par[idx];
                if ((f & 1u) != 0u) continue;
                const uint32_t mine = mine4[u];
                const uint32_t w3[3] = {idx + 256u, idx + 16u, idx + 1u};
                const uint32_t bit3[3] = {0x20u, 0x40u, 0x80u};
                const bool linked[3] = {vi < 15u && (f & 0x20u) != 0u, vj < 15u && (f & 0x40u) != 0u,
                                        vk < 15u && (f & 0x80u) != 0u && (s_flags[(idx + 1u) & 4095u] & 1u) == 0u};
                uint32_t my_step = 0xFFFFFFFFu;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (!linked[d]) continue;
                    const uint32_t theirs = par[w3[d]];
                    if (theirs == mine) continue;
                    if (my_step == 0xFFFFFFFFu) my_step = lowest_step(idx);
                    if (my_step != 0u && my_step != (w3[d] - idx) && lowest_step(w3[d]) == my_step) {
                        // the voxels one step down: is v - s linked to w - s the same way?
                        const uint32_t fl = s_flags[idx - my_step];
                        const bool lower_link = (fl & bit3[d]) != 0u && (d != 2 || (s_flags[w3[d] - my_step] & 1u) == 0u);
                        if (lower_link) continue;
                    }
                    const uint32_t key = (min(mine, theirs) << 12) | max(mine, theirs);
                    if (s_seen[(key * 0x9E3779B1u) >> 24] == key) continue;
                    other[count] = theirs;
                    keys[count] = key;
                    owner[count] = (uint8_t)u;
                    count++;
                }
            }
            __syncwarp();
            // (several lanes may note pairs that share an entry: any of them may stay, so an exchange, not a plain store)
            for (uint32_t q = 0; q < count; ++q) atomicExch(&s_seen[(keys[q] * 0x9E3779B1u) >> 24], keys[q]);
            uint32_t offset = count;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, offset, d);
                if (lane >= d) offset += y;
            }
            const uint32_t batch = __shfl_sync(0xffffffffu, offset, 31);
            offset -= count;
            if (n_events + batch > MAX_EVENTS) {
                overflow = true;
                break;
            }
            for (uint32_t q = 0; q < count; ++q) {
                s_events[2u * (n_events + offset + q)] = (uint16_t)mine4[owner[q]];
                s_events[2u * (n_events + offset + q) + 1u] = (uint16_t)other[q];
            }
            n_events += batch;
            __syncwarp();
        }
        if (!overflow) {
            __syncwarp();
            if (lane == 0) {
                for (uint32_t e = 0; e < n_events; ++e) {
                    const uint32_t ra = find_root_compress(s_par, s_events[2u * e]);
                    const uint32_t rb = find_root_compress(s_par, s_events[2u * e + 1u]);
                    if (ra != rb) s_par[rb] = (uint16_t)ra;
                }
            }
            __syncwarp();
        } else {
            // more links between trees than the list holds (a very ragged chunk): the reference's sequence, one k-run of
            // linked voxels per step — lane 0 finds the run's root, the lanes of the run's voxels merge their upper
            // neighbours' sets under it concurrently (racing writes all store the same root; path compression only ever
            // stores an ancestor)
            __syncwarp();
            for (int idx = lane; idx < 4096; idx += 32) s_par[idx] = (uint16_t)idx;
            __syncwarp();
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
        {
            // nobody writes the forest any more: plain loads, four voxels in flight per lane
            const uint16_t* par = const_cast<const uint16_t*>(s_par);
#pragma unroll 4
            for (int idx = lane; idx < 4096; idx += 32) {
                if (s_flags[idx] & 1u) continue;
                uint32_t r = idx;
                while (par[r] != r) r = par[r];
                if (r != (uint32_t)idx) s_lab[idx] = s_lab[r];
            }
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
                                                                        const uint32_t* __restrict__ regions,
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
        // FaceVoxelDistribution of the two touching faces (0 Empty, 1 Full, 2 Mixed; Uniform chunks are full): an empty
        // face connects nothing, two full faces of chunks with one region each connect exactly those two — no labels read
        const uint32_t face_lo = lo.kind == 1u ? 1u : lo.face[2 * d + 1], face_up = up.kind == 1u ? 1u : up.face[2 * d];
        if (face_lo == 0u || face_up == 0u) continue;
        if (face_lo == 1u && face_up == 1u && (regions[c] & 255u) == 1u && (regions[cu] & 255u) == 1u) {
            if (lane == 0) seen[0] = 0u;
            count = 1;
        } else {
            const uint8_t* ll = lo.kind == 2u ? labels + (size_t)lo.slot * 4096 : nullptr;
            const uint8_t* lu = up.kind == 2u ? labels + (size_t)up.slot * 4096 : nullptr;
            const uint32_t stride_a = d == 0 ? 16u : 256u, stride_b = d == 2 ? 16u : 1u;  // the two in-face axes
            const uint32_t upper_face = d == 0 ? 15u * 256u : (d == 1 ? 15u * 16u : 15u);
            for (int round = 0; round < 8; ++round) {
                const uint32_t p = (uint32_t)round * 32u + (uint32_t)lane;  // 0..255
                const uint32_t off = (p >> 4) * stride_a + (p & 15u) * stride_b;
                const uint32_t la = ll ? ll[upper_face + off] : 0u;
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
        obj->region_total = 0;
        obj->split_valid = true;
        return IVX_OK;
    }
    if (n > (1u << 24)) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "GlobalRegionLabel holds 24 bits of chunk index");
    cudaStream_t st = ctx->stream;
    const auto t_begin = std::chrono::steady_clock::now();
    if (obj->label_slots < obj->slot_capacity || !obj->d_labels) {
        // the voxel storage grew (chunks converted to NonUniform took new slots): the labels move along; the new slots
        // belong to chunks that were marked stale when they were converted
        uint8_t* grown = static_cast<uint8_t*>(ctx->alloc(std::max<size_t>(1, (size_t)obj->slot_capacity) * 4096));
        if (!grown) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "region labels (%u chunks): out of device memory", obj->slot_capacity);
        if (obj->d_labels && obj->label_slots)
            CU(ctx, cudaMemcpyAsync(grown, obj->d_labels, (size_t)obj->label_slots * 4096, cudaMemcpyDeviceToDevice, st));
        else if (obj->d_label_stale)
            CU(ctx, cudaMemsetAsync(obj->d_label_stale, 1, n, st));  // no labels at all yet
        ctx->release(obj->d_labels);
        obj->d_labels = grown;
        obj->label_slots = obj->slot_capacity;
    }
    if (!obj->d_regions) {
        obj->d_regions = static_cast<uint32_t*>(ctx->alloc((size_t)n * 4));
        obj->d_label_stale = static_cast<uint8_t*>(ctx->alloc(n));
        if (!obj->d_regions || !obj->d_label_stale) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "split detection: out of device memory");
        CU(ctx, cudaMemsetAsync(obj->d_label_stale, 1, n, st));  // nothing labelled yet
    }
    if (!obj->d_region_first) {
        obj->d_region_first = static_cast<uint32_t*>(ctx->alloc((size_t)n * 4));
        if (!obj->d_region_first) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "split detection: out of device memory");
    }
    // Capacities from the last resolve of this object where there was one; the kernels never write past them and report
    // what they needed, so a guess that turns out too small costs one more pass, not a wrong answer.
    uint32_t region_cap = std::max(obj->region_cap, n + 1024u);
    uint32_t record_cap = obj->region_records ? obj->region_records + obj->region_records / 4u + 4096u : std::max<uint32_t>(4096u, 4u * n);
    if (std::getenv("IVX_REGIONS_TINY_CAPACITIES")) {  // tests: every capacity starts too small, the retries find the sizes
        region_cap = std::max(obj->region_cap, 8u);
        record_cap = 16u;
    }
    uint32_t* words = ctx->d_scratch + 64;  // RegionWord
    uint32_t h[RW_COUNT];

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
    cudaEventRecord(ev.a, st);
    int shared_optin = 0;
    cudaDeviceGetAttribute(&shared_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    for (int attempt = 0;; ++attempt) {
        if (attempt == 4) IVX_FAIL(ctx, IVX_ERR_CUDA, "split detection: capacities did not settle");
        Tmp tmp(ctx);
        if (obj->region_cap < region_cap || !obj->d_region_root) {
            ctx->release(obj->d_region_root);
            ctx->release(obj->d_region_label);
            obj->d_region_root = static_cast<uint32_t*>(ctx->alloc((size_t)region_cap * 4));
            obj->d_region_label = static_cast<uint32_t*>(ctx->alloc((size_t)region_cap * 4));
            obj->region_cap = (obj->d_region_root && obj->d_region_label) ? region_cap : 0;
            if (!obj->region_cap) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "split detection: out of device memory");
        }
        uint32_t slots = 1024;
        while (slots < 2u * record_cap) slots <<= 1;
        uint32_t* flag = tmp.get<uint32_t>(n);
        uint32_t* scan = tmp.get<uint32_t>(n);
        uint32_t* work = tmp.get<uint32_t>(n);
        uint32_t* per_region = tmp.get<uint32_t>((size_t)region_cap * 6);   // lowest, tree, tree_number, tree_vertex, tree_root, tree_parent
        uint32_t* zeroed = tmp.get<uint32_t>((size_t)region_cap * 4);       // degree, fresh_flag, event_count, event_cursor
        uint32_t* event_offset = tmp.get<uint32_t>(region_cap);
        uint2* records = tmp.get<uint2>(record_cap);
        uint2* edges = tmp.get<uint2>(record_cap);
        uint2* events = tmp.get<uint2>(record_cap);
        unsigned long long* slot_keys = tmp.get<unsigned long long>(slots);
        uint32_t* slot_values = tmp.get<uint32_t>(slots);
        if (!flag || !scan || !work || !per_region || !zeroed || !event_offset || !records || !edges || !events || !slot_keys || !slot_values)
            IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "split detection: out of device memory");
        KL(ctx, launch_region_result_init(words, st));
        CU(ctx, cudaMemsetAsync(zeroed, 0, (size_t)region_cap * 16, st));
        CU(ctx, cudaMemsetAsync(slot_keys, 0xFF, (size_t)slots * 8, st));
        CU(ctx, cudaMemsetAsync(slot_values, 0xFF, (size_t)slots * 4, st));

        // chunk-local labels of the chunks modified since the last resolve
        ctx->launches++;
        k_init_regions<<<(n + 255) / 256, 256, 0, st>>>(obj->d_chunks, n, obj->d_regions, obj->d_label_stale, flag);
        CU(ctx, cudaGetLastError());
        KL(ctx, launch_exclusive_scan(flag, scan, n, words + RW_WORK, st));
        KL(ctx, launch_scatter_active(flag, scan, n, work, st));
        ctx->launches++;
        k_local_regions<<<std::min<uint32_t>(n, (uint32_t)ctx->sm_count * 13u), 32, 0, st>>>(
            obj->d_chunks, work, words + RW_WORK, obj->d_voxels, obj->d_labels, obj->d_regions, words + RW_LABEL_ERROR);
        CU(ctx, cudaGetLastError());
        // connections across chunk faces
        ctx->launches++;
        k_region_connections<<<(n + CONN_WARPS - 1) / CONN_WARPS, CONN_WARPS * 32, 0, st>>>(
            obj->d_chunks, n, make_uint3(obj->nb[0], obj->nb[1], obj->nb[2]), obj->d_regions, obj->d_labels, records, record_cap,
            words + RW_RECORDS,
            words + RW_LABEL_ERROR);
        CU(ctx, cudaGetLastError());
        // roots of the connected regions
        RegionPass p{};
        p.n = n;
        p.stride0 = obj->nb[1] * obj->nb[2];
        p.stride1 = obj->nb[2];
        p.regions = obj->d_regions;
        p.words = words;
        p.cap = region_cap;
        p.counts = flag;
        p.first = obj->d_region_first;
        p.label = obj->d_region_label;
        p.root = obj->d_region_root;
        p.lowest = per_region;
        p.tree = per_region + (size_t)region_cap;
        p.tree_number = per_region + (size_t)region_cap * 2;
        p.tree_vertex = per_region + (size_t)region_cap * 3;
        p.tree_root = per_region + (size_t)region_cap * 4;
        p.tree_parent = per_region + (size_t)region_cap * 5;
        p.degree = zeroed;
        p.fresh_flag = zeroed + (size_t)region_cap;
        p.event_count = zeroed + (size_t)region_cap * 2;
        p.event_cursor = zeroed + (size_t)region_cap * 3;
        p.event_offset = event_offset;
        p.records = records;
        p.record_cap = record_cap;
        p.edges = edges;
        p.events = events;
        p.slot_keys = slot_keys;
        p.slot_values = slot_values;
        p.slot_mask = slots - 1u;
        uint32_t launched = 0;
        CU(ctx, launch_region_global_pass(p, region_cap, &launched, shared_optin, st));
        ctx->launches += launched;
        cudaEventRecord(ev.b, st);
        if (int rc = ivx_read_words(ctx, words, RW_COUNT, h)) return rc;  // the one synchronisation of a resolve
        if (h[RW_LABEL_ERROR] == 1u)
            IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "a chunk has more than 254 connected regions (the reference asserts, split_detection.rs:798)");
        if (h[RW_LABEL_ERROR] == 2u)
            IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "a chunk face has more than %d distinct region connections", CONN_SEEN);
        bool again = false;
        if (h[RW_RECORDS] > record_cap) {
            record_cap = h[RW_RECORDS] + h[RW_RECORDS] / 8u + 1024u;
            again = true;
        }
        if (h[RW_TOTAL] > region_cap) {
            region_cap = h[RW_TOTAL] + h[RW_TOTAL] / 4u + 1024u;
            again = true;
        }
        if (h[RW_ERROR] == RERR_PAIR_TABLE_FULL) {
            record_cap *= 2u;
            again = true;
        }
        if (again) continue;
        if (h[RW_ERROR] == RERR_TOO_MANY_CONNECTIONS)
            IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "chunk %u region %u has %u adjacent regions, the reference keeps %u (split_detection.rs:1519-1546)",
                     h[RW_ERROR_INFO], h[RW_ERROR_INFO + 1], h[RW_ERROR_INFO + 2], h[RW_ERROR_INFO + 3]);
        if (h[RW_ERROR]) IVX_FAIL(ctx, IVX_ERR_CUDA, "split detection: unexpected error word %u", h[RW_ERROR]);
        break;
    }
    cudaEventElapsedTime(&out->device_ms, ev.a, ev.b);
    obj->region_total = h[RW_TOTAL];
    obj->region_records = h[RW_RECORDS];
    obj->region_two[0] = h[RW_FIRST_ROOT];
    obj->region_two[1] = h[RW_SECOND_ROOT];

    // count_regions / find_two_disconnected_regions (split_detection.rs:193-301) and the choice of
    // extract_smallest_region (extraction.rs:121-281)
    out->n_regions = h[RW_ROOTS];
    out->has_two = h[RW_ROOTS] >= 2u ? 1u : 0u;
    out->n_local_regions = h[RW_TOTAL];
    out->n_connections = h[RW_RECORDS];
    out->n_relabelled_chunks = h[RW_WORK];
    if (out->has_two) {
        for (int q = 0; q < 2; ++q) {
            ivx_region_candidate& cd = out->candidates[q];
            const uint32_t* sw = h + RW_CANDIDATES + q * 8;
            cd.label = h[RW_FIRST_LABEL + q];
            cd.chunk_count = sw[0];
            cd.non_uniform_chunk_count = sw[1];
            for (int d = 0; d < 3; ++d) {
                cd.chunk_min[d] = sw[2 + d];
                cd.chunk_max[d] = sw[5 + d];
            }
        }
        const ivx_region_candidate &a = out->candidates[0], &b = out->candidates[1];
        if (a.non_uniform_chunk_count != b.non_uniform_chunk_count) out->smallest = a.non_uniform_chunk_count < b.non_uniform_chunk_count ? 0u : 1u;
        else out->smallest = a.chunk_count < b.chunk_count ? 0u : 1u;
    }
    // what the host spent around the device work (launches, the wait); there is no host-side pass over the regions
    out->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count() - out->device_ms;
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
    if (per_chunk && n) {
        if (chunk_capacity < n) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u chunk entries", n);
        std::vector<uint32_t> creg(n), first(n);
        CU(ctx, cudaMemcpyAsync(creg.data(), obj->d_regions, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaMemcpyAsync(first.data(), obj->d_region_first, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        for (uint32_t c = 0; c < n; ++c) {
            per_chunk[c].region_count = (uint16_t)(creg[c] & 255u);
            per_chunk[c].boundary_region_count = (uint16_t)((creg[c] >> 8) & 255u);
            per_chunk[c].first_region = first[c];
        }
    }
    if (region_roots && obj->region_total) {
        if (region_capacity < obj->region_total) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u region roots", obj->region_total);
        Tmp tmp(ctx);
        uint32_t* d_out = tmp.get<uint32_t>(obj->region_total);
        if (!d_out) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "download: out of device memory");
        KL(ctx, launch_region_root_labels(obj->d_region_root, obj->d_region_label, obj->region_total, d_out, ctx->stream));
        CU(ctx, cudaMemcpyAsync(region_roots, d_out, (size_t)obj->region_total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
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
