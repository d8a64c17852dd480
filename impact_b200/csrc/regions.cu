// Global pass of connected-region ("split") detection on the device: which local region ends up as the ROOT of each
// connected set, exactly as the reference's sequential disjoint-set pass leaves it
// (resolve_connected_regions_between_all_chunks, split_detection.rs:323-488; set_root_for_region :1914-1947).
//
// The reference visits the boundary regions in linear (chunk, region) order; the visited region v takes its current root
// R and puts the root of every adjacent region under R. Which region is the final root is a property of that visiting
// order, so it cannot come from a generic parallel labelling. It does not need the sequential pass over all ~10^5
// regions and ~2 x 10^5 connections either:
//
//   * An adjacent region u < v was visited before v and merged v's set into its own then, so only connections to HIGHER
//     regions do anything when v is visited.
//   * Let a(v) be the lowest region adjacent to v if that is lower than v, else v ("fresh": nobody reaches v before its
//     own visit). A non-fresh v is an untouched singleton until a(v) is visited, joins a(v)'s set at that moment without
//     changing its root, and never leaves it. Following a() to its fixed point m(v) (a fresh region) therefore gives a
//     forest of TREES whose members are in one set from the moment they join, and only fresh regions are ever roots.
//   * What is left are the connections (u, w), u < w, between different trees: "at time u the set of m(u) absorbs the
//     set of m(w)". Of all connections between the same two trees only the earliest can be a real merge. On a 1024^3
//     asteroid that is ~10^3 events between a few hundred trees instead of 1.8 x 10^5 connections between 6 x 10^4
//     regions; they are replayed in time order by one thread on a disjoint-set forest in shared memory.
//
// Everything before and after the replay is data parallel: a() by atomicMin over the connections, m() by chasing a()
// (chains are as long as the object is wide in chunks), the earliest connection per tree pair by a hash table with
// atomicMin, time order by counting per visiting region + prefix sum, and finally root(v) = root of m(v)'s tree.
// tests/test_gpu_split_detection.py, test_gpu_extraction.py and test_gpu_fuzz.py hold the roots against the oracle's
// sequential pass region for region.
#include "api_internal.cuh"

namespace ivx {

__device__ __forceinline__ uint32_t chunk_region_capacity(uint32_t creg) {
    // the reference gives each boundary region 256 / boundary_region_count connection slots (uniform chunks: 256,
    // split_detection.rs:1519-1546, 1769-1774)
    const uint32_t kind = creg >> 16, bc = (creg >> 8) & 255u;
    return kind == 1u ? 256u : 256u / (bc ? bc : 1u);
}

__global__ void k_region_counts(const uint32_t* __restrict__ regions, uint32_t n, uint32_t* __restrict__ counts) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) counts[c] = regions[c] & 255u;
}

// per chunk: labels of its regions, every region its own lowest neighbour so far
__global__ void k_region_vertices(RegionPass p) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = p.words[RW_TOTAL];
    if (c == 0 && total > p.cap) atomicCAS(&p.words[RW_ERROR], 0u, RERR_REGION_CAPACITY);
    if (c >= p.n) return;
    const uint32_t count = p.regions[c] & 255u, f = p.first[c];
    for (uint32_t r = 0; r < count; ++r) {
        const uint32_t e = f + r;
        if (e >= p.cap) return;
        p.label[e] = (c << 8) | r;
        p.lowest[e] = e;
    }
}

// per connection record: both region indices, lowest lower neighbour of the upper one, connection counts
__global__ void k_region_edges(RegionPass p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_rec = min(p.words[RW_RECORDS], p.record_cap);
    if (i >= n_rec || p.words[RW_TOTAL] > p.cap) return;
    const uint2 rec = p.records[i];
    const uint32_t c = rec.x, d = rec.y >> 16, la = (rec.y >> 8) & 255u, lb = rec.y & 255u;
    const uint32_t cu = c + (d == 0 ? p.stride0 : (d == 1 ? p.stride1 : 1u));
    const uint32_t u = p.first[c] + la, w = p.first[cu] + lb;
    p.edges[i] = make_uint2(u, w);
    atomicMin(&p.lowest[w], u);
    atomicAdd(&p.degree[u], 1u);
    atomicAdd(&p.degree[w], 1u);
}

// per region: its tree m(v); fresh regions with connections are the trees that take part in the replay
__global__ void k_region_trees(RegionPass p) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = p.words[RW_TOTAL];
    if (e >= total || total > p.cap) return;
    uint32_t x = e, up;
    while ((up = p.lowest[x]) != x) x = up;
    p.tree[e] = x;
    const uint32_t deg = p.degree[e];
    p.fresh_flag[e] = (x == e && deg > 0u) ? 1u : 0u;
    const uint32_t lab = p.label[e], room = chunk_region_capacity(p.regions[lab >> 8]);
    if (deg > room && atomicCAS(&p.words[RW_ERROR], 0u, RERR_TOO_MANY_CONNECTIONS) == 0u) {
        p.words[RW_ERROR_INFO + 0] = lab >> 8;
        p.words[RW_ERROR_INFO + 1] = lab & 255u;
        p.words[RW_ERROR_INFO + 2] = deg;
        p.words[RW_ERROR_INFO + 3] = room;
    }
}

__device__ __forceinline__ uint32_t hash_pair(uint32_t lo, uint32_t hi) {
    uint32_t h = lo * 0x9E3779B1u ^ (hi + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0xC2B2AE3Du;
    return h ^ (h >> 13);
}

// per connection between different trees: the earliest one of its tree pair (value = time << 1 | "the absorbing tree is
// the higher of the two"); connections of one visiting region all have the same absorbing tree, so ties agree
__global__ void k_region_pairs(RegionPass p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_rec = min(p.words[RW_RECORDS], p.record_cap);
    if (i >= n_rec || p.words[RW_TOTAL] > p.cap) return;
    const uint2 ed = p.edges[i];
    const uint32_t mu = p.tree[ed.x], mw = p.tree[ed.y];
    if (mu == mw) return;
    const uint32_t lo = min(mu, mw), hi = max(mu, mw);
    const unsigned long long key = ((unsigned long long)lo << 32) | hi;
    const uint32_t value = (ed.x << 1) | (mu == lo ? 0u : 1u);
    uint32_t h = hash_pair(lo, hi) & p.slot_mask;
    for (uint32_t probe = 0; probe <= p.slot_mask; ++probe) {
        const unsigned long long seen = atomicCAS(&p.slot_keys[h], ~0ull, key);
        if (seen == ~0ull || seen == key) {
            atomicMin(&p.slot_values[h], value);
            return;
        }
        h = (h + 1u) & p.slot_mask;
    }
    atomicCAS(&p.words[RW_ERROR], 0u, RERR_PAIR_TABLE_FULL);
}

__global__ void k_region_event_counts(RegionPass p) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > p.slot_mask || p.slot_keys[s] == ~0ull) return;
    atomicAdd(&p.event_count[p.slot_values[s] >> 1], 1u);
}

// events in time order: (absorbing tree, absorbed tree) as tree numbers
__global__ void k_region_events(RegionPass p) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > p.slot_mask) return;
    const unsigned long long key = p.slot_keys[s];
    if (key == ~0ull) return;
    const uint32_t value = p.slot_values[s], u = value >> 1;
    const uint32_t lo = (uint32_t)(key >> 32), hi = (uint32_t)key;
    const uint32_t a = (value & 1u) ? hi : lo, b = (value & 1u) ? lo : hi;
    const uint32_t pos = p.event_offset[u] + atomicAdd(&p.event_cursor[u], 1u);
    if (pos < p.record_cap) p.events[pos] = make_uint2(p.tree_number[a], p.tree_number[b]);
}

__global__ void k_region_tree_list(RegionPass p) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = p.words[RW_TOTAL];
    if (e >= total || total > p.cap) return;
    if (p.fresh_flag[e]) p.tree_vertex[p.tree_number[e]] = e;
}

// The replay: one thread, disjoint-set forest over the trees (shared memory while they fit), the events staged through
// shared memory by the whole block. "The absorbing set keeps its root" is the reference's set_root_for_region.
constexpr uint32_t REPLAY_THREADS = 1024;
constexpr uint32_t REPLAY_STAGE = 2048;
__global__ void __launch_bounds__(REPLAY_THREADS) k_region_replay(RegionPass p, uint32_t shared_trees) {
    extern __shared__ uint32_t s_dyn[];
    __shared__ uint2 s_events[REPLAY_STAGE];
    const uint32_t n_trees = p.words[RW_TREES], n_events = min(p.words[RW_EVENTS], p.record_cap);
    if (p.words[RW_TOTAL] > p.cap || p.words[RW_ERROR] != 0u) return;
    uint32_t* parent = n_trees <= shared_trees ? s_dyn : p.tree_parent;
    for (uint32_t t = threadIdx.x; t < n_trees; t += REPLAY_THREADS) parent[t] = t;
    __syncthreads();
    for (uint32_t base = 0; base < n_events; base += REPLAY_STAGE) {
        const uint32_t chunk = min(REPLAY_STAGE, n_events - base);
        for (uint32_t t = threadIdx.x; t < chunk; t += REPLAY_THREADS) s_events[t] = p.events[base + t];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (uint32_t t = 0; t < chunk; ++t) {
                uint32_t a = s_events[t].x, b = s_events[t].y, up;
                while ((up = parent[a]) != a) {  // path halving
                    const uint32_t upup = parent[up];
                    parent[a] = upup;
                    a = upup;
                }
                while ((up = parent[b]) != b) {
                    const uint32_t upup = parent[up];
                    parent[b] = upup;
                    b = upup;
                }
                if (a != b) parent[b] = a;
            }
        }
        __syncthreads();
    }
    for (uint32_t t = threadIdx.x; t < n_trees; t += REPLAY_THREADS) {
        uint32_t x = t, up;
        while ((up = parent[x]) != x) x = up;
        p.tree_root[t] = p.tree_vertex[x];
    }
}

// per region: its root; the roots are counted and the lowest one is kept (find_two_disconnected_regions takes the
// first two in linear order, split_detection.rs:193-250)
__global__ void k_region_roots(RegionPass p) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = p.words[RW_TOTAL];
    if (e >= total || total > p.cap || p.words[RW_ERROR] != 0u) return;
    const uint32_t m = p.tree[e];
    const uint32_t root = p.fresh_flag[m] ? p.tree_root[p.tree_number[m]] : m;
    p.root[e] = root;
    if (root == e) {
        atomicAdd(&p.words[RW_ROOTS], 1u);
        atomicMin(&p.words[RW_FIRST_ROOT], e);
    }
}

__global__ void k_region_second_root(RegionPass p) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = p.words[RW_TOTAL];
    if (e >= total || total > p.cap || p.words[RW_ERROR] != 0u) return;
    if (p.root[e] == e && e != p.words[RW_FIRST_ROOT]) atomicMin(&p.words[RW_SECOND_ROOT], e);
}

// per chunk: what extract_smallest_region compares (extraction.rs:137-245): chunks, NonUniform chunks and chunk bounds
// of the two regions
__global__ void k_region_candidates(RegionPass p) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && p.words[RW_ROOTS] >= 2u) {
        p.words[RW_FIRST_LABEL] = p.label[p.words[RW_FIRST_ROOT]];
        p.words[RW_SECOND_LABEL] = p.label[p.words[RW_SECOND_ROOT]];
    }
    if (c >= p.n || p.words[RW_ROOTS] < 2u || p.words[RW_TOTAL] > p.cap || p.words[RW_ERROR] != 0u) return;
    const uint32_t creg = p.regions[c], count = creg & 255u, f = p.first[c];
    if (count == 0u) return;
    const uint32_t two[2] = {p.words[RW_FIRST_ROOT], p.words[RW_SECOND_ROOT]};
    bool found[2] = {false, false};
    for (uint32_t r = 0; r < count; ++r) {
        const uint32_t root = p.root[f + r];
        found[0] = found[0] || root == two[0];
        found[1] = found[1] || root == two[1];
    }
    const uint32_t idx[3] = {c / (p.stride0), (c / p.stride1) % (p.stride0 / p.stride1), c % p.stride1};
    for (int q = 0; q < 2; ++q) {
        if (!found[q]) continue;
        uint32_t* st = p.words + RW_CANDIDATES + q * 8;
        atomicAdd(&st[0], 1u);
        if ((creg >> 16) == 2u) atomicAdd(&st[1], 1u);
        for (int d = 0; d < 3; ++d) {
            atomicMin(&st[2 + d], idx[d]);
            atomicMax(&st[5 + d], idx[d]);
        }
    }
}

__global__ void k_region_result_init(uint32_t* __restrict__ words) {
    const uint32_t t = threadIdx.x;
    if (t >= RW_COUNT) return;
    uint32_t v = 0u;
    if (t == RW_FIRST_ROOT || t == RW_SECOND_ROOT) v = 0xFFFFFFFFu;
    if (t >= RW_CANDIDATES && ((t - RW_CANDIDATES) % 8u) >= 2u && ((t - RW_CANDIDATES) % 8u) < 5u) v = 0xFFFFFFFFu;  // chunk min
    words[t] = v;
}

__global__ void k_region_root_labels(const uint32_t* __restrict__ root, const uint32_t* __restrict__ label, uint32_t total,
                                     uint32_t* __restrict__ out) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) out[e] = label[root[e]];
}

// ---- extraction: which chunks of the region's bounding box move, and how (extraction.rs:137-245, 339-349) ----
__global__ void k_extract_classify(ExtractPlanArgs a) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n_ext) return;
    const uint32_t k = e % a.ext[2], j = (e / a.ext[2]) % a.ext[1], i = e / (a.ext[2] * a.ext[1]);
    const uint32_t c = ((a.lo[0] + i) * a.nb1 + (a.lo[1] + j)) * a.nb2 + (a.lo[2] + k);
    const uint32_t creg = a.regions[c], count = creg & 255u, kind = creg >> 16, f = a.first[c];
    bool in_region = false, mixed = false;
    for (uint32_t r = 0; r < count; ++r) {
        if (a.root[f + r] == a.region_root) in_region = true;
        else mixed = true;
    }
    uint8_t mode = 0;
    uint32_t src = 0xFFFFFFFFu, non_uniform = 0u;
    if (in_region && kind != 0u) {
        src = c;
        if (kind == 1u) {
            mode = 1;
            atomicAdd(a.uniform_count, 1u);
        } else {
            mode = mixed ? 3 : 2;
            non_uniform = 1u;
        }
    }
    a.mode[e] = mode;
    a.src_index[e] = src;
    a.first_region[e] = (in_region && kind != 0u) ? f : 0u;
    a.non_uniform_flag[e] = non_uniform;
}

__global__ void k_extract_slots(const uint32_t* __restrict__ flag, uint32_t n, uint32_t* __restrict__ slot) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n && !flag[e]) slot[e] = 0xFFFFFFFFu;
}

__global__ void k_region_membership(const uint32_t* __restrict__ root, uint32_t total, uint32_t region_root,
                                    uint8_t* __restrict__ is_member) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) is_member[e] = root[e] == region_root ? 1 : 0;
}

// ---------------------------------------------------------------------------
static inline uint32_t blocks(uint32_t n) { return (n + 255u) / 256u; }

cudaError_t launch_region_global_pass(const RegionPass& p, uint32_t n_trees_scan, uint32_t* launches, int max_shared_bytes,
                                      cudaStream_t st) {
    // sizes known to the host are capacities; the kernels read the actual counts from p.words
    const uint32_t slots = p.slot_mask + 1u;
    k_region_counts<<<blocks(p.n), 256, 0, st>>>(p.regions, p.n, p.counts);
    if (cudaError_t e = launch_exclusive_scan(p.counts, p.first, p.n, p.words + RW_TOTAL, st)) return e;
    k_region_vertices<<<blocks(p.n), 256, 0, st>>>(p);
    k_region_edges<<<blocks(p.record_cap), 256, 0, st>>>(p);
    k_region_trees<<<blocks(p.cap), 256, 0, st>>>(p);
    if (cudaError_t e = launch_exclusive_scan(p.fresh_flag, p.tree_number, n_trees_scan, p.words + RW_TREES, st)) return e;
    k_region_tree_list<<<blocks(p.cap), 256, 0, st>>>(p);
    k_region_pairs<<<blocks(p.record_cap), 256, 0, st>>>(p);
    k_region_event_counts<<<blocks(slots), 256, 0, st>>>(p);
    if (cudaError_t e = launch_exclusive_scan(p.event_count, p.event_offset, n_trees_scan, p.words + RW_EVENTS, st)) return e;
    k_region_events<<<blocks(slots), 256, 0, st>>>(p);
    uint32_t shared_trees = (uint32_t)std::max(0, max_shared_bytes - (int)(REPLAY_STAGE * sizeof(uint2)) - 1024) / 4u;
    if (std::getenv("IVX_REGIONS_GLOBAL_FOREST")) shared_trees = 0;  // tests: the path of objects with more trees than fit
    if (cudaError_t e = cudaFuncSetAttribute(k_region_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(shared_trees * 4u)))
        return e;
    k_region_replay<<<1, REPLAY_THREADS, shared_trees * 4u, st>>>(p, shared_trees);
    k_region_roots<<<blocks(p.cap), 256, 0, st>>>(p);
    k_region_second_root<<<blocks(p.cap), 256, 0, st>>>(p);
    k_region_candidates<<<blocks(p.n), 256, 0, st>>>(p);
    *launches += 16;
    return cudaGetLastError();
}

cudaError_t launch_region_result_init(uint32_t* words, cudaStream_t st) {
    k_region_result_init<<<1, 64, 0, st>>>(words);
    return cudaGetLastError();
}

cudaError_t launch_region_root_labels(const uint32_t* root, const uint32_t* label, uint32_t total, uint32_t* out, cudaStream_t st) {
    if (total == 0) return cudaSuccess;
    k_region_root_labels<<<blocks(total), 256, 0, st>>>(root, label, total, out);
    return cudaGetLastError();
}

cudaError_t launch_extract_plan(const ExtractPlanArgs& a, uint32_t* slot_total, cudaStream_t st) {
    k_extract_classify<<<blocks(a.n_ext), 256, 0, st>>>(a);
    if (cudaError_t e = launch_exclusive_scan(a.non_uniform_flag, a.dst_slot, a.n_ext, slot_total, st)) return e;
    k_extract_slots<<<blocks(a.n_ext), 256, 0, st>>>(a.non_uniform_flag, a.n_ext, a.dst_slot);
    if (a.total) k_region_membership<<<blocks(a.total), 256, 0, st>>>(a.root, a.total, a.region_root, a.is_member);
    return cudaGetLastError();
}

}  // namespace ivx
