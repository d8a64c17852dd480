// Collision probes of a meshed voxel object (sm_100a): `VoxelObjectCollisionProbes` (collidable.rs:97-101, 346-780) —
// per meshed chunk and per block of 1^3 .. 8^3 voxels, the mesh vertex with the lowest (most convex) curvature; the
// points the physics uses to probe other objects. `MeshedVoxelObject::create` computes them for all chunks right after
// the mesh, `sync_mesh_with_object` for the invalidated chunks right after the mesh sync (mesh.rs:156-205).
//
//   k_probe_points   one CTA per chunk submesh. The reference walks the triangles in index order and adds a curvature
//                    sample (normal . outgoing edge - normal . incoming edge) to each of the three vertices; f32 sums
//                    depend on that order, so every thread takes ONE vertex and walks the chunk's triangles in order
//                    (the index loads are the same address for the whole warp: one broadcast per triangle). The block a
//                    vertex falls in keeps the vertex with the smallest mean curvature, the first one of equals:
//                    atomicMin on (ordered curvature bits << 32 | vertex) in shared memory. The survivors are written in
//                    block order to a scratch row of the chunk.
//   placement        on the host from the per-chunk point counts: all chunks in submesh order (recompute_for_all_chunks),
//                    or, after a mesh sync, the invalidated chunks through the reference's RangeAllocator (smallest free
//                    range that fits, else the end) — the same code the synced mesh uses.
//   k_probe_scatter  scratch rows → their places in the persistent point buffer.
#include "mesh_sync.cuh"

struct ivx_probes {
    // chunk_point_ranges by linear chunk index: a flat table (start == NONE: the chunk has no points) instead of the
    // reference's HashMap — thousands of chunks are placed per call
    static constexpr uint32_t NONE = 0xFFFFFFFFu;
    std::vector<std::pair<uint32_t, uint32_t>> range_of_chunk;
    uint32_t n_with_points = 0;
    bool has(uint32_t chunk) const { return range_of_chunk[chunk].first != NONE; }
    void put(uint32_t chunk, uint32_t a, uint32_t b) {
        if (!has(chunk)) ++n_with_points;
        range_of_chunk[chunk] = {a, b};
    }
    void drop(uint32_t chunk) {
        if (has(chunk)) --n_with_points;
        range_of_chunk[chunk] = {NONE, NONE};
    }
    ivx_ranges::Ranges free_points;                                                // point_range_allocator
    uint32_t n_points = 0;    // length of the point buffer (holes included)
    uint32_t cap_points = 0;
    float* d_points = nullptr;
    uint32_t log2_block_size = 3;
};

void ivx_probes_free(ivx_ctx* ctx, ivx_probes* p) {
    if (!p) return;
    ctx->release(p->d_points);
    delete p;
}

namespace {

using namespace ivx_ranges;

struct ProbeArgs {
    const float* positions;
    const float* normals;
    const uint32_t* indices;
    const ivx_chunk_submesh* submeshes;
    const uint32_t* vertex_ranges;
    const uint32_t* rows;  // submesh rows to work on (null: all of them)
    uint32_t n_items;
    uint32_t log2_block_size;
    float inverse_voxel_extent;
    uint32_t nb1, nb2;
    float* scratch_points;  // n_items x blocks-per-chunk x 3
    uint32_t* counts;       // per item: points
    uint32_t* chunk_of;     // per item: linear chunk index
    uint32_t walk_all;      // 1: every chunk takes the path of chunks too large for the shared-memory lists (tests)
};

constexpr int PROBE_THREADS = 128;
constexpr uint32_t PROBE_MAX_VERTICES = 2048;  // of a chunk whose triangle corners are sorted by vertex in shared memory
constexpr uint32_t PROBE_MAX_CORNERS = 12288;  // (a chunk has ~300 vertices and ~1600 corners on average, at most 4913 / ~29 000)

__device__ __forceinline__ uint32_t ordered_bits(float f) {  // monotonic in f; -0 == +0
    uint32_t u = __float_as_uint(f == 0.0f ? 0.0f : f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(PROBE_THREADS) k_probe_points(ProbeArgs a) {
    extern __shared__ unsigned long long s_best[];  // per block of the chunk: ordered curvature << 32 | vertex
    __shared__ uint32_t s_warp[PROBE_THREADS / 32];
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_first[PROBE_MAX_VERTICES + 1];  // per vertex: where its corners start in s_corner
    __shared__ uint32_t s_fill[PROBE_MAX_VERTICES];
    __shared__ uint16_t s_corner[PROBE_MAX_CORNERS];      // positions in the chunk's index list, grouped by vertex
    const uint32_t log2_blocks = 4u - a.log2_block_size, n_blocks = 1u << (3u * log2_blocks);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const uint32_t row = a.rows ? a.rows[item] : item;
        const ivx_chunk_submesh sm = a.submeshes[row];
        const uint32_t v0 = a.vertex_ranges[2 * (size_t)row], v1 = a.vertex_ranges[2 * (size_t)row + 1];
        const uint32_t n_vertices = v1 - v0, n_triangles = sm.index_count / 3u;
        const uint32_t* idx = a.indices + sm.index_offset;
        for (uint32_t b = tid; b < n_blocks; b += PROBE_THREADS) s_best[b] = ~0ull;
        __syncthreads();
        float lower[3], upper[3];
        for (int d = 0; d < 3; ++d) {
            lower[d] = (float)(sm.chunk_indices[d] * 16u);
            upper[d] = (float)((sm.chunk_indices[d] + 1u) * 16u);
        }
        // Which triangle corners use a vertex, in index order. Chunks of ordinary size (PROBE_MAX_VERTICES vertices,
        // PROBE_MAX_CORNERS corners) sort the corners by vertex in shared memory — count, prefix sum, fill, then every
        // vertex orders its own few corners — so the work is linear in the chunk's mesh; larger chunks let every vertex
        // walk all triangles.
        const uint32_t n_corners = 3u * n_triangles;
        const bool listed = n_vertices <= PROBE_MAX_VERTICES && n_corners <= PROBE_MAX_CORNERS && !a.walk_all;
        if (listed) {
            for (uint32_t v = tid; v <= n_vertices; v += PROBE_THREADS) s_first[v] = 0u;
            __syncthreads();
            for (uint32_t q = tid; q < n_corners; q += PROBE_THREADS) atomicAdd(&s_first[idx[q] - v0], 1u);
            __syncthreads();
            // exclusive prefix sum over the vertices, PROBE_MAX_VERTICES / PROBE_THREADS consecutive entries per thread
            constexpr uint32_t PER = PROBE_MAX_VERTICES / PROBE_THREADS;
            uint32_t mine[PER], total = 0;
#pragma unroll
            for (uint32_t q = 0; q < PER; ++q) {
                const uint32_t v = tid * PER + q;
                mine[q] = v < n_vertices ? s_first[v] : 0u;
                total += mine[q];
            }
            uint32_t incl = total;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += y;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            uint32_t before = incl - total;
            for (int w = 0; w < warp; ++w) before += s_warp[w];
#pragma unroll
            for (uint32_t q = 0; q < PER; ++q) {
                const uint32_t v = tid * PER + q;
                if (v < n_vertices) {
                    s_first[v] = before;
                    s_fill[v] = before;
                }
                before += mine[q];
            }
            if (tid == PROBE_THREADS - 1) s_first[n_vertices] = before;
            __syncthreads();
            for (uint32_t q = tid; q < n_corners; q += PROBE_THREADS) s_corner[atomicAdd(&s_fill[idx[q] - v0], 1u)] = (uint16_t)q;
            __syncthreads();
        }
        for (uint32_t v = tid; v < n_vertices; v += PROBE_THREADS) {
            const uint32_t me = v0 + v;
            const float px = a.positions[3 * (size_t)me], py = a.positions[3 * (size_t)me + 1], pz = a.positions[3 * (size_t)me + 2];
            const float nx = a.normals[3 * (size_t)me], ny = a.normals[3 * (size_t)me + 1], nz = a.normals[3 * (size_t)me + 2];
            float sum = 0.0f, count = 0.0f;
            // corner c of a triangle: outgoing edge to corner c + 1, incoming edge from corner c - 1
            const auto add_corner = [&](uint32_t nxt, uint32_t prv) {
                const float ox = a.positions[3 * (size_t)nxt] - px, oy = a.positions[3 * (size_t)nxt + 1] - py,
                            oz = a.positions[3 * (size_t)nxt + 2] - pz;  // edge c -> c + 1
                const float ix = px - a.positions[3 * (size_t)prv], iy = py - a.positions[3 * (size_t)prv + 1],
                            iz = pz - a.positions[3 * (size_t)prv + 2];  // edge c - 1 -> c
                const float out_dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, ox), __fmul_rn(ny, oy)), __fmul_rn(nz, oz));
                const float in_dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, ix), __fmul_rn(ny, iy)), __fmul_rn(nz, iz));
                sum = __fadd_rn(sum, __fsub_rn(out_dot, in_dot));
                count += 2.0f;
            };
            if (listed) {
                const uint32_t b0 = s_first[v], b1 = s_first[v + 1];
                for (uint32_t q = b0 + 1; q < b1; ++q) {  // the fill order is arbitrary: insertion sort of a handful
                    const uint16_t key = s_corner[q];
                    uint32_t at = q;
                    while (at > b0 && s_corner[at - 1] > key) {
                        s_corner[at] = s_corner[at - 1];
                        --at;
                    }
                    s_corner[at] = key;
                }
                for (uint32_t q = b0; q < b1; ++q) {
                    const uint32_t corner = s_corner[q], t3 = corner - corner % 3u, c = corner % 3u;
                    add_corner(idx[t3 + (c + 1u) % 3u], idx[t3 + (c + 2u) % 3u]);
                }
            } else {
                for (uint32_t t = 0; t < n_triangles; ++t) {
                    const uint32_t tri[3] = {idx[3 * t], idx[3 * t + 1], idx[3 * t + 2]};
                    if (tri[0] != me && tri[1] != me && tri[2] != me) continue;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (tri[c] == me) add_corner(tri[(c + 1) % 3], tri[(c + 2) % 3]);
                }
            }
            if (count == 0.0f) continue;  // a vertex no triangle of the chunk uses
            uint32_t block[3];
            const float p3[3] = {px, py, pz};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float norm = __fmul_rn(p3[d], a.inverse_voxel_extent);
                const float clamped = fminf(fmaxf(norm, lower[d]), upper[d]);
                const uint32_t voxel = (uint32_t)clamped;  // `as usize`
                block[d] = (voxel & 15u) >> a.log2_block_size;  // clamped to the upper face: block 0, like the reference
            }
            const uint32_t b = (block[0] << (2u * log2_blocks)) + (block[1] << log2_blocks) + block[2];
            const float curvature = __fdiv_rn(sum, count);
            if (!(curvature < INFINITY)) continue;  // `curvature < min_curvature` never holds for these
            atomicMin(&s_best[b], ((unsigned long long)ordered_bits(curvature) << 32) | v);
        }
        __syncthreads();
        // the blocks that have a vertex, in block order
        float* out = a.scratch_points + (size_t)item * n_blocks * 3;
        if (tid == 0) s_base = 0;
        __syncthreads();
        for (uint32_t b0 = 0; b0 < n_blocks; b0 += PROBE_THREADS) {
            const uint32_t b = b0 + tid;
            const unsigned long long key = b < n_blocks ? s_best[b] : ~0ull;
            const bool has = key != ~0ull;
            const uint32_t ballot = __ballot_sync(0xffffffffu, has);
            if (lane == 0) s_warp[warp] = __popc(ballot);
            __syncthreads();
            uint32_t before = s_base;
            for (int w = 0; w < warp; ++w) before += s_warp[w];
            if (has) {
                const uint32_t at = before + __popc(ballot & ((1u << lane) - 1u));
                const uint32_t me = v0 + (uint32_t)(key & 0xFFFFFFFFull);
                out[3 * at] = a.positions[3 * (size_t)me];
                out[3 * at + 1] = a.positions[3 * (size_t)me + 1];
                out[3 * at + 2] = a.positions[3 * (size_t)me + 2];
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t all = 0;
                for (int w = 0; w < PROBE_THREADS / 32; ++w) all += s_warp[w];
                s_base += all;
            }
            __syncthreads();
        }
        if (tid == 0) {
            a.counts[item] = s_base;
            a.chunk_of[item] = (sm.chunk_indices[0] * a.nb1 + sm.chunk_indices[1]) * a.nb2 + sm.chunk_indices[2];
        }
        __syncthreads();
    }
}

// scratch rows → the point buffer; offsets[item] = 0xFFFFFFFF: nothing to copy
__global__ void k_probe_scatter(const float* __restrict__ scratch, const uint32_t* __restrict__ counts,
                                const uint32_t* __restrict__ offsets, uint32_t n_items, uint32_t blocks_per_chunk,
                                float* __restrict__ points) {
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t off = offsets[item];
        if (off == 0xFFFFFFFFu) continue;
        const uint32_t words = counts[item] * 3u;
        const float* src = scratch + (size_t)item * blocks_per_chunk * 3;
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) points[3 * (size_t)off + w] = src[w];
    }
}

int ensure_points(ivx_ctx* ctx, ivx_probes& pr, uint32_t keep, uint32_t want) {
    if (want <= pr.cap_points && pr.d_points) return IVX_OK;
    const uint32_t cap = want + want / 4 + 64;
    float* grown = static_cast<float*>(ctx->alloc((size_t)cap * 12));
    if (!grown) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "collision probes: out of device memory");
    if (pr.d_points && keep) CU(ctx, cudaMemcpyAsync(grown, pr.d_points, (size_t)keep * 12, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->release(pr.d_points);
    pr.d_points = grown;
    pr.cap_points = cap;
    return IVX_OK;
}

// determine_log2_block_size_for_object (collidable.rs:451-471)
uint32_t log2_block_size_for(const ivx_object* obj) {
    uint32_t min_extent = 0xFFFFFFFFu;
    for (int d = 0; d < 3; ++d) min_extent = std::min(min_extent, obj->occ_voxels[3 + d] - obj->occ_voxels[d]);
    return min_extent >= 16u ? 3u : (min_extent >= 8u ? 2u : (min_extent >= 4u ? 1u : 0u));
}

// runs k_probe_points over `rows` (null: all submeshes) and brings the counts back
int probe_items(ivx_ctx* ctx, ivx_object* obj, const uint32_t* d_rows, uint32_t n_items, uint32_t log2_bs, Tmp& tmp, float*& scratch,
                uint32_t*& d_counts, std::vector<uint32_t>& counts, std::vector<uint32_t>& chunk_of) {
    const DeviceMesh& m = obj->mesh;
    const uint32_t blocks_per_chunk = 1u << (3u * (4u - log2_bs));
    scratch = tmp.get<float>((size_t)n_items * blocks_per_chunk * 3);
    d_counts = tmp.get<uint32_t>(2 * (size_t)n_items);
    if (!scratch || !d_counts) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "collision probes: out of device memory");
    ProbeArgs a{};
    a.positions = m.positions;
    a.normals = m.normals;
    a.indices = m.indices;
    a.submeshes = m.submeshes;
    a.vertex_ranges = m.vertex_ranges;
    a.rows = d_rows;
    a.n_items = n_items;
    a.log2_block_size = log2_bs;
    a.inverse_voxel_extent = 1.0f / obj->voxel_extent;
    a.nb1 = obj->nb[1];
    a.nb2 = obj->nb[2];
    a.scratch_points = scratch;
    a.counts = d_counts;
    a.chunk_of = d_counts + n_items;
    a.walk_all = std::getenv("IVX_PROBES_WALK_ALL") ? 1u : 0u;
    const size_t smem = (size_t)blocks_per_chunk * 8;
    CU(ctx, cudaFuncSetAttribute(k_probe_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->launches++;
    k_probe_points<<<ivx_persistent_grid(ctx, n_items, 8), PROBE_THREADS, smem, ctx->stream>>>(a);
    CU(ctx, cudaGetLastError());
    counts.resize(n_items);
    chunk_of.resize(n_items);
    CU(ctx, cudaMemcpyAsync(counts.data(), d_counts, (size_t)n_items * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(chunk_of.data(), d_counts + n_items, (size_t)n_items * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IVX_OK;
}

int scatter_items(ivx_ctx* ctx, ivx_probes& pr, const float* scratch, const uint32_t* d_counts, const std::vector<uint32_t>& offsets,
                  uint32_t log2_bs, Tmp& tmp) {
    const uint32_t n_items = (uint32_t)offsets.size();
    if (n_items == 0) return IVX_OK;
    uint32_t* d_off = tmp.get<uint32_t>(n_items);
    if (!d_off) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "collision probes: out of device memory");
    CU(ctx, cudaMemcpyAsync(d_off, offsets.data(), (size_t)n_items * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches++;
    k_probe_scatter<<<ivx_persistent_grid(ctx, n_items, 8), 64, 0, ctx->stream>>>(scratch, d_counts, d_off, n_items,
                                                                                1u << (3u * (4u - log2_bs)), pr.d_points);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaStreamSynchronize(ctx->stream));  // `offsets` is the caller's
    return IVX_OK;
}

void fill_info(const ivx_probes& pr, ivx_probes_info* out) {
    out->log2_block_size = pr.log2_block_size;
    out->n_points = pr.n_points;
    out->n_chunks = pr.n_with_points;
    out->d_points = pr.d_points;
}

int check_object(ivx_ctx* ctx, const ivx_object* obj) {
    if (obj->derive_pending || obj->first_i != 0 || obj->nb[0] != obj->chunk_counts[0])
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "collision probes are kept for whole objects");
    if (obj->mesh_is_patch || !obj->mesh.positions)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "the object has no mesh: call ivx_object_mesh first (ivx_object_remesh_dirty "
                 "replaces the mesh by a patch)");
    return IVX_OK;
}

}  // namespace

extern "C" {

int ivx_object_collision_probes(ivx_ctx* ctx, ivx_object* obj, ivx_probes_info* out) {
    if (!ctx || !obj || !out) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    std::memset(out, 0, sizeof(*out));
    if (int rc = check_object(ctx, obj)) return rc;
    if (!obj->probes) {
        obj->probes = new (std::nothrow) ivx_probes();
        if (!obj->probes) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    ivx_probes& pr = *obj->probes;
    pr.range_of_chunk.assign(obj->n_chunks, {ivx_probes::NONE, ivx_probes::NONE});
    pr.n_with_points = 0;
    pr.free_points.clear();
    pr.n_points = 0;
    pr.log2_block_size = log2_block_size_for(obj);
    const uint32_t n_items = obj->mesh.n_submeshes;
    if (n_items == 0) {
        fill_info(pr, out);
        return IVX_OK;
    }
    Tmp tmp(ctx);
    float* scratch = nullptr;
    uint32_t* d_counts = nullptr;
    std::vector<uint32_t> counts, chunk_of;
    if (int rc = probe_items(ctx, obj, nullptr, n_items, pr.log2_block_size, tmp, scratch, d_counts, counts, chunk_of)) return rc;
    std::vector<uint32_t> offsets(n_items, 0xFFFFFFFFu);
    uint32_t total = 0;
    for (uint32_t q = 0; q < n_items; ++q) {
        if (counts[q] == 0u) continue;
        offsets[q] = total;
        pr.put(chunk_of[q], total, total + counts[q]);
        total += counts[q];
    }
    if (int rc = ensure_points(ctx, pr, 0, total)) return rc;
    pr.n_points = total;
    if (int rc = scatter_items(ctx, pr, scratch, d_counts, offsets, pr.log2_block_size, tmp)) return rc;
    fill_info(pr, out);
    return IVX_OK;
}

int ivx_object_collision_probes_sync(ivx_ctx* ctx, ivx_object* obj, ivx_probes_info* out) {
    if (!ctx || !obj || !out) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    std::memset(out, 0, sizeof(*out));
    if (int rc = check_object(ctx, obj)) return rc;
    if (!obj->probes || !obj->sync)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "collision probes follow a synced mesh: ivx_object_collision_probes once, then "
                 "ivx_object_mesh_sync + ivx_object_collision_probes_sync after every modification");
    ivx_probes& pr = *obj->probes;
    const ivx_mesh_sync& s = *obj->sync;
    pr.log2_block_size = log2_block_size_for(obj);
    const std::vector<uint32_t>& dirty = s.last_dirty;
    // the invalidated chunks that (still) have a submesh are probed; all of them are placed or removed in the order the
    // mesh sync took them
    std::vector<uint32_t> rows, item_of(dirty.size(), 0xFFFFFFFFu);
    for (size_t q = 0; q < dirty.size(); ++q) {
        auto it = s.row_of_chunk.find(dirty[q]);
        if (it == s.row_of_chunk.end()) continue;
        item_of[q] = (uint32_t)rows.size();
        rows.push_back(it->second);
    }
    Tmp tmp(ctx);
    float* scratch = nullptr;
    uint32_t* d_counts = nullptr;
    std::vector<uint32_t> counts, chunk_of;
    const uint32_t n_items = (uint32_t)rows.size();
    if (n_items) {
        uint32_t* d_rows = tmp.get<uint32_t>(n_items);
        if (!d_rows) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "collision probes: out of device memory");
        CU(ctx, cudaMemcpyAsync(d_rows, rows.data(), (size_t)n_items * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (int rc = probe_items(ctx, obj, d_rows, n_items, pr.log2_block_size, tmp, scratch, d_counts, counts, chunk_of)) return rc;
    }
    // update_for_chunk (collidable.rs:542-612)
    std::vector<uint32_t> offsets(n_items, 0xFFFFFFFFu);
    const uint32_t kept_points = pr.n_points;
    for (size_t q = 0; q < dirty.size(); ++q) {
        const uint32_t chunk = dirty[q];
        const uint32_t count = item_of[q] == 0xFFFFFFFFu ? 0u : counts[item_of[q]];
        const bool had = pr.has(chunk);
        const std::pair<uint32_t, uint32_t> old = pr.range_of_chunk[chunk];
        if (count == 0u) {
            if (had) {
                release_range(pr.free_points, old.first, old.second);
                pr.drop(chunk);
            }
            continue;
        }
        if (had) release_range(pr.free_points, old.first, old.second);
        uint32_t start = 0;
        if (!take_range(pr.free_points, count, start)) {
            start = pr.n_points;
            pr.n_points += count;
        }
        pr.put(chunk, start, start + count);
        offsets[item_of[q]] = start;
    }
    coalesce(pr.free_points);  // merge_consecutive_ranges
    if (int rc = ensure_points(ctx, pr, kept_points, pr.n_points)) return rc;
    if (int rc = scatter_items(ctx, pr, scratch, d_counts, offsets, pr.log2_block_size, tmp)) return rc;
    fill_info(pr, out);
    return IVX_OK;
}

int ivx_collision_probes_download(ivx_ctx* ctx, const ivx_object* obj, float* points, size_t capacity_points, ivx_probe_range* ranges,
                                  size_t capacity_ranges) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    const ivx_probes* pr = obj->probes;
    if (!pr) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "call ivx_object_collision_probes first");
    if (points) {
        if (capacity_points < pr->n_points) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u probe points", pr->n_points);
        if (pr->n_points) {
            CU(ctx, cudaMemcpyAsync(points, pr->d_points, (size_t)pr->n_points * 12, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    if (ranges) {
        if (capacity_ranges < pr->n_with_points)
            IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u chunk point ranges", pr->n_with_points);
        size_t q = 0;
        for (uint32_t c = 0; c < (uint32_t)pr->range_of_chunk.size(); ++c) {
            if (!pr->has(c)) continue;
            const auto& r = pr->range_of_chunk[c];
            ranges[q].chunk_indices[0] = c / (obj->nb[1] * obj->nb[2]);
            ranges[q].chunk_indices[1] = (c / obj->nb[2]) % obj->nb[1];
            ranges[q].chunk_indices[2] = c % obj->nb[2];
            ranges[q].point_start = r.first;
            ranges[q].point_end = r.second;
            ++q;
        }
    }
    return IVX_OK;
}

}  // extern "C"

// ---- mutual voxel-object contacts (collidable.rs:859-1050, 1288-1440) ----------------------------------------------------
// for_each_mutual_voxel_object_contact: the probes of one object that lie in the box of the intersection are taken to the
// other object's voxel space, where the distance field is sampled trilinearly (2x2x2 samples around the point) for the
// penetration depth and its gradient for the contact normal; deep inside (uniform chunk, or the clamped minimum distance)
// the direction from the centre of mass stands in for the normal. One thread per probe point; the contacts keep the
// order of the points (flag, prefix sum, compaction).
namespace {

struct MutualArgs2 {
    // the probing object
    const float* points;
    const uint32_t* range_first;  // per selected chunk: first point / first flattened position (n_ranges + 1 entries)
    const uint32_t* flat_first;
    uint32_t n_ranges, n_points;
    float q_from[4], t_from[3], inv_extent_from;
    float lo[3], hi[3];           // box of the intersection in the probing object's space, expanded
    // the probed object
    const DevChunk* chunks;
    const unsigned char* voxels;
    uint32_t nb1, nb2, dims[3];
    float q_into[4], t_into[3], extent_into, inv_extent_into, norm_center[3];
    uint32_t flip_normal;
    uint32_t* flag;
    ivx_voxel_contact* records;
};

__device__ __forceinline__ f3 rot(const float q[4], f3 v, bool inverse) {  // glam Quat::mul_vec3a (modify.cu quat_rotate)
    const float s = inverse ? -1.0f : 1.0f;
    const float bx = s * q[0], by = s * q[1], bz = s * q[2], w = q[3];
    const float b2 = (bx * bx + by * by) + bz * bz, vb = (v.x * bx + v.y * by) + v.z * bz;
    const float s1 = w * w - b2, s2 = vb * 2.0f, s3 = w * 2.0f;
    const float cx = by * v.z - bz * v.y, cy = bz * v.x - bx * v.z, cz = bx * v.y - by * v.x;
    return mk3((v.x * s1 + bx * s2) + cx * s3, (v.y * s1 + by * s2) + cy * s3, (v.z * s1 + bz * s2) + cz * s3);
}
__device__ __forceinline__ bool sign_set(float f) { return (__float_as_uint(f) >> 31) != 0u; }
__device__ __forceinline__ float voxel_distance(const MutualArgs2& a, uint32_t i, uint32_t j, uint32_t k) {
    const DevChunk c = a.chunks[((i >> 4) * a.nb1 + (j >> 4)) * a.nb2 + (k >> 4)];
    if (c.kind == 0) return sd_decode(127);
    if (c.kind == 1) return sd_decode((int)c.u_sd);
    const int8_t code = reinterpret_cast<const int8_t*>(a.voxels + (size_t)c.slot * SLOT_BYTES + PLANE_SD)[((i & 15u) << 8) | ((j & 15u) << 4) | (k & 15u)];
    return sd_decode((int)code);
}
__device__ __forceinline__ bool unit_if_above(f3 v, f3& out) {  // normalized_from_if_above(v, 1e-8)
    const float n2 = dot3(v, v);
    if (!(n2 > 1e-8f * 1e-8f)) return false;
    const float n = sqrtf(n2);
    out = mk3(v.x / n, v.y / n, v.z / n);
    return true;
}

__global__ void k_mutual_contacts(MutualArgs2 a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_points) return;
    a.flag[t] = 0u;
    // which selected chunk range this flattened position belongs to
    uint32_t lo_r = 0, hi_r = a.n_ranges;
    while (hi_r - lo_r > 1u) {
        const uint32_t mid = (lo_r + hi_r) >> 1;
        if (a.flat_first[mid] <= t) lo_r = mid;
        else hi_r = mid;
    }
    const uint32_t pt = a.range_first[lo_r] + (t - a.flat_first[lo_r]);
    const f3 pf = mk3(a.points[3 * (size_t)pt], a.points[3 * (size_t)pt + 1], a.points[3 * (size_t)pt + 2]);
    if (sign_set(pf.x - a.lo[0]) || sign_set(pf.y - a.lo[1]) || sign_set(pf.z - a.lo[2]) || sign_set(a.hi[0] - pf.x) ||
        sign_set(a.hi[1] - pf.y) || sign_set(a.hi[2] - pf.z))
        return;
    const f3 world = rot(a.q_from, mk3(pf.x - a.t_from[0], pf.y - a.t_from[1], pf.z - a.t_from[2]), true);
    const f3 ri = rot(a.q_into, world, false);
    const f3 p = mk3((ri.x + a.t_into[0]) * a.inv_extent_into, (ri.y + a.t_into[1]) * a.inv_extent_into, (ri.z + a.t_into[2]) * a.inv_extent_into);
    // determine_sdf_value_and_normal_at_point_if_intersecting
    const float MIN_SD = 0.02f * -128.0f, HALF_DIAGONAL = 0.5f * 1.7320508f;
    const f3 nc = mk3(a.norm_center[0], a.norm_center[1], a.norm_center[2]);
    const f3 lower = mk3(p.x - 0.5f, p.y - 0.5f, p.z - 0.5f);
    if (sign_set(lower.x) || sign_set(lower.y) || sign_set(lower.z)) return;
    if (!(lower.x < 4.0e9f && lower.y < 4.0e9f && lower.z < 4.0e9f)) return;  // (`as usize` saturates: beyond any grid; NaN too)
    const uint32_t li = (uint32_t)lower.x, lj = (uint32_t)lower.y, lk = (uint32_t)lower.z;
    if ((li + 1u >= a.dims[0]) | (lj + 1u >= a.dims[1]) | (lk + 1u >= a.dims[2])) return;
    const uint32_t ci = (uint32_t)p.x, cj = (uint32_t)p.y, ck = (uint32_t)p.z;
    const DevChunk chunk = a.chunks[((ci >> 4) * a.nb1 + (cj >> 4)) * a.nb2 + (ck >> 4)];
    float sd = MIN_SD;
    f3 normal;
    bool deep = chunk.kind == 1;
    if (chunk.kind == 0) return;
    if (!deep) {
        if (voxel_distance(a, ci, cj, ck) > HALF_DIAGONAL) return;
        const float d[8] = {voxel_distance(a, li, lj, lk),         voxel_distance(a, li, lj, lk + 1),
                            voxel_distance(a, li, lj + 1, lk),     voxel_distance(a, li, lj + 1, lk + 1),
                            voxel_distance(a, li + 1, lj, lk),     voxel_distance(a, li + 1, lj, lk + 1),
                            voxel_distance(a, li + 1, lj + 1, lk), voxel_distance(a, li + 1, lj + 1, lk + 1)};
        const f3 o = mk3(lower.x - floorf(lower.x), lower.y - floorf(lower.y), lower.z - floorf(lower.z));
        const f3 r = mk3(1.0f - o.x, 1.0f - o.y, 1.0f - o.z);
        // evaluate_sdf_from_corner_samples (object/sdf.rs:579-592)
        const float d00 = d[0] * r.x + d[4] * o.x, d01 = d[1] * r.x + d[5] * o.x;
        const float d10 = d[2] * r.x + d[6] * o.x, d11 = d[3] * r.x + d[7] * o.x;
        const float d0 = d00 * r.y + d10 * o.y, d1 = d01 * r.y + d11 * o.y;
        sd = d0 * r.z + d1 * o.z;
        if (sd > 0.0f) return;
        if (fabsf(sd - MIN_SD) < 1e-3f) {
            deep = true;
        } else {
            // compute_sdf_gradient_from_corner_samples (object/sdf.rs:603-633)
            const f3 e00 = mk3(d[4] - d[0], d[2] - d[0], d[1] - d[0]), e01 = mk3(d[5] - d[1], d[6] - d[4], d[3] - d[2]);
            const f3 e10 = mk3(d[6] - d[2], d[3] - d[1], d[5] - d[4]), e11 = mk3(d[7] - d[3], d[7] - d[5], d[7] - d[6]);
            f3 g;
            g.x = (((r.y * r.z) * e00.x + (r.y * o.z) * e01.x) + (o.y * r.z) * e10.x) + (o.y * o.z) * e11.x;
            g.y = (((r.z * r.x) * e00.y + (r.z * o.x) * e01.y) + (o.z * r.x) * e10.y) + (o.z * o.x) * e11.y;
            g.z = (((r.x * r.y) * e00.z + (r.x * o.y) * e01.z) + (o.x * r.y) * e10.z) + (o.x * o.y) * e11.z;
            if (!unit_if_above(g, normal)) return;
        }
    }
    if (deep) {  // estimate_sdf_value_and_normal_at_point_deep_inside
        sd = MIN_SD;
        if (!unit_if_above(mk3(p.x - nc.x, p.y - nc.y, p.z - nc.z), normal)) return;
    }
    f3 n = rot(a.q_into, normal, true);
    if (a.flip_normal) n = mk3(-n.x, -n.y, -n.z);
    ivx_voxel_contact rec;
    rec.indices[0] = (uint32_t)(pf.x * a.inv_extent_from);
    rec.indices[1] = (uint32_t)(pf.y * a.inv_extent_from);
    rec.indices[2] = (uint32_t)(pf.z * a.inv_extent_from);
    rec.position[0] = world.x;
    rec.position[1] = world.y;
    rec.position[2] = world.z;
    rec.surface_normal[0] = n.x;
    rec.surface_normal[1] = n.y;
    rec.surface_normal[2] = n.z;
    rec.penetration_depth = -sd * a.extent_into;
    a.records[t] = rec;
    a.flag[t] = 1u;
}

__global__ void k_compact_contacts(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ scan, const ivx_voxel_contact* __restrict__ in,
                                   uint32_t n, uint32_t capacity, ivx_voxel_contact* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && flag[t] && scan[t] < capacity) out[scan[t]] = in[t];
}

// the probes of `from` against the distance field of `into` → contacts appended to d_out (device) from `base` on
int probes_against(ivx_ctx* ctx, const ivx_object* from, const ivx_isometry* world_to_from, const ivx_object* into,
                   const ivx_inertial_moments* inertial_into, const ivx_isometry* world_to_into, const uint32_t ranges_in_from[6],
                   float aabb_margin, bool flip_normal, ivx_voxel_contact* d_out, uint32_t base, uint32_t capacity, uint32_t* out_count) {
    *out_count = 0;
    const ivx_probes& pr = *from->probes;
    uint32_t cr[3][2];
    for (int d = 0; d < 3; ++d) {
        cr[d][0] = ranges_in_from[2 * d] / 16u;
        cr[d][1] = std::min(from->nb[d], (ranges_in_from[2 * d + 1] + 15u) / 16u);
    }
    std::vector<uint32_t> first, flat;
    uint32_t total = 0;
    for (uint32_t i = cr[0][0]; i < cr[0][1]; ++i)
        for (uint32_t j = cr[1][0]; j < cr[1][1]; ++j)
            for (uint32_t k = cr[2][0]; k < cr[2][1]; ++k) {
                const uint32_t c = (i * from->nb[1] + j) * from->nb[2] + k;
                if (c >= pr.range_of_chunk.size() || !pr.has(c)) continue;
                first.push_back(pr.range_of_chunk[c].first);
                flat.push_back(total);
                total += pr.range_of_chunk[c].second - pr.range_of_chunk[c].first;
            }
    if (total == 0) return IVX_OK;
    flat.push_back(total);
    first.push_back(0);
    Tmp tmp(ctx);
    const uint32_t n_ranges = (uint32_t)first.size() - 1u;
    uint32_t* d_first = tmp.get<uint32_t>(first.size());
    uint32_t* d_flat = tmp.get<uint32_t>(flat.size());
    uint32_t* flag = tmp.get<uint32_t>(total);
    uint32_t* scan = tmp.get<uint32_t>(total);
    ivx_voxel_contact* recs = tmp.get<ivx_voxel_contact>(total);
    if (!d_first || !d_flat || !flag || !scan || !recs) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mutual contacts: out of device memory");
    cudaStream_t st = ctx->stream;
    CU(ctx, cudaMemcpyAsync(d_first, first.data(), first.size() * 4, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(d_flat, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, st));
    MutualArgs2 a{};
    a.points = pr.d_points;
    a.range_first = d_first;
    a.flat_first = d_flat;
    a.n_ranges = n_ranges;
    a.n_points = total;
    for (int q = 0; q < 4; ++q) {
        a.q_from[q] = world_to_from->rotation[q];
        a.q_into[q] = world_to_into->rotation[q];
    }
    const float inv_into = 1.0f / into->voxel_extent;
    for (int d = 0; d < 3; ++d) {
        a.t_from[d] = world_to_from->translation[d];
        a.t_into[d] = world_to_into->translation[d];
        a.lo[d] = from->voxel_extent * (float)ranges_in_from[2 * d] - aabb_margin;
        a.hi[d] = from->voxel_extent * (float)ranges_in_from[2 * d + 1] + aabb_margin;
        a.dims[d] = into->nb[d] * 16u;
        // derive_center_of_mass() * inverse_voxel_extent
        a.norm_center[d] = (inertial_into->moments[d] / inertial_into->mass) * inv_into;
    }
    a.inv_extent_from = 1.0f / from->voxel_extent;
    a.chunks = into->d_chunks;
    a.voxels = into->d_voxels;
    a.nb1 = into->nb[1];
    a.nb2 = into->nb[2];
    a.extent_into = into->voxel_extent;
    a.inv_extent_into = inv_into;
    a.flip_normal = flip_normal ? 1u : 0u;
    a.flag = flag;
    a.records = recs;
    ctx->launches++;
    k_mutual_contacts<<<(total + 127) / 128, 128, 0, st>>>(a);
    CU(ctx, cudaGetLastError());
    KL(ctx, launch_exclusive_scan(flag, scan, total, ctx->d_scratch + 30, st));
    ctx->launches++;
    k_compact_contacts<<<(total + 255) / 256, 256, 0, st>>>(flag, scan, recs, total, capacity > base ? capacity - base : 0u, d_out + base);
    CU(ctx, cudaGetLastError());
    uint32_t cnt = 0;
    if (int rc = ivx_read_words(ctx, ctx->d_scratch + 30, 1, &cnt)) return rc;  // (also keeps `first` / `flat` alive long enough)
    *out_count = cnt;
    return IVX_OK;
}

}  // namespace

extern "C" int ivx_objects_mutual_contacts(ivx_ctx* ctx, const ivx_object* a, const ivx_object* b, const ivx_isometry* world_to_a,
                                           const ivx_isometry* world_to_b, const uint32_t ranges_in_a[6], const uint32_t ranges_in_b[6],
                                           const ivx_inertial_moments* inertial_a, const ivx_inertial_moments* inertial_b,
                                           ivx_voxel_contact* out, size_t capacity, uint64_t* out_count_a_against_b,
                                           uint64_t* out_count_b_against_a) {
    if (!ctx || !a || !b || !world_to_a || !world_to_b || !ranges_in_a || !ranges_in_b || !inertial_a || !inertial_b ||
        !out_count_a_against_b || !out_count_b_against_a)
        return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    *out_count_a_against_b = *out_count_b_against_a = 0;
    if (!a->probes || !b->probes) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "both objects need collision probes: ivx_object_collision_probes");
    for (const ivx_object* o : {a, b})
        if (o->derive_pending || o->first_i != 0 || o->nb[0] != o->chunk_counts[0])
            IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "contact queries on a slab-partitioned object are not supported");
    const uint32_t cap = (uint32_t)std::min<size_t>(capacity, 0xFFFFFFFFu);
    Tmp tmp(ctx);
    ivx_voxel_contact* d_out = tmp.get<ivx_voxel_contact>(std::max<size_t>(1, cap));
    if (!d_out) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mutual contacts: out of device memory");
    uint32_t n_ab = 0, n_ba = 0;
    if (int rc = probes_against(ctx, a, world_to_a, b, inertial_b, world_to_b, ranges_in_a, a->voxel_extent, false, d_out, 0, cap, &n_ab))
        return rc;
    // (the reference expands B's box by A's voxel extent too, collidable.rs:969-973)
    if (int rc = probes_against(ctx, b, world_to_b, a, inertial_a, world_to_a, ranges_in_b, a->voxel_extent, true, d_out,
                                std::min(n_ab, cap), cap, &n_ba))
        return rc;
    *out_count_a_against_b = n_ab;
    *out_count_b_against_a = n_ba;
    if ((uint64_t)n_ab + n_ba > cap) return out ? IVX_ERR_CAPACITY : IVX_OK;
    if (out && n_ab + n_ba) {
        CU(ctx, cudaMemcpyAsync(out, d_out, (size_t)(n_ab + n_ba) * sizeof(ivx_voxel_contact), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return IVX_OK;
}

