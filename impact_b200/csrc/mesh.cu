// Surface Nets meshing kernels (sm_100a).
//
// Replaces, per exposed chunk,
//   VoxelObject::fill_sdf_for_chunk_if_exposed   (object/sdf.rs:181-508)
//   compute_surface_nets_mesh                    (object/sdf/surface_nets.rs:131-381)
//   calculate_all_index_materials                (surface_nets.rs:540-637)
// and, across chunks, VoxelObjectMesh::recreate  (mesh.rs:286-354).
//
// One CTA per exposed chunk. The 18³ brick (chunk + 1-voxel halo gathered from
// up to 26 neighbours) is staged in shared memory as signed-distance codes and
// voxel types (2 B per cell, 11.7 KB) and decoded on use. Vertices are
// emitted in the reference's i→j→k cube order and quads in vertex order
// (x, y, z axis per vertex) through ballot / prefix-sum compaction, so the
// output buffers are identical in ORDER, not just in content. The kernel runs
// twice: a counting pass sizes every chunk's vertex / index ranges, an
// exclusive scan over the chunk list (linear chunk order) places them, and the
// emit pass writes positions, normals, indices and index materials in place.
#include "common.cuh"
#include "kernels.h"

namespace ivx {

constexpr int MESH_THREADS = 256;
constexpr int N_CUBES = 17 * 17 * 17;
constexpr int CUBES_PER_THREAD = (N_CUBES + MESH_THREADS - 1) / MESH_THREADS;  // 20

// SurfaceNetsVertexMaterials (surface_nets.rs:440-451) packed into registers: byte q of `idx` / `wgt` is
// indices[q] / weights[q]; indices[7] holds the material count.
struct VertexMaterials {
    uint64_t idx, wgt;
};

__device__ __forceinline__ uint32_t byte_of(uint64_t v, int q) { return (uint32_t)(v >> (8 * q)) & 0xFFu; }

// SurfaceNetsVertexMaterials::compute + sort_descending (surface_nets.rs:453-538)
// `types`: byte c = voxel type at cube corner c (CUBE_CORNERS order, surface_nets.rs:639-648)
__device__ __noinline__ VertexMaterials vertex_materials(uint32_t neg_mask, uint64_t types) {
    uint64_t idx = 0, wgt = 0;
    int count = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if ((neg_mask >> c) & 1u) {
            const uint32_t m = (uint32_t)(types >> (8 * c)) & 0xFFu;
            int found = -1;
#pragma unroll
            for (int q = 0; q < 7; ++q)
                if (q < count && found < 0 && byte_of(idx, q) == m) found = q;
            if (found < 0) {
                idx |= (uint64_t)m << (8 * count);
                wgt |= 1ull << (8 * count);
                count++;
            } else {
                wgt += 1ull << (8 * found);
            }
        }
    }
    idx |= (uint64_t)count << 56;
    const int NET[17][2] = {{0, 6}, {1, 5}, {2, 4}, {0, 3}, {1, 2}, {4, 5}, {0, 1}, {2, 3}, {4, 6},
                            {5, 6}, {1, 4}, {3, 5}, {1, 2}, {3, 4}, {5, 6}, {2, 3}, {4, 5}};
#pragma unroll
    for (int s = 0; s < 17; ++s) {
        const int i = NET[s][0], j = NET[s][1];
        const uint32_t wi = byte_of(wgt, i), wj = byte_of(wgt, j);
        if (wi < wj) {
            const uint64_t xw = (uint64_t)(wi ^ wj);
            wgt ^= (xw << (8 * i)) | (xw << (8 * j));
            const uint64_t xi = (uint64_t)(byte_of(idx, i) ^ byte_of(idx, j));
            idx ^= (xi << (8 * i)) | (xi << (8 * j));
        }
    }
    return VertexMaterials{idx, wgt};
}

// calculate_index_materials_for_triangle (surface_nets.rs:556-637); out = indices[4] | weights[4] << 32
__device__ __noinline__ void triangle_index_materials(const VertexMaterials vm[3], uint64_t out[3]) {
    const uint32_t cnt[3] = {byte_of(vm[0].idx, 7), byte_of(vm[1].idx, 7), byte_of(vm[2].idx, 7)};
    if (cnt[0] == 1 && cnt[1] == 1 && cnt[2] == 1) {
        const uint32_t index = byte_of(vm[0].idx, 0);
        if (byte_of(vm[1].idx, 0) == index && byte_of(vm[2].idx, 0) == index) {
            out[0] = out[1] = out[2] = (uint64_t)index | (1ull << 32);
            return;
        }
    }
    uint32_t top = 0;  // 4 packed bytes
    int n_top = 0;
    int off[3] = {0, 0, 0};
    for (int t = 0; t < 4; ++t) {
        uint32_t w[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) w[i] = byte_of(vm[i].wgt, off[i]);
        const int mx = (w[0] >= w[1]) ? ((w[0] >= w[2]) ? 0 : 2) : ((w[1] >= w[2]) ? 1 : 2);
        const uint32_t wmx = mx == 0 ? w[0] : (mx == 1 ? w[1] : w[2]);
        if (wmx == 0) break;
        const uint64_t imx = mx == 0 ? vm[0].idx : (mx == 1 ? vm[1].idx : vm[2].idx);
        const int omx = mx == 0 ? off[0] : (mx == 1 ? off[1] : off[2]);
        top |= byte_of(imx, omx) << (8 * t);
        n_top++;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            for (;;) {
                if (off[i] >= (int)cnt[i]) break;
                const uint32_t cand = byte_of(vm[i].idx, off[i]);
                bool is_top = false;
                for (int q = 0; q < n_top; ++q) is_top = is_top || (((top >> (8 * q)) & 0xFFu) == cand);
                if (!is_top) break;
                off[i]++;
            }
        }
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        uint32_t weights = 0;
        for (int i = 0; i < n_top; ++i) {
            const uint32_t want = (top >> (8 * i)) & 0xFFu;
            for (int j = 0; j < (int)cnt[v]; ++j)
                if (byte_of(vm[v].idx, j) == want) {
                    weights |= byte_of(vm[v].wgt, j) << (8 * i);
                    break;
                }
        }
        out[v] = (uint64_t)top | ((uint64_t)weights << 32);
    }
}

__device__ __forceinline__ int corner_off(int c) { return ((c >> 2) & 1) * 324 + ((c >> 1) & 1) * 18 + (c & 1); }

// block-wide exclusive scan of one small value per thread; returns the
// exclusive prefix and adds the block total to `running`
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < MESH_THREADS / 32; ++w) {
        const uint32_t s = s_warp[w];
        if (w < warp) woff += s;
        tot += s;
    }
    total = tot;
    return woff + x - v;
}

// what the material merge needs of one surface vertex: the signs and voxel types of its cube's 8 corners
__device__ __forceinline__ void vertex_corner_data(int lin, const uint32_t* s_neg, const uint8_t* s_type, uint32_t& neg,
                                                   uint64_t& types) {
    const int r = (lin / 324) * 18 + (lin / 18) % 18, k = lin % 18;
    neg = ((s_neg[r] >> k) & 3u) | (((s_neg[r + 1] >> k) & 3u) << 2) | (((s_neg[r + 18] >> k) & 3u) << 4) |
          (((s_neg[r + 19] >> k) & 3u) << 6);
    types = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) types |= (uint64_t)s_type[lin + corner_off(c)] << (8 * c);
}

// index materials of the two triangles of a quad whose corner vertices carry several materials
// (calculate_all_index_materials, surface_nets.rs:540-637, slow path)
__device__ __forceinline__ void emit_quad_materials(const uint32_t neg[4], const uint64_t types[4], uint32_t order, uint64_t* IM) {
    VertexMaterials vm[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) vm[c] = vertex_materials(neg[c], types[c]);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        VertexMaterials tv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t sel = (order >> (3 * (3 * t + c))) & 3u;
            tv[c] = sel == 0 ? vm[0] : (sel == 1 ? vm[1] : (sel == 2 ? vm[2] : vm[3]));
        }
        uint64_t im[3];
        triangle_index_materials(tv, im);
#pragma unroll
        for (int c = 0; c < 3; ++c) IM[3 * t + c] = im[c];
    }
}

// One thread per recorded multi-material quad. Entry layout (3 x uint4): types of the four vertices' cube corners
// (4 x u64), then {first index of the quad in the object's index buffer, corner order | neg masks, -, -}.
__global__ void __launch_bounds__(128) k_mesh_materials(const uint4* __restrict__ entries, const uint32_t* __restrict__ count,
                                                         uint32_t capacity, ivx_index_materials* __restrict__ index_materials) {
    const uint32_t n = min(*count, capacity);
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint4 a = entries[3 * (size_t)e], b = entries[3 * (size_t)e + 1], c = entries[3 * (size_t)e + 2];
        const uint64_t types[4] = {(uint64_t)a.x | ((uint64_t)a.y << 32), (uint64_t)a.z | ((uint64_t)a.w << 32),
                                   (uint64_t)b.x | ((uint64_t)b.y << 32), (uint64_t)b.z | ((uint64_t)b.w << 32)};
        const uint32_t neg[4] = {c.z & 0xFFu, (c.z >> 8) & 0xFFu, (c.z >> 16) & 0xFFu, c.z >> 24};
        emit_quad_materials(neg, types, c.y, reinterpret_cast<uint64_t*>(index_materials + c.x));
    }
}

template <bool EMIT>
__global__ void __launch_bounds__(MESH_THREADS, 4) k_mesh(MeshArgs a) {
    __shared__ __align__(16) int8_t s_sd[5832];
    __shared__ __align__(16) uint8_t s_type[EMIT ? 5832 : 16];
    __shared__ __align__(4) uint16_t s_l2v[EMIT ? 5832 : 2];
    __shared__ uint32_t s_neg[324];
    __shared__ uint16_t s_surf[N_CUBES];
    __shared__ uint8_t s_vmat[EMIT ? N_CUBES : 1];
    __shared__ uint32_t s_warp[MESH_THREADS / 32];
    __shared__ uint32_t s_adj_up[3];

    const int tid = threadIdx.x;
    for (uint32_t w = blockIdx.x; w < a.n_work; w += gridDim.x) {
        const uint32_t chunk = a.work[w];
        const uint32_t ck = chunk % a.nb[2], cj = (chunk / a.nb[2]) % a.nb[1], ci = chunk / (a.nb[2] * a.nb[1]);

        // ---- stage the 18³ brick ----
        if (tid < 3) {
            int n[3] = {(int)ci, (int)cj, (int)ck};
            n[tid] += 1;
            uint32_t up = 0;
            if (n[0] < (int)a.nb[0] && n[1] < (int)a.nb[1] && n[2] < (int)a.nb[2])
            {
                const DevChunk nc = a.chunks[(n[0] * a.nb[1] + n[1]) * a.nb[2] + n[2]];
                up = (nc.kind == 2 || (nc.kind == 1 && nc.pre == PRE_CONVERTED_HALO)) ? 1u : 0u;
            }
            s_adj_up[tid] = up;
        }
        // interior 16³: the chunk's own planes, one 16-byte row (i, j, 0..15) per thread and plane
        {
            const DevChunk me = a.chunks[chunk];  // NonUniform by construction of the work list
            const unsigned char* slot = a.voxels + (size_t)me.slot * SLOT_BYTES;
            const uint4 wsd = *reinterpret_cast<const uint4*>(slot + PLANE_SD + tid * 16);
            const uint32_t ws[4] = {wsd.x, wsd.y, wsd.z, wsd.w};
            const int row = bidx((tid >> 4) + 1, (tid & 15) + 1, 1);
#pragma unroll
            for (int k = 0; k < 16; ++k) s_sd[row + k] = (int8_t)((ws[k >> 2] >> (8 * (k & 3))) & 0xFFu);
            if (EMIT) {  // the counting pass only needs signs
                const uint4 wty = *reinterpret_cast<const uint4*>(slot + PLANE_TYPE + tid * 16);
                const uint32_t wt[4] = {wty.x, wty.y, wty.z, wty.w};
#pragma unroll
                for (int k = 0; k < 16; ++k) s_type[row + k] = (uint8_t)((wt[k >> 2] >> (8 * (k & 3))) & 0xFFu);
            }
        }
        // the 1-voxel halo (1736 cells) from the up to 26 neighbours (object/sdf.rs:410-508): the two i planes, then
        // the two j planes without their i borders, then the two k planes without their i and j borders
        for (int h = tid; h < 1736; h += MESH_THREADS) {
            int bi, bj, bk;
            if (h < 648) {
                bi = (h / 324) * 17;
                bj = (h % 324) / 18;
                bk = h % 18;
            } else if (h < 1224) {
                const int r = (h - 648) % 288;
                bi = 1 + r / 18;
                bj = ((h - 648) / 288) * 17;
                bk = r % 18;
            } else {
                const int r = (h - 1224) % 256;
                bi = 1 + (r >> 4);
                bj = 1 + (r & 15);
                bk = ((h - 1224) / 256) * 17;
            }
            const int cell = bidx(bi, bj, bk);
            const int gi = (int)ci * 16 + bi - 1, gj = (int)cj * 16 + bj - 1, gk = (int)ck * 16 + bk - 1;
            int8_t sd = 127;
            uint8_t ty = 255;
            if (gi >= 0 && gj >= 0 && gk >= 0) {
                const uint32_t ni = gi >> 4, nj = gj >> 4, nk = gk >> 4;
                if (ni < a.nb[0] && nj < a.nb[1] && nk < a.nb[2]) {
                    const DevChunk nc = a.chunks[(ni * a.nb[1] + nj) * a.nb[2] + nk];
                    if (nc.kind == 1) {
                        sd = -128;
                        ty = nc.u_type;
                    } else if (nc.kind == 2) {
                        const unsigned char* slot = a.voxels + (size_t)nc.slot * SLOT_BYTES;
                        const int v = vidx(gi & 15, gj & 15, gk & 15);
                        sd = (int8_t)slot[PLANE_SD + v];
                        if (EMIT) ty = slot[PLANE_TYPE + v];
                    }
                }
            }
            s_sd[cell] = sd;
            if (EMIT) s_type[cell] = ty;
        }
        if (EMIT)
            for (int cell = tid; cell < 5832 / 2; cell += MESH_THREADS) reinterpret_cast<uint32_t*>(s_l2v)[cell] = 0xFFFFFFFFu;
        __syncthreads();
        // sign bits per brick row (i, j): bit k = cell (i, j, k) is negative
        for (int r = tid; r < 324; r += MESH_THREADS) {
            uint32_t m = 0;
#pragma unroll
            for (int k = 0; k < 18; ++k) m |= (s_sd[r * 18 + k] < 0 ? 1u : 0u) << k;
            s_neg[r] = m;
        }
        __syncthreads();

        const float extent = a.voxel_extent;
        const float chunk_extent = extent * 16.0f;
        // vertex_position_offset_for_chunk (mesh.rs:559-577); chunk index in the FULL grid
        const f3 offset = mk3((float)(ci + a.first_i) * chunk_extent - 0.5f * extent,
                              (float)cj * chunk_extent - 0.5f * extent, (float)ck * chunk_extent - 0.5f * extent);
        const uint32_t voff = EMIT ? a.vertex_offset[w] : 0u;
        const uint32_t ioff = EMIT ? a.index_offset[w] : 0u;
        bool skip_emit = EMIT && a.index_count[w] == 0u;  // empty mesh: nothing is appended (mesh.rs:319-321)
        if (EMIT && a.cap_indices != 0u && !skip_emit &&
            (voff + a.vertex_count[w] > a.cap_vertices || ioff + a.index_count[w] > a.cap_indices ||
             a.submesh_ord[w] >= a.cap_submeshes))
            skip_emit = true;  // buffers sized from a stale plan: stay inside them, k_check_plan flags the call

        // ---- vertex pass: estimate_surface_nets_surface (surface_nets.rs:152-244) ----
        // 1) compact the surface cubes in cube order (i → j → k), which is the reference's vertex order:
        //    each thread owns CUBES_PER_THREAD consecutive cubes, one block scan places them
        uint32_t n_vertices = 0;
        {
            const int q0 = tid * CUBES_PER_THREAD;
            int i = q0 / 289, j = (q0 / 17) % 17, k = q0 % 17;
            const int i0 = i, j0 = j, k0 = k;
            uint32_t smask = 0;
#pragma unroll 4
            for (int u = 0; u < CUBES_PER_THREAD; ++u) {
                if (q0 + u < N_CUBES) {
                    // the cube's 8 corner signs: bits k, k+1 of the four rows (i, j), (i, j+1), (i+1, j), (i+1, j+1)
                    const int r = i * 18 + j;
                    const uint32_t neg = ((s_neg[r] >> k) & 3u) | (((s_neg[r + 1] >> k) & 3u) << 2) |
                                         (((s_neg[r + 18] >> k) & 3u) << 4) | (((s_neg[r + 19] >> k) & 3u) << 6);
                    if (neg != 0u && neg != 0xFFu) smask |= 1u << u;
                }
                if (++k == 17) {
                    k = 0;
                    if (++j == 17) {
                        j = 0;
                        ++i;
                    }
                }
            }
            uint32_t v = block_exclusive_scan(__popc(smask), s_warp, n_vertices);
            i = i0, j = j0, k = k0;
            for (int u = 0; u < CUBES_PER_THREAD && smask; ++u) {
                if ((smask >> u) & 1u) {
                    const int lin = bidx(i, j, k);
                    s_surf[v] = (uint16_t)lin;
                    if (EMIT) s_l2v[lin] = (uint16_t)v;
                    ++v;
                }
                if (++k == 17) {
                    k = 0;
                    if (++j == 17) {
                        j = 0;
                        ++i;
                    }
                }
            }
        }
        __syncthreads();

        // 2) one thread per surface vertex (dense lanes): position, normal, material summary
        if (EMIT && !skip_emit) {
            for (uint32_t v = tid; v < n_vertices; v += MESH_THREADS) {
                const int lin = s_surf[v];
                const int i = lin / 324, j = (lin / 18) % 18, k = lin % 18;
                float d[8];
                uint32_t neg = 0;
                uint32_t mat_first = 256u;
                bool single = true;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int code = (int)s_sd[lin + corner_off(c)];
                    d[c] = sd_decode(code);
                    if (code < 0) {
                        neg |= 1u << c;
                        const uint32_t m = s_type[lin + corner_off(c)];
                        if (mat_first == 256u) mat_first = m;
                        single = single && (m == mat_first);
                    }
                }
                s_vmat[v] = single ? (uint8_t)mat_first : (uint8_t)255;  // 255 = several materials (types are < 255)
                // centroid_of_edge_intersections (surface_nets.rs:384-418)
                const int E[12][2] = {{0, 1}, {0, 2}, {0, 4}, {1, 3}, {1, 5}, {2, 3},
                                      {2, 6}, {3, 7}, {4, 5}, {4, 6}, {5, 7}, {6, 7}};
                int count = 0;
                f3 sum = mk3(0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int e = 0; e < 12; ++e) {
                    const int c1 = E[e][0], c2 = E[e][1];
                    if (((neg >> c1) ^ (neg >> c2)) & 1u) {
                        count++;
                        const float interp1 = d[c1] / (d[c1] - d[c2]);
                        const float interp2 = 1.0f - interp1;
                        f3 p1 = mk3((float)((c1 >> 2) & 1), (float)((c1 >> 1) & 1), (float)(c1 & 1));
                        f3 p2 = mk3((float)((c2 >> 2) & 1), (float)((c2 >> 1) & 1), (float)(c2 & 1));
                        sum = sum + (interp2 * p1 + interp1 * p2);
                    }
                }
                const float fc = (float)count;
                const f3 o = mk3(sum.x / fc, sum.y / fc, sum.z / fc);
                // compute_sdf_gradient_from_corner_samples (object/sdf.rs:603-633)
                const f3 r = mk3(1.0f - o.x, 1.0f - o.y, 1.0f - o.z);
                const f3 d00 = mk3(d[4] - d[0], d[2] - d[0], d[1] - d[0]);
                const f3 d01 = mk3(d[5] - d[1], d[6] - d[4], d[3] - d[2]);
                const f3 d10 = mk3(d[6] - d[2], d[3] - d[1], d[5] - d[4]);
                const f3 d11 = mk3(d[7] - d[3], d[7] - d[5], d[7] - d[6]);
                // rev.yzx * rev.zxy * d00 + rev.yzx * o.zxy * d01 + o.yzx * rev.zxy * d10 + o.yzx * o.zxy * d11
                f3 g;
                g.x = (((r.y * r.z) * d00.x + (r.y * o.z) * d01.x) + (o.y * r.z) * d10.x) + (o.y * o.z) * d11.x;
                g.y = (((r.z * r.x) * d00.y + (r.z * o.x) * d01.y) + (o.z * r.x) * d10.y) + (o.z * o.x) * d11.y;
                g.z = (((r.x * r.y) * d00.z + (r.x * o.y) * d01.z) + (o.x * r.y) * d10.z) + (o.x * o.y) * d11.z;
                const float len = norm3(g);
                const f3 nrm = mk3(g.x / len, g.y / len, g.z / len);
                const f3 pos = mk3(extent * (o.x + (float)i) + offset.x, extent * (o.y + (float)j) + offset.y,
                                   extent * (o.z + (float)k) + offset.z);
                float* P = a.positions + 3 * (size_t)(voff + v);
                float* N = a.normals + 3 * (size_t)(voff + v);
                P[0] = pos.x; P[1] = pos.y; P[2] = pos.z;
                N[0] = nrm.x; N[1] = nrm.y; N[2] = nrm.z;
            }
            __syncthreads();
        }

        // ---- quad pass: make_all_surface_nets_quads (surface_nets.rs:251-381) ----
        const int up0 = 17 - (int)s_adj_up[0], up1 = 17 - (int)s_adj_up[1], up2 = 17 - (int)s_adj_up[2];
        uint32_t n_quads = 0;
        for (uint32_t base = 0; base < n_vertices; base += MESH_THREADS) {
            const uint32_t v = base + tid;
            uint32_t qmask = 0;  // bit a: a quad on axis a
            int lin = 0;
            if (v < n_vertices) {
                lin = s_surf[v];
                const int i = lin / 324, j = (lin / 18) % 18, k = lin % 18;
                const bool n1 = s_sd[lin] < 0;
                if (j != 0 && k != 0 && i < up0 && (n1 != (s_sd[lin + 324] < 0))) qmask |= 1u;
                if (i != 0 && k != 0 && j < up1 && (n1 != (s_sd[lin + 18] < 0))) qmask |= 2u;
                if (i != 0 && j != 0 && k < up2 && (n1 != (s_sd[lin + 1] < 0))) qmask |= 4u;
            }
            uint32_t tile_total;
            const uint32_t pre = block_exclusive_scan(__popc(qmask), s_warp, tile_total);
            if (EMIT && !skip_emit && qmask) {
                uint32_t qi = n_quads + pre;
                // d1 negative, d2 positive → negative_face = false (surface_nets.rs:345-349)
                const bool negative_face = !(s_sd[lin] < 0);
#pragma unroll 1
                for (int ax = 0; ax < 3; ++ax) {
                    if (!((qmask >> ax) & 1u)) continue;
                    const int axb = ax == 0 ? 18 : (ax == 1 ? 1 : 324);
                    const int axc = ax == 0 ? 1 : (ax == 1 ? 324 : 18);
                    const int cl[4] = {lin, lin - axb, lin - axc, lin - axb - axc};
                    uint32_t vid[4];
                    f3 p[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        vid[c] = s_l2v[cl[c]];
                        const float* P = a.positions + 3 * (size_t)(voff + vid[c]);
                        p[c] = mk3(P[0], P[1], P[2]);
                    }
                    // triangle corner → quad corner, 3 bits each: [t0c0 t0c1 t0c2 t1c0 t1c1 t1c2]
                    uint32_t order;
                    if (norm3(p[0] - p[3]) < norm3(p[1] - p[2]))
                        order = negative_face ? (0u | 3u << 3 | 1u << 6 | 0u << 9 | 2u << 12 | 3u << 15)
                                              : (0u | 1u << 3 | 3u << 6 | 0u << 9 | 3u << 12 | 2u << 15);
                    else
                        order = negative_face ? (1u | 2u << 3 | 3u << 6 | 1u << 9 | 0u << 12 | 2u << 15)
                                              : (1u | 3u << 3 | 2u << 6 | 1u << 9 | 2u << 12 | 0u << 15);
                    uint32_t* I = a.indices + (size_t)ioff + 6 * (size_t)qi;
                    uint64_t* IM = reinterpret_cast<uint64_t*>(a.index_materials + (size_t)ioff + 6 * (size_t)qi);
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const uint32_t sel = (order >> (3 * c)) & 3u;
                        I[c] = voff + (sel == 0 ? vid[0] : (sel == 1 ? vid[1] : (sel == 2 ? vid[2] : vid[3])));
                    }
                    const uint32_t m0 = s_vmat[vid[0]], m1 = s_vmat[vid[1]], m2 = s_vmat[vid[2]], m3 = s_vmat[vid[3]];
                    if (m0 != 255u && m0 == m1 && m0 == m2 && m0 == m3) {
                        // every vertex has the one material: {index, 0, 0, 0} / {1, 0, 0, 0} (surface_nets.rs:563-577)
                        const uint64_t im = (uint64_t)m0 | (1ull << 32);
#pragma unroll
                        for (int c = 0; c < 6; ++c) IM[c] = im;
                    } else {
                        // several materials meet here: the quad is recorded and finished by k_mesh_materials with one
                        // thread per quad, instead of by the few lanes of this warp that sit on a material boundary
                        uint32_t neg4[4];
                        uint64_t types4[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) vertex_corner_data(cl[c], s_neg, s_type, neg4[c], types4[c]);
                        const uint32_t e = atomicAdd(a.mq_count, 1u);
                        if (e < a.mq_capacity) {
                            uint4* dst = a.mq_entries + 3 * (size_t)e;
                            dst[0] = make_uint4((uint32_t)types4[0], (uint32_t)(types4[0] >> 32), (uint32_t)types4[1], (uint32_t)(types4[1] >> 32));
                            dst[1] = make_uint4((uint32_t)types4[2], (uint32_t)(types4[2] >> 32), (uint32_t)types4[3], (uint32_t)(types4[3] >> 32));
                            dst[2] = make_uint4(ioff + 6u * qi, order, neg4[0] | (neg4[1] << 8) | (neg4[2] << 16) | (neg4[3] << 24), 0u);
                        } else {
                            emit_quad_materials(neg4, types4, order, IM);  // no room: in place
                        }
                    }
                    qi++;
                }
            }
            n_quads += tile_total;
        }

        if (!EMIT) {
            if (tid == 0) {
                const uint32_t ni = 6u * n_quads;
                a.index_count[w] = ni;
                a.vertex_count[w] = ni ? n_vertices : 0u;
                a.has_submesh[w] = ni ? 1u : 0u;
            }
        } else if (tid == 0 && !skip_emit) {
            const uint32_t so = a.submesh_ord[w];
            ivx_chunk_submesh sm;
            sm.chunk_indices[0] = ci + a.first_i;
            sm.chunk_indices[1] = cj;
            sm.chunk_indices[2] = ck;
            sm.index_offset = ioff;
            sm.index_count = a.index_count[w];
            const uint8_t fl = a.chunks[chunk].flags;
            // compute_directional_obscuredness_table (mesh.rs:611-635)
            for (int x = 0; x < 2; ++x)
                for (int y = 0; y < 2; ++y)
                    for (int z = 0; z < 2; ++z) {
                        const bool ox = fl & (1u << (x == 0 ? 0 : 3));
                        const bool oy = fl & (1u << (y == 0 ? 1 : 4));
                        const bool oz = fl & (1u << (z == 0 ? 2 : 5));
                        sm.is_obscured_from_direction[(x * 2 + y) * 2 + z] = (ox && oy && oz) ? 1u : 0u;
                    }
            a.submeshes[so] = sm;
            a.vertex_ranges[2 * so] = voff;
            a.vertex_ranges[2 * so + 1] = voff + a.vertex_count[w];
        }
        __syncthreads();
    }
}

// chunks to mesh = NonUniform with at least one unobscured face (object/sdf.rs:196, object.rs:3016-3019)
__global__ void k_exposed_flags(const DevChunk* __restrict__ chunks, uint32_t n, uint3 nb, uint32_t own_lo,
                                uint32_t own_hi, uint32_t* __restrict__ flag) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const DevChunk ch = chunks[c];
    const uint32_t ci = c / (nb.z * nb.y);
    flag[c] = (ch.kind == 2 && (ch.flags & 0x3F) != 0x3F && ci >= own_lo && ci < own_hi) ? 1u : 0u;
}

cudaError_t launch_exposed_flags(const DevChunk* chunks, uint32_t n, const uint32_t nb[3], uint32_t own_lo,
                                 uint32_t own_hi, uint32_t* flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_exposed_flags<<<(n + 255) / 256, 256, 0, st>>>(chunks, n, make_uint3(nb[0], nb[1], nb[2]), own_lo, own_hi, flag);
    return cudaGetLastError();
}

cudaError_t launch_mesh_materials(const uint4* entries, const uint32_t* count, uint32_t capacity,
                                  ivx_index_materials* index_materials, cudaStream_t st) {
    if (capacity == 0) return cudaSuccess;
    // the count lives on the device: a fixed grid-stride launch (the bulk of the surface has one material)
    k_mesh_materials<<<148 * 4, 128, 0, st>>>(entries, count, capacity, index_materials);
    return cudaGetLastError();
}

cudaError_t launch_mesh(bool emit, const MeshArgs& a, uint32_t grid, cudaStream_t st) {
    if (a.n_work == 0 || grid == 0) return cudaSuccess;
    if (emit)
        k_mesh<true><<<grid, MESH_THREADS, 0, st>>>(a);
    else
        k_mesh<false><<<grid, MESH_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ivx
