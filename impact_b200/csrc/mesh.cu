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

struct VertexMaterials {
    uint8_t indices[8];
    uint8_t weights[8];
};

// SurfaceNetsVertexMaterials::compute + sort_descending (surface_nets.rs:453-538)
__device__ __forceinline__ VertexMaterials vertex_materials(uint32_t neg_mask, const uint8_t* mat) {
    VertexMaterials m;
#pragma unroll
    for (int q = 0; q < 8; ++q) m.indices[q] = m.weights[q] = 0;
    int count = 0;
    for (int c = 0; c < 8; ++c) {
        if ((neg_mask >> c) & 1u) {
            int found = -1;
            for (int q = 0; q < count; ++q)
                if (m.indices[q] == mat[c] && found < 0) found = q;
            if (found < 0) {
                m.indices[count] = mat[c];
                m.weights[count] = 1;
                count++;
            } else {
                m.weights[found] += 1;
            }
        }
    }
    m.indices[7] = (uint8_t)count;
    const int NET[17][2] = {{0, 6}, {1, 5}, {2, 4}, {0, 3}, {1, 2}, {4, 5}, {0, 1}, {2, 3}, {4, 6},
                            {5, 6}, {1, 4}, {3, 5}, {1, 2}, {3, 4}, {5, 6}, {2, 3}, {4, 5}};
#pragma unroll
    for (int s = 0; s < 17; ++s) {
        const int i = NET[s][0], j = NET[s][1];
        if (m.weights[i] < m.weights[j]) {
            uint8_t t = m.indices[i]; m.indices[i] = m.indices[j]; m.indices[j] = t;
            t = m.weights[i]; m.weights[i] = m.weights[j]; m.weights[j] = t;
        }
    }
    return m;
}

// calculate_index_materials_for_triangle (surface_nets.rs:556-637)
__device__ __forceinline__ void triangle_index_materials(const VertexMaterials* const vm[3], ivx_index_materials out[3]) {
    if (vm[0]->indices[7] == 1 && vm[1]->indices[7] == 1 && vm[2]->indices[7] == 1) {
        const uint8_t index = vm[0]->indices[0];
        if (vm[1]->indices[0] == index && vm[2]->indices[0] == index) {
            ivx_index_materials im = {{index, 0, 0, 0}, {1, 0, 0, 0}};
            out[0] = out[1] = out[2] = im;
            return;
        }
    }
    uint8_t top[4] = {0, 0, 0, 0};
    int n_top = 0;
    int off[3] = {0, 0, 0};
    for (int t = 0; t < 4; ++t) {
        uint8_t w[3];
        for (int i = 0; i < 3; ++i) w[i] = vm[i]->weights[off[i]];
        const int mx = (w[0] >= w[1]) ? ((w[0] >= w[2]) ? 0 : 2) : ((w[1] >= w[2]) ? 1 : 2);
        if (w[mx] == 0) break;
        top[t] = vm[mx]->indices[off[mx]];
        n_top++;
        for (int i = 0; i < 3; ++i) {
            for (;;) {
                if (off[i] >= (int)vm[i]->indices[7]) break;
                const uint8_t cand = vm[i]->indices[off[i]];
                bool is_top = false;
                for (int q = 0; q < n_top; ++q) is_top = is_top || (top[q] == cand);
                if (!is_top) break;
                off[i]++;
            }
        }
    }
    for (int v = 0; v < 3; ++v) {
        ivx_index_materials im = {{top[0], top[1], top[2], top[3]}, {0, 0, 0, 0}};
        for (int i = 0; i < n_top; ++i)
            for (int j = 0; j < (int)vm[v]->indices[7]; ++j)
                if (vm[v]->indices[j] == top[i]) {
                    im.weights[i] = vm[v]->weights[j];
                    break;
                }
        out[v] = im;
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

template <bool EMIT>
__global__ void __launch_bounds__(MESH_THREADS) k_mesh(MeshArgs a) {
    __shared__ __align__(16) int8_t s_sd[5832];
    __shared__ __align__(16) uint8_t s_type[5832];
    __shared__ uint16_t s_l2v[EMIT ? 5832 : 1];
    __shared__ uint16_t s_surf[N_CUBES];
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
        for (int cell = tid; cell < 5832; cell += MESH_THREADS) {
            const int bi = cell / 324, bj = (cell / 18) % 18, bk = cell % 18;
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
                        ty = slot[PLANE_TYPE + v];
                    }
                }
            }
            s_sd[cell] = sd;
            s_type[cell] = ty;
            if (EMIT) s_l2v[cell] = 0xFFFF;
        }
        __syncthreads();

        const float extent = a.voxel_extent;
        const float chunk_extent = extent * 16.0f;
        // vertex_position_offset_for_chunk (mesh.rs:559-577); chunk index in the FULL grid
        const f3 offset = mk3((float)(ci + a.first_i) * chunk_extent - 0.5f * extent,
                              (float)cj * chunk_extent - 0.5f * extent, (float)ck * chunk_extent - 0.5f * extent);
        const uint32_t voff = EMIT ? a.vertex_offset[w] : 0u;
        const uint32_t ioff = EMIT ? a.index_offset[w] : 0u;
        const bool skip_emit = EMIT && a.index_count[w] == 0u;  // empty mesh: nothing is appended (mesh.rs:319-321)

        // ---- vertex pass: estimate_surface_nets_surface (surface_nets.rs:152-244) ----
        uint32_t n_vertices = 0;
        for (int base = 0; base < N_CUBES; base += MESH_THREADS) {
            const int q = base + tid;
            uint32_t neg = 0;
            int lin = 0, i = 0, j = 0, k = 0;
            if (q < N_CUBES) {
                i = q / 289;
                j = (q / 17) % 17;
                k = q % 17;
                lin = bidx(i, j, k);
#pragma unroll
                for (int c = 0; c < 8; ++c) neg |= (s_sd[lin + corner_off(c)] < 0 ? 1u : 0u) << c;
            }
            const bool surf = q < N_CUBES && neg != 0u && neg != 0xFFu;
            uint32_t tile_total;
            const uint32_t pre = block_exclusive_scan(surf ? 1u : 0u, s_warp, tile_total);
            if (surf) {
                const uint32_t v = n_vertices + pre;
                s_surf[v] = (uint16_t)lin;
                if (EMIT && !skip_emit) {
                    s_l2v[lin] = (uint16_t)v;
                    float d[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) d[c] = sd_decode((int)s_sd[lin + corner_off(c)]);
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
            }
            n_vertices += tile_total;
        }
        __syncthreads();

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
                const bool n1 = s_sd[lin] < 0;
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    if (!((qmask >> ax) & 1u)) continue;
                    const int axb = ax == 0 ? 18 : (ax == 1 ? 1 : 324);
                    const int axc = ax == 0 ? 1 : (ax == 1 ? 324 : 18);
                    // d1 negative, d2 positive → negative_face = false (surface_nets.rs:345-349)
                    const bool negative_face = !n1;
                    const int cl[4] = {lin, lin - axb, lin - axc, lin - axb - axc};
                    uint32_t vid[4];
                    f3 p[4];
                    VertexMaterials vm[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        vid[c] = s_l2v[cl[c]];
                        const float* P = a.positions + 3 * (size_t)(voff + vid[c]);
                        p[c] = mk3(P[0], P[1], P[2]);
                        uint32_t neg = 0;
                        uint8_t mat[8];
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) {
                            neg |= (s_sd[cl[c] + corner_off(cc)] < 0 ? 1u : 0u) << cc;
                            mat[cc] = s_type[cl[c] + corner_off(cc)];
                        }
                        vm[c] = vertex_materials(neg, mat);
                    }
                    int order[6];
                    if (norm3(p[0] - p[3]) < norm3(p[1] - p[2])) {
                        if (negative_face) { order[0]=0; order[1]=3; order[2]=1; order[3]=0; order[4]=2; order[5]=3; }
                        else               { order[0]=0; order[1]=1; order[2]=3; order[3]=0; order[4]=3; order[5]=2; }
                    } else if (negative_face) { order[0]=1; order[1]=2; order[2]=3; order[3]=1; order[4]=0; order[5]=2; }
                    else                      { order[0]=1; order[1]=3; order[2]=2; order[3]=1; order[4]=2; order[5]=0; }
                    uint32_t* I = a.indices + (size_t)ioff + 6 * (size_t)qi;
                    ivx_index_materials* IM = a.index_materials + (size_t)ioff + 6 * (size_t)qi;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const VertexMaterials* tv[3] = {&vm[order[3 * t]], &vm[order[3 * t + 1]], &vm[order[3 * t + 2]]};
                        ivx_index_materials im[3];
                        triangle_index_materials(tv, im);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            I[3 * t + c] = voff + vid[order[3 * t + c]];
                            IM[3 * t + c] = im[c];
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

cudaError_t launch_mesh(bool emit, const MeshArgs& a, uint32_t grid, cudaStream_t st) {
    if (a.n_work == 0 || grid == 0) return cudaSuccess;
    if (emit)
        k_mesh<true><<<grid, MESH_THREADS, 0, st>>>(a);
    else
        k_mesh<false><<<grid, MESH_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ivx
