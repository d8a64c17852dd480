// Cross-chunk derived state (sm_100a): uniform → non-uniform conversion,
// outward adjacency bits on chunk faces and face obscuredness.
//
// Replaces VoxelObject::update_all_chunk_boundary_adjacencies →
// VoxelChunk::update_mutual_face_adjacencies
// (engine/crates/impact_voxel/src/object.rs:1659-1785, 2077-2652).
//
// The reference visits chunk pairs sequentially and mutates both sides. Its end
// state is a pure function of the chunk kinds and face distributions, which is
// what these kernels evaluate, one CTA per chunk, all chunks in parallel:
//   * a Uniform chunk stays uniform iff each of its 6 neighbours is Uniform or
//     NonUniform with a Full facing face (object.rs:2118-2160, 2262-2330);
//     otherwise it is converted: 4096 copies of its voxel with full adjacency,
//     all faces Full, all faces obscured (object.rs:2530-2550), then treated
//     like any other non-uniform chunk;
//   * for a NonUniform chunk face whose own distribution is not Empty, the
//     outward adjacency bit of the 256 face voxels is cleared if the
//     neighbour is Void / has an Empty facing face, set if the neighbour is
//     Uniform / has a Full facing face, and for a Mixed facing face set or
//     cleared per non-empty voxel from the adjacent voxel (object.rs:2552-2660);
//   * a face is obscured iff the neighbour is Uniform or its facing face is Full.
// Faces not selected by `face_mask` are left untouched, which is how the
// modification path restricts the update to the touched chunk range
// (object/intersection.rs:391-393).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace ivx {

struct Neighbour {
    uint8_t kind;      // 0 void / outside, 1 uniform, 2 non-uniform
    uint8_t face;      // facing face distribution when non-uniform
    uint32_t slot;
};

__device__ __forceinline__ Neighbour neighbour_of(const DevChunk* __restrict__ chunks, uint3 nb, int i, int j, int k,
                                                  int dim, int side, const uint32_t* __restrict__ convert_flag) {
    int n[3] = {i, j, k};
    n[dim] += side ? 1 : -1;
    Neighbour r{0, 0, 0};
    if (n[0] < 0 || n[1] < 0 || n[2] < 0 || n[0] >= (int)nb.x || n[1] >= (int)nb.y || n[2] >= (int)nb.z) return r;
    const uint32_t idx = (n[0] * nb.y + n[1]) * nb.z + n[2];
    const DevChunk c = chunks[idx];
    r.kind = c.kind;
    r.slot = c.slot;
    if (c.kind == 2) r.face = c.face[dim * 2 + (1 - side)];
    // a uniform neighbour that is being converted in this pass presents Full faces either way
    if (c.kind == 1 || (convert_flag && convert_flag[idx])) { r.kind = 1; }
    return r;
}

// face bit: bit (dim, side) of a 6-bit mask, index dim*2+side
__global__ void k_boundary_classify(const DevChunk* __restrict__ chunks, uint32_t n, uint3 nb,
                                    const uint8_t* __restrict__ face_mask, uint32_t* __restrict__ convert_flag,
                                    uint32_t own_lo, uint32_t own_hi) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    uint32_t conv = 0;
    const DevChunk me = chunks[c];
    const uint32_t plane = c / (nb.z * nb.y);
    if (me.kind == 1 && plane >= own_lo && plane < own_hi) {
        const int k = c % nb.z, j = (c / nb.z) % nb.y, i = c / (nb.z * nb.y);
        const uint8_t mask = face_mask ? face_mask[c] : 0x3F;
        for (int f = 0; f < 6; ++f) {
            if (!((mask >> f) & 1)) continue;
            Neighbour nbh = neighbour_of(chunks, nb, i, j, k, f >> 1, f & 1, nullptr);
            const bool full = nbh.kind == 1 || (nbh.kind == 2 && nbh.face == 1);
            if (!full) conv = 1;
        }
    }
    convert_flag[c] = conv;
}

__global__ void __launch_bounds__(256, 8) k_boundary_apply(DevChunk* __restrict__ chunks, uint32_t n, uint3 nb,
                                                         const uint8_t* __restrict__ face_mask,
                                                         const uint32_t* __restrict__ convert_flag,
                                                         const uint32_t* __restrict__ slot_of,
                                                         unsigned char* __restrict__ voxels,
                                                         const uint32_t* __restrict__ work_list, uint32_t n_work,
                                                         uint32_t work_first, uint32_t own_lo, uint32_t own_hi) {
    __shared__ __align__(16) uint8_t s_flags[4096];
    __shared__ Neighbour s_nb[6];
    const int tid = threadIdx.x;
    for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const uint32_t c = work_list ? work_list[w] : w + work_first;
        DevChunk me = chunks[c];
        const bool converting = convert_flag[c] != 0;
        const uint32_t plane = c / (nb.z * nb.y);
        if ((me.kind != 2 && !converting) || plane < own_lo || plane >= own_hi) continue;  // uniform across the CTA
        const uint8_t mask = face_mask ? face_mask[c] : 0x3F;
        const int ck = c % nb.z, cj = (c / nb.z) % nb.y, ci = c / (nb.z * nb.y);
        unsigned char* slot;
        if (converting) {
            // convert_to_non_uniform_if_uniform (object.rs:2530-2550)
            me.kind = 2;
            me.slot = slot_of[c];
            for (int q = 0; q < 6; ++q) me.face[q] = 1;
            me.flags = 0x3F;
            slot = voxels + (size_t)me.slot * SLOT_BYTES;
            const uint32_t sdw = 0x80808080u;
            const uint32_t tw = (uint32_t)me.u_type * 0x01010101u;
            *reinterpret_cast<uint4*>(slot + PLANE_SD + tid * 16) = make_uint4(sdw, sdw, sdw, sdw);
            *reinterpret_cast<uint4*>(slot + PLANE_TYPE + tid * 16) = make_uint4(tw, tw, tw, tw);
            *reinterpret_cast<uint4*>(&s_flags[tid * 16]) = make_uint4(0xFCFCFCFCu, 0xFCFCFCFCu, 0xFCFCFCFCu, 0xFCFCFCFCu);
        } else {
            slot = voxels + (size_t)me.slot * SLOT_BYTES;
            *reinterpret_cast<uint4*>(&s_flags[tid * 16]) = *reinterpret_cast<const uint4*>(slot + PLANE_FLAGS + tid * 16);
        }
        // the six neighbours' descriptors in one round trip
        if (tid < 6) s_nb[tid] = neighbour_of(chunks, nb, ci, cj, ck, tid >> 1, tid & 1, convert_flag);
        __syncthreads();
        uint8_t cflags = me.flags;
        // this thread's face voxel per face, and for faces against a Mixed neighbour face the adjacent voxel's
        // emptiness, all fetched before any of them is used (independent loads instead of six dependent ones)
        int fidx[6];
        uint32_t adj_empty = 0, mixed = 0;
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            const int dim = f >> 1, side = f & 1;
            const int b = tid >> 4, cc = tid & 15, p = side ? 15 : 0;
            int v[3];
            v[dim] = p;
            v[dim == 0 ? 1 : 0] = b;
            v[dim == 2 ? 1 : 2] = cc;
            fidx[f] = vidx(v[0], v[1], v[2]);
            const Neighbour nbh = s_nb[f];
            if (((mask >> f) & 1) && me.face[f] != 0 && nbh.kind == 2 && nbh.face == 2 && !(s_flags[fidx[f]] & 1)) {
                int a3[3] = {v[0], v[1], v[2]};
                a3[dim] = side ? 0 : 15;
                const unsigned char* nslot = voxels + (size_t)nbh.slot * SLOT_BYTES;
                mixed |= 1u << f;
                if (nslot[PLANE_FLAGS + vidx(a3[0], a3[1], a3[2])] & 1) adj_empty |= 1u << f;
            }
        }
        __syncthreads();  // every IS_EMPTY read above precedes the first byte update below
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            if (!((mask >> f) & 1)) continue;
            const int dim = f >> 1, side = f & 1;
            const Neighbour nbh = s_nb[f];
            const uint8_t own = me.face[f];
            // obscuredness
            const uint8_t obit = (uint8_t)(1u << (side == 0 ? dim : 3 + dim));
            const bool obscured = nbh.kind == 1 || (nbh.kind == 2 && nbh.face == 1);
            if (obscured) cflags |= obit; else cflags &= (uint8_t)~obit;
            if (own != 0) {
                const uint8_t abit = (uint8_t)(1u << ((side == 0 ? 2 : 5) + dim));
                const int idx = fidx[f];
                if (nbh.kind == 0 || (nbh.kind == 2 && nbh.face == 0)) {
                    s_flags[idx] &= (uint8_t)~abit;
                } else if (nbh.kind == 1 || nbh.face == 1) {
                    s_flags[idx] |= abit;
                } else if ((mixed >> f) & 1u) {
                    if ((adj_empty >> f) & 1u) s_flags[idx] &= (uint8_t)~abit; else s_flags[idx] |= abit;
                }
            }
            __syncthreads();  // edge / corner voxels sit on several faces
        }
        *reinterpret_cast<uint4*>(slot + PLANE_FLAGS + tid * 16) = *reinterpret_cast<const uint4*>(&s_flags[tid * 16]);
        if (tid == 0) {
            me.flags = (uint8_t)((me.flags & 0xC0) | (cflags & 0x3F));
            chunks[c] = me;
        }
        __syncthreads();
    }
}

cudaError_t launch_boundary_classify(const DevChunk* chunks, uint32_t n, const uint32_t nb[3], const uint8_t* face_mask,
                                     uint32_t* convert_flag, uint32_t own_lo, uint32_t own_hi, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_boundary_classify<<<(n + 255) / 256, 256, 0, st>>>(chunks, n, make_uint3(nb[0], nb[1], nb[2]), face_mask,
                                                         convert_flag, own_lo, own_hi);
    return cudaGetLastError();
}
cudaError_t launch_boundary_apply(DevChunk* chunks, uint32_t n, const uint32_t nb[3], const uint8_t* face_mask,
                                  const uint32_t* convert_flag, const uint32_t* slot_of, unsigned char* voxels,
                                  const uint32_t* work_list, uint32_t n_work, uint32_t own_lo, uint32_t own_hi,
                                  uint32_t grid, cudaStream_t st) {
    if (n_work == 0) return cudaSuccess;
    uint32_t work_first = 0;
    if (!work_list) {
        // without a work list only the chunks of planes [own_lo, own_hi) can have work: walk just those
        const uint32_t plane = nb[1] * nb[2];
        const uint32_t lo = std::min(own_lo, nb[0]), hi = std::min(own_hi, nb[0]);
        if (hi <= lo) return cudaSuccess;
        work_first = lo * plane;
        n_work = std::min(n_work, (hi - lo) * plane);
        grid = std::max(1u, std::min(grid, n_work));
    }
    k_boundary_apply<<<grid, 256, 0, st>>>(chunks, n, make_uint3(nb[0], nb[1], nb[2]), face_mask, convert_flag, slot_of,
                                           voxels, work_list, n_work, work_first, own_lo, own_hi);
    return cudaGetLastError();
}

}  // namespace ivx
