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

// position of chunk (i, j, k) in the arrays of a box-local pass, or ~0 outside the box
__device__ __forceinline__ uint32_t box_index(const ChunkBox& box, int i, int j, int k) {
    const uint32_t a = (uint32_t)i - box.c0[0], b = (uint32_t)j - box.c0[1], c = (uint32_t)k - box.c0[2];
    return (a < box.d[0] && b < box.d[1] && c < box.d[2]) ? (a * box.d[1] + b) * box.d[2] + c : 0xFFFFFFFFu;
}

// `box`: convert_flag is indexed by position in that box (chunks outside it are not being converted) instead of by chunk
__device__ __forceinline__ Neighbour neighbour_of(const DevChunk* __restrict__ chunks, uint3 nb, int i, int j, int k,
                                                  int dim, int side, const uint32_t* __restrict__ convert_flag,
                                                  const ChunkBox* box = nullptr) {
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
    bool converting = false;
    if (convert_flag) {
        const uint32_t at = box ? box_index(*box, n[0], n[1], n[2]) : idx;
        converting = at != 0xFFFFFFFFu && convert_flag[at] != 0;
    }
    if (c.kind == 1 || converting) { r.kind = 1; }
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

// BOX: the pass of a modification — the work items are the chunks of `box` and face_mask / convert_flag / slot_of are
// indexed by position in it (written by k_boundary_prep_box), so that nothing of the size of the chunk table is touched.
template <bool BOX>
__global__ void __launch_bounds__(256, 8) k_boundary_apply(DevChunk* __restrict__ chunks, uint32_t n, uint3 nb,
                                                         const uint8_t* __restrict__ face_mask,
                                                         const uint32_t* __restrict__ convert_flag,
                                                         const uint32_t* __restrict__ slot_of,
                                                         unsigned char* __restrict__ voxels,
                                                         const uint32_t* __restrict__ work_list, uint32_t n_work,
                                                         uint32_t work_first, uint32_t own_lo, uint32_t own_hi,
                                                         ChunkBox box) {
    __shared__ __align__(16) uint8_t s_flags[4096];
    __shared__ Neighbour s_nb[6];
    const int tid = threadIdx.x;
    for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        uint32_t c, at;  // the chunk, and its position in the per-pass arrays
        if (BOX) {
            const uint32_t bk = w % box.d[2], bj = (w / box.d[2]) % box.d[1], bi = w / (box.d[2] * box.d[1]);
            c = ((box.c0[0] + bi) * nb.y + box.c0[1] + bj) * nb.z + box.c0[2] + bk;
            at = w;
        } else {
            c = work_list ? work_list[w] : w + work_first;
            at = c;
        }
        DevChunk me = chunks[c];
        const bool converting = convert_flag[at] != 0;
        const uint32_t plane = c / (nb.z * nb.y);
        if ((me.kind != 2 && !converting) || plane < own_lo || plane >= own_hi) continue;  // uniform across the CTA
        const uint8_t mask = face_mask ? face_mask[at] : 0x3F;
        if (BOX && mask == 0) continue;  // (a corner of the box: none of its faces belongs to a refreshed pair)
        const int ck = c % nb.z, cj = (c / nb.z) % nb.y, ci = c / (nb.z * nb.y);
        unsigned char* slot;
        if (converting) {
            // convert_to_non_uniform_if_uniform (object.rs:2530-2550)
            me.kind = 2;
            me.slot = slot_of[at];
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
        if (tid < 6) s_nb[tid] = neighbour_of(chunks, nb, ci, cj, ck, tid >> 1, tid & 1, convert_flag, BOX ? &box : nullptr);
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
    k_boundary_apply<false><<<grid, 256, 0, st>>>(chunks, n, make_uint3(nb[0], nb[1], nb[2]), face_mask, convert_flag,
                                                  slot_of, voxels, work_list, n_work, work_first, own_lo, own_hi, ChunkBox{});
    return cudaGetLastError();
}

// ---- the boundary refresh of a modification, local to the refreshed box -------------------------------------
// One CTA does for the chunks of `box` what k_absorb_face_mask, k_boundary_classify, k_need_slot_for_convert, the prefix
// sum and k_assign_slots do for a whole chunk table: which faces belong to a refreshed pair (`range`: the chunks whose
// upper faces are refreshed, intersection.rs:391-393), which Uniform chunks stop being uniform, and the slots those get —
// `first_slot` + *first_extra + their rank among the converted chunks of the box in chunk order, the numbering the
// whole-table pass gives. Outputs are indexed by position in the box; *total = the slots handed out.
// what the refresh decides for chunk `w` of the box: its face mask, whether it stops being uniform, the slot it has
struct BoxChunk {
    uint32_t c, conv, own_slot;
    uint8_t mask;
};
__device__ __forceinline__ BoxChunk box_chunk(const DevChunk* __restrict__ chunks, uint3 nb, const ChunkBox& box,
                                              const AbsorbRange& range, uint32_t w) {
    BoxChunk r{0, 0, 0, 0};
    const int i = (int)(box.c0[0] + w / (box.d[2] * box.d[1])), j = (int)(box.c0[1] + (w / box.d[2]) % box.d[1]),
              k = (int)(box.c0[2] + w % box.d[2]);
    r.c = ((uint32_t)i * nb.y + (uint32_t)j) * nb.z + (uint32_t)k;
    auto in_range = [&](int x, int y, int z) {
        return (uint32_t)x >= range.c0[0] && (uint32_t)x < range.c1[0] && (uint32_t)y >= range.c0[1] &&
               (uint32_t)y < range.c1[1] && (uint32_t)z >= range.c0[2] && (uint32_t)z < range.c1[2];
    };
    if (in_range(i, j, k)) r.mask |= (1u << 1) | (1u << 3) | (1u << 5);  // upper faces
    if (i > 0 && in_range(i - 1, j, k)) r.mask |= 1u << 0;               // lower faces: the pair belongs to the lower chunk
    if (j > 0 && in_range(i, j - 1, k)) r.mask |= 1u << 2;
    if (k > 0 && in_range(i, j, k - 1)) r.mask |= 1u << 4;
    const DevChunk me = chunks[r.c];
    r.own_slot = me.slot;
    if (me.kind == 1) {
        for (int f = 0; f < 6; ++f) {
            if (!((r.mask >> f) & 1)) continue;
            const Neighbour nbh = neighbour_of(chunks, nb, i, j, k, f >> 1, f & 1, nullptr);
            if (!(nbh.kind == 1 || (nbh.kind == 2 && nbh.face == 1))) r.conv = 1;
        }
    }
    return r;
}

__global__ void __launch_bounds__(1024) k_boundary_prep_box(const DevChunk* __restrict__ chunks, uint3 nb, ChunkBox box,
                                                            AbsorbRange range, uint32_t first_slot,
                                                            const uint32_t* __restrict__ first_extra,
                                                            uint8_t* __restrict__ face_mask, uint32_t* __restrict__ convert_flag,
                                                            uint32_t* __restrict__ slot_of, uint8_t* __restrict__ label_stale,
                                                            uint32_t* __restrict__ total) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_box = box.d[0] * box.d[1] * box.d[2];
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const uint32_t first = first_slot + (first_extra ? *first_extra : 0u);
    for (uint32_t base = 0; base < n_box; base += 1024) {
        const uint32_t w = base + tid;
        BoxChunk bc{0, 0, 0, 0};
        uint32_t need = 0;
        if (w < n_box) {
            bc = box_chunk(chunks, nb, box, range, w);
            need = (bc.conv && bc.own_slot == 0xFFFFFFFFu) ? 1u : 0u;
        }
        // exclusive prefix sum of `need` over the CTA, carried from tile to tile
        const uint32_t bal = __ballot_sync(0xffffffffu, need != 0);
        if (lane == 0) s_warp[warp] = (uint32_t)__popc(bal);
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = s_warp[lane];
            uint32_t incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += t;
            }
            s_warp[lane] = incl - v;
        }
        __syncthreads();
        const uint32_t ord = s_carry + s_warp[warp] + (uint32_t)__popc(bal & ((1u << lane) - 1u));
        if (w < n_box) {
            face_mask[w] = bc.mask;
            convert_flag[w] = bc.conv;
            slot_of[w] = need ? first + ord : bc.own_slot;
            if (bc.conv && label_stale) label_stale[bc.c] = 1;  // a chunk that becomes NonUniform has no region labels yet
        }
        __syncthreads();
        if (tid == 1023) s_carry = ord + need;
        __syncthreads();
    }
    if (tid == 0) *total = s_carry;
}

// the same in three steps for a box too large for one CTA: flags over a grid, the library's prefix sum, slots over a grid
__global__ void k_boundary_box_flags(const DevChunk* __restrict__ chunks, uint3 nb, ChunkBox box, AbsorbRange range,
                                     uint32_t n_box, uint8_t* __restrict__ face_mask, uint32_t* __restrict__ convert_flag,
                                     uint32_t* __restrict__ need, uint32_t* __restrict__ slot_of,
                                     uint8_t* __restrict__ label_stale) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_box) return;
    const BoxChunk bc = box_chunk(chunks, nb, box, range, w);
    face_mask[w] = bc.mask;
    convert_flag[w] = bc.conv;
    need[w] = (bc.conv && bc.own_slot == 0xFFFFFFFFu) ? 1u : 0u;
    slot_of[w] = bc.own_slot;
    if (bc.conv && label_stale) label_stale[bc.c] = 1;
}
__global__ void k_boundary_box_slots(const uint32_t* __restrict__ need, const uint32_t* __restrict__ ord, uint32_t n_box,
                                     uint32_t first_slot, const uint32_t* __restrict__ first_extra,
                                     uint32_t* __restrict__ slot_of) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_box || !need[w]) return;
    slot_of[w] = first_slot + (first_extra ? *first_extra : 0u) + ord[w];
}

cudaError_t launch_boundary_refresh_box(DevChunk* chunks, uint32_t n, const uint32_t nb[3], const ChunkBox& box,
                                        const AbsorbRange& range, uint32_t first_slot, const uint32_t* first_extra,
                                        uint8_t* face_mask, uint32_t* convert_flag, uint32_t* slot_of, uint8_t* label_stale,
                                        uint32_t* total, uint32_t* need, uint32_t* ord, bool prep, bool apply,
                                        unsigned char* voxels, uint32_t grid, cudaStream_t st) {
    const uint32_t n_box = box.d[0] * box.d[1] * box.d[2];
    if (n_box == 0) return cudaSuccess;
    const uint3 nb3 = make_uint3(nb[0], nb[1], nb[2]);
    if (prep && n_box <= BOUNDARY_BOX_ONE_CTA) {
        k_boundary_prep_box<<<1, 1024, 0, st>>>(chunks, nb3, box, range, first_slot, first_extra, face_mask, convert_flag,
                                                slot_of, label_stale, total);
    } else if (prep) {
        if (!need || !ord) return cudaErrorInvalidValue;
        k_boundary_box_flags<<<(n_box + 255) / 256, 256, 0, st>>>(chunks, nb3, box, range, n_box, face_mask, convert_flag, need,
                                                                  slot_of, label_stale);
        if (cudaError_t e = launch_exclusive_scan(need, ord, n_box, total, st)) return e;
        k_boundary_box_slots<<<(n_box + 255) / 256, 256, 0, st>>>(need, ord, n_box, first_slot, first_extra, slot_of);
    }
    if (apply)
        k_boundary_apply<true><<<std::max(1u, std::min(grid, n_box)), 256, 0, st>>>(chunks, n, nb3, face_mask, convert_flag,
                                                                                   slot_of, voxels, nullptr, n_box, 0, 0,
                                                                                   nb[0], box);
    return cudaGetLastError();
}

}  // namespace ivx
