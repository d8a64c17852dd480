// Slab-boundary exchange support (multi-GPU, sm_100a).
//
// The object is partitioned into x-slabs of chunk planes, one per rank (the reference's own thread
// split of the x-major linear chunk index, engine/crates/impact_voxel/src/object.rs:423-427).
// Generation needs no exchange. Derived state and meshing of a rank's outer chunk planes read the
// neighbouring rank's boundary plane: the 1-voxel brick padding (object/sdf.rs:35, 410-428), the face
// adjacency / obscuredness rules (object.rs:1682-1704) and, for the "+x neighbour is non-uniform"
// quad-ownership rule (object/sdf/surface_nets.rs:252-261), that plane's final chunk kinds.
// These kernels pack a boundary plane into one contiguous fixed-size buffer (chunk descriptors, then one
// 768-byte voxel layer per chunk) for NCCL send/recv, and unpack it into the halo plane on the receiving side.
#include "common.cuh"
#include "kernels.h"

namespace ivx {

// Per non-uniform chunk only the voxel layer that touches the neighbouring slab travels: 256 voxels x 3 planes =
// 768 bytes at a fixed position (chunk t of the plane at head + t * 768), so a message has a fixed size known to
// both sides and needs no size handshake.
constexpr uint32_t HALO_LAYER_BYTES = 3u * 256u;

// CTA per chunk of the plane: descriptor + the layer `layer_i` (i = 0 or 15) of its three planes
__global__ void __launch_bounds__(64) k_halo_pack(const DevChunk* __restrict__ chunks, uint32_t plane_first,
                                                  uint32_t plane_chunks, uint32_t layer_i,
                                                  const unsigned char* __restrict__ voxels, unsigned char* __restrict__ dst) {
    const uint32_t t = blockIdx.x;
    if (t >= plane_chunks) return;
    DevChunk c = chunks[plane_first + t];
    DevChunk* out_desc = reinterpret_cast<DevChunk*>(dst);
    unsigned char* payload = dst + (size_t)plane_chunks * sizeof(DevChunk) + (size_t)t * HALO_LAYER_BYTES;
    const uint32_t q = threadIdx.x;  // 48 x 16 bytes
    if (q < 48u) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (c.kind == 2)
            v = *reinterpret_cast<const uint4*>(voxels + (size_t)c.slot * SLOT_BYTES + (q >> 4) * 4096u + layer_i * 256u + (q & 15u) * 16u);
        *reinterpret_cast<uint4*>(payload + q * 16u) = v;
    }
    if (q == 0) {
        c.slot = c.kind == 2 ? t : 0xFFFFFFFFu;
        out_desc[t] = c;
    }
}

__global__ void __launch_bounds__(64) k_halo_unpack(DevChunk* __restrict__ chunks, uint32_t plane_first,
                                                    uint32_t plane_chunks, uint32_t first_slot, uint32_t layer_i,
                                                    unsigned char* __restrict__ voxels, const unsigned char* __restrict__ src) {
    const uint32_t t = blockIdx.x;
    if (t >= plane_chunks) return;
    const DevChunk* in_desc = reinterpret_cast<const DevChunk*>(src);
    const unsigned char* payload = src + (size_t)plane_chunks * sizeof(DevChunk) + (size_t)t * HALO_LAYER_BYTES;
    DevChunk c = in_desc[t];
    const uint32_t q = threadIdx.x;
    if (c.kind == 2) {
        const uint32_t slot = first_slot + t;
        if (q < 48u)
            *reinterpret_cast<uint4*>(voxels + (size_t)slot * SLOT_BYTES + (q >> 4) * 4096u + layer_i * 256u + (q & 15u) * 16u) =
                *reinterpret_cast<const uint4*>(payload + q * 16u);
        c.slot = slot;
    }
    if (q == 0) {
        c.pre = PRE_ACTIVE;
        chunks[plane_first + t] = c;
    }
}

// second exchange: whether each chunk of the boundary plane ends up non-uniform (after conversion)
__global__ void k_halo_kinds_pack(const DevChunk* __restrict__ chunks, const uint32_t* __restrict__ convert_flag,
                                  uint32_t plane_first, uint32_t plane_chunks, uint8_t* __restrict__ dst) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= plane_chunks) return;
    const DevChunk c = chunks[plane_first + t];
    dst[t] = (c.kind == 2 || (c.kind == 1 && convert_flag[plane_first + t])) ? 1 : 0;
}
__global__ void k_halo_kinds_unpack(DevChunk* __restrict__ chunks, uint32_t plane_first, uint32_t plane_chunks,
                                    const uint8_t* __restrict__ src) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= plane_chunks) return;
    if (src[t] && chunks[plane_first + t].kind == 1) chunks[plane_first + t].pre = PRE_CONVERTED_HALO;
}

size_t halo_message_bytes(uint32_t plane_chunks) { return (size_t)plane_chunks * (sizeof(DevChunk) + HALO_LAYER_BYTES); }

cudaError_t launch_halo_pack(const DevChunk* chunks, uint32_t plane_first, uint32_t plane_chunks, uint32_t layer_i,
                             const unsigned char* voxels, unsigned char* dst, cudaStream_t st) {
    if (plane_chunks == 0) return cudaSuccess;
    k_halo_pack<<<plane_chunks, 64, 0, st>>>(chunks, plane_first, plane_chunks, layer_i, voxels, dst);
    return cudaGetLastError();
}
cudaError_t launch_halo_unpack(DevChunk* chunks, uint32_t plane_first, uint32_t plane_chunks, uint32_t first_slot,
                               uint32_t layer_i, unsigned char* voxels, const unsigned char* src, cudaStream_t st) {
    if (plane_chunks == 0) return cudaSuccess;
    k_halo_unpack<<<plane_chunks, 64, 0, st>>>(chunks, plane_first, plane_chunks, first_slot, layer_i, voxels, src);
    return cudaGetLastError();
}
cudaError_t launch_halo_kinds_pack(const DevChunk* chunks, const uint32_t* convert_flag, uint32_t plane_first,
                                   uint32_t plane_chunks, uint8_t* dst, cudaStream_t st) {
    if (plane_chunks == 0) return cudaSuccess;
    k_halo_kinds_pack<<<(plane_chunks + 255) / 256, 256, 0, st>>>(chunks, convert_flag, plane_first, plane_chunks, dst);
    return cudaGetLastError();
}
cudaError_t launch_halo_kinds_unpack(DevChunk* chunks, uint32_t plane_first, uint32_t plane_chunks, const uint8_t* src,
                                     cudaStream_t st) {
    if (plane_chunks == 0) return cudaSuccess;
    k_halo_kinds_unpack<<<(plane_chunks + 255) / 256, 256, 0, st>>>(chunks, plane_first, plane_chunks, src);
    return cudaGetLastError();
}


// ---- mesh gather over peer memory ----------------------------------------------------------------------
// One slab's mesh arrays are written straight into the merged mesh that lives in ANOTHER GPU's memory (mapped through
// CUDA IPC, reached over NVLink / NVSwitch), already rebased: word i of `src` lands at dst[i], plus `add` when
// i % period == col (period 1: every word — the u32 vertex indices; period 13, col 3: ChunkSubmesh::index_offset;
// period 0: plain copy). 16-byte accesses on both sides; `n_words` need not be a multiple of 4.
__global__ void __launch_bounds__(256) k_push_words(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n_words,
                                                     uint32_t add, uint32_t period, uint32_t col) {
    const size_t n4 = n_words / 4;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto fix = [&](uint32_t w, size_t i) { return (period == 1u || (period > 1u && i % period == col)) ? w + add : w; };
    if (aligned) {
        for (size_t q = t; q < n4; q += stride) {
            uint4 v = reinterpret_cast<const uint4*>(src)[q];
            v.x = fix(v.x, 4 * q);
            v.y = fix(v.y, 4 * q + 1);
            v.z = fix(v.z, 4 * q + 2);
            v.w = fix(v.w, 4 * q + 3);
            reinterpret_cast<uint4*>(dst)[q] = v;
        }
        for (size_t i = 4 * n4 + t; i < n_words; i += stride) dst[i] = fix(src[i], i);
    } else {
        for (size_t i = t; i < n_words; i += stride) dst[i] = fix(src[i], i);
    }
}
cudaError_t launch_push_words(const void* src, void* dst, size_t n_words, uint32_t add, uint32_t period, uint32_t col,
                              uint32_t grid, cudaStream_t st) {
    if (n_words == 0) return cudaSuccess;
    k_push_words<<<grid, 256, 0, st>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), n_words, add, period, col);
    return cudaGetLastError();
}

}  // namespace ivx
