// Disconnected-region extraction kernels (sm_100a).
//
// Replace the data movement of
//   VoxelObject::extract_disconnected_region                     (object/extraction.rs:297-600)
//   create_extracted_voxel_object_in_single_chunk_if_possible    (object/extraction.rs:1976-2141)
// The region to extract, its chunk list and which local regions of every chunk belong to it come from the
// connected-region resolution (split.cu + the host union-find); these kernels move the chunks / voxels between
// the two objects and refresh the in-chunk state of every chunk they rewrite.
//
//   k_extract_chunks   one CTA per chunk of the extracted object's chunk grid (the bounding box of the region's
//                      chunks): padding → Void; Uniform → moved; NonUniform holding only the region → all three
//                      planes moved, source chunk becomes Void; NonUniform shared with other regions ("mixed") →
//                      per voxel: empty voxels are copied, the region's voxels move (the source gets
//                      Voxel::maximally_outside), other regions' voxels read as maximally_outside in the copy; both
//                      chunks then get update_all_internal_state_and_determine_sparseness (object.rs:2761-2874).
//   k_repack_single    an extracted object of at most 2 x 2 x 2 chunks whose non-empty voxels span <= 14 per axis
//                      is re-packed into one chunk with an empty boundary layer (extraction.rs:2003-2120).
#include "common.cuh"
#include "kernels.h"

namespace ivx {

// The in-chunk flag refresh of update_all_internal_state_and_determine_sparseness / update_internal_adjacencies
// (object.rs:2673-2874) as a pure function of (previous flags, emptiness), for the 16 voxels of thread (ti, tj):
// adjacency bits of non-empty voxels follow their in-chunk neighbours; an empty voxel keeps its (possibly stale)
// upper bits, its lower bits are cleared where the lower neighbour is empty; bits facing the chunk boundary are
// left alone. Returns the mask of empty voxels of the column.
__device__ __forceinline__ uint32_t refresh_column_flags(const uint8_t* s_fl, int ti, int tj, uint8_t nf[16]) {
    uint32_t empty_mask = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int idx = vidx(ti, tj, k);
        uint8_t f = s_fl[idx];
        const bool e = (f & 1) != 0;
        if (e) empty_mask |= 1u << k;
        const int up[3] = {ti < 15 ? vidx(ti + 1, tj, k) : -1, tj < 15 ? vidx(ti, tj + 1, k) : -1, k < 15 ? vidx(ti, tj, k + 1) : -1};
        const int dn[3] = {ti > 0 ? vidx(ti - 1, tj, k) : -1, tj > 0 ? vidx(ti, tj - 1, k) : -1, k > 0 ? vidx(ti, tj, k - 1) : -1};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const uint8_t ub = (uint8_t)(1u << (5 + d)), db = (uint8_t)(1u << (2 + d));
            if (up[d] >= 0 && !e) {
                if (s_fl[up[d]] & 1) f &= (uint8_t)~ub; else f |= ub;
            }
            if (dn[d] >= 0) {
                if (s_fl[dn[d]] & 1) f &= (uint8_t)~db;
                else if (!e) f |= db;
            }
        }
        nf[k] = f;
    }
    return empty_mask;
}

// face empty counts → FaceVoxelDistribution (object.rs:2967-2981); s_cnt[6] zeroed by the caller, block-wide
__device__ __forceinline__ void count_face_empties(uint32_t empty_mask, int ti, int tj, uint32_t* s_cnt) {
    const uint32_t ne = __popc(empty_mask);
    if (ti == 0) atomicAdd(&s_cnt[0], ne);
    if (ti == 15) atomicAdd(&s_cnt[1], ne);
    if (tj == 0) atomicAdd(&s_cnt[2], ne);
    if (tj == 15) atomicAdd(&s_cnt[3], ne);
    if (empty_mask & 1u) atomicAdd(&s_cnt[4], 1u);
    if (empty_mask & 0x8000u) atomicAdd(&s_cnt[5], 1u);
}

__device__ __forceinline__ uint4 pack16(const uint8_t b[16]) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k >> 2] |= (uint32_t)b[k] << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void unpack16(uint4 v, uint8_t b[16]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) b[k] = (uint8_t)((w[k >> 2] >> (8 * (k & 3))) & 0xFFu);
}

__global__ void __launch_bounds__(256) k_extract_chunks(ExtractArgs a) {
    __shared__ __align__(16) uint8_t s_fl_dst[4096];
    __shared__ __align__(16) uint8_t s_fl_src[4096];
    __shared__ uint32_t s_cnt[12];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    for (uint32_t e = blockIdx.x; e < a.n_ext; e += gridDim.x) {
        const uint32_t mode = a.mode[e];
        DevChunk dst{};
        dst.slot = 0xFFFFFFFFu;
        if (mode == 0u) {  // padding of the extracted object's chunk grid
            if (tid == 0) a.dst_chunks[e] = dst;
            continue;
        }
        const uint32_t c = a.src_index[e];
        DevChunk src = a.src_chunks[c];
        DevChunk gone{};
        gone.slot = 0xFFFFFFFFu;
        if (mode == 1u) {  // Uniform: moved as it is (extraction.rs:512-531)
            if (tid == 0) {
                dst = src;
                dst.slot = 0xFFFFFFFFu;
                dst.flags = 0;
                dst.pre = PRE_UNIFORM;
                a.dst_chunks[e] = dst;
                a.src_chunks[c] = gone;
            }
            continue;
        }
        const unsigned char* sslot = a.src_voxels + (size_t)src.slot * SLOT_BYTES;
        unsigned char* dslot = a.dst_voxels + (size_t)a.dst_slot[e] * SLOT_BYTES;
        dst.kind = 2;
        dst.pre = PRE_ACTIVE;
        dst.slot = a.dst_slot[e];
        const uint4 wsd = *reinterpret_cast<const uint4*>(sslot + PLANE_SD + tid * 16);
        const uint4 wty = *reinterpret_cast<const uint4*>(sslot + PLANE_TYPE + tid * 16);
        const uint4 wfl = *reinterpret_cast<const uint4*>(sslot + PLANE_FLAGS + tid * 16);
        if (mode == 2u) {  // only the extracted region lives here: the chunk changes owner (extraction.rs:447-477)
            *reinterpret_cast<uint4*>(dslot + PLANE_SD + tid * 16) = wsd;
            *reinterpret_cast<uint4*>(dslot + PLANE_TYPE + tid * 16) = wty;
            *reinterpret_cast<uint4*>(dslot + PLANE_FLAGS + tid * 16) = wfl;
            uint8_t fl[16];
            unpack16(wfl, fl);
            uint32_t non_empty = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) non_empty += (fl[k] & 1) ? 0u : 1u;
            if (non_empty) atomicAdd(a.non_empty_count, non_empty);
            if (tid == 0) {
                for (int q = 0; q < 6; ++q) dst.face[q] = src.face[q];
                dst.flags = 0;
                a.dst_chunks[e] = dst;
                a.src_chunks[c] = gone;
                a.src_dirty[c] = 1;
            }
            continue;
        }
        // mixed chunk (extraction.rs:389-446)
        const uint4 wlb = *reinterpret_cast<const uint4*>(a.src_labels + (size_t)src.slot * 4096 + tid * 16);
        uint8_t sd[16], ty[16], fl[16], lb[16], dsd[16], dty[16], dfl[16];
        unpack16(wsd, sd);
        unpack16(wty, ty);
        unpack16(wfl, fl);
        unpack16(wlb, lb);
        const uint8_t* is_r = a.region_is_r + a.first_region[e];
        uint32_t non_empty = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (fl[k] & 1) {  // empty voxels shape the mesh next to the surface: copied unconditionally
                dsd[k] = sd[k]; dty[k] = ty[k]; dfl[k] = fl[k];
            } else if (is_r[lb[k]]) {
                dsd[k] = sd[k]; dty[k] = ty[k]; dfl[k] = fl[k];
                sd[k] = 127; ty[k] = 255; fl[k] = 1;  // Voxel::maximally_outside
                non_empty++;
            } else {
                dsd[k] = 127; dty[k] = 255; dfl[k] = 1;
            }
        }
        if (non_empty) atomicAdd(a.non_empty_count, non_empty);
        *reinterpret_cast<uint4*>(&s_fl_dst[tid * 16]) = pack16(dfl);
        *reinterpret_cast<uint4*>(&s_fl_src[tid * 16]) = pack16(fl);
        if (tid < 12) s_cnt[tid] = 0;
        __syncthreads();
        uint8_t nfd[16], nfs[16];
        const uint32_t em_d = refresh_column_flags(s_fl_dst, ti, tj, nfd);
        const uint32_t em_s = refresh_column_flags(s_fl_src, ti, tj, nfs);
        count_face_empties(em_d, ti, tj, s_cnt);
        count_face_empties(em_s, ti, tj, s_cnt + 6);
        const int only_empty_d = __syncthreads_and(em_d == 0xFFFFu);
        const int only_empty_s = __syncthreads_and(em_s == 0xFFFFu);
        *reinterpret_cast<uint4*>(dslot + PLANE_SD + tid * 16) = pack16(dsd);
        *reinterpret_cast<uint4*>(dslot + PLANE_TYPE + tid * 16) = pack16(dty);
        *reinterpret_cast<uint4*>(dslot + PLANE_FLAGS + tid * 16) = pack16(nfd);
        unsigned char* wslot = a.src_voxels + (size_t)src.slot * SLOT_BYTES;
        *reinterpret_cast<uint4*>(wslot + PLANE_SD + tid * 16) = pack16(sd);
        *reinterpret_cast<uint4*>(wslot + PLANE_TYPE + tid * 16) = pack16(ty);
        *reinterpret_cast<uint4*>(wslot + PLANE_FLAGS + tid * 16) = pack16(nfs);
        if (tid == 0) {
            for (int q = 0; q < 6; ++q) {
                dst.face[q] = s_cnt[q] == 256u ? 0 : (s_cnt[q] == 0u ? 1 : 2);
                src.face[q] = s_cnt[6 + q] == 256u ? 0 : (s_cnt[6 + q] == 0u ? 1 : 2);
            }
            dst.flags = only_empty_d ? (uint8_t)(1u << 6) : 0;
            if (only_empty_s) src.flags |= (uint8_t)(1u << 6); else src.flags &= (uint8_t)~(1u << 6);
            a.dst_chunks[e] = dst;
            a.src_chunks[c] = src;
            a.src_dirty[c] = 1;
            if (a.src_label_stale) a.src_label_stale[c] = 1;
        }
        __syncthreads();
    }
}

// One chunk out of an object of at most 2 x 2 x 2 chunks: destination voxel p takes the voxel at org + p of the small
// object (maximally_outside beyond its NonUniform chunks), then update_internal_adjacencies; the face distributions
// follow from where the occupied range touches the chunk (extraction.rs:2085-2103).
__global__ void __launch_bounds__(256) k_repack_single(const DevChunk* __restrict__ chunks, uint3 nb,
                                                        const unsigned char* __restrict__ voxels, uint3 org, uint3 occ_lo,
                                                        uint3 occ_hi, DevChunk* __restrict__ out_chunk,
                                                        unsigned char* __restrict__ out_slot) {
    __shared__ __align__(16) uint8_t s_fl[4096];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    uint8_t sd[16], ty[16], fl[16];
    const uint32_t si = org.x + ti, sj = org.y + tj;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t sk = org.z + k;
        sd[k] = 127; ty[k] = 255; fl[k] = 1;
        if (si < nb.x * 16u && sj < nb.y * 16u && sk < nb.z * 16u) {
            const DevChunk c = chunks[((si >> 4) * nb.y + (sj >> 4)) * nb.z + (sk >> 4)];
            if (c.kind == 2) {
                const unsigned char* slot = voxels + (size_t)c.slot * SLOT_BYTES;
                const int v = vidx(si & 15, sj & 15, sk & 15);
                sd[k] = slot[PLANE_SD + v];
                ty[k] = slot[PLANE_TYPE + v];
                fl[k] = slot[PLANE_FLAGS + v];
            }
        }
    }
    *reinterpret_cast<uint4*>(&s_fl[tid * 16]) = pack16(fl);
    __syncthreads();
    uint8_t nf[16];
    refresh_column_flags(s_fl, ti, tj, nf);
    *reinterpret_cast<uint4*>(out_slot + PLANE_SD + tid * 16) = pack16(sd);
    *reinterpret_cast<uint4*>(out_slot + PLANE_TYPE + tid * 16) = pack16(ty);
    *reinterpret_cast<uint4*>(out_slot + PLANE_FLAGS + tid * 16) = pack16(nf);
    if (tid == 0) {
        DevChunk d{};
        d.kind = 2;
        d.pre = PRE_ACTIVE;
        d.slot = 0;
        d.flags = 0;
        const uint32_t o3[3] = {org.x, org.y, org.z}, lo3[3] = {occ_lo.x, occ_lo.y, occ_lo.z}, hi3[3] = {occ_hi.x, occ_hi.y, occ_hi.z};
        for (int dim = 0; dim < 3; ++dim) {
            d.face[2 * dim] = o3[dim] == lo3[dim] ? 2 : 0;
            d.face[2 * dim + 1] = o3[dim] + 16u == hi3[dim] ? 2 : 0;
        }
        *out_chunk = d;
    }
}

// ---- ingest of host-generated chunks -------------------------------------------------------------------------
// VoxelChunk::create_for_generated_voxels (object.rs:1890-1964) + update_internal_adjacencies (object.rs:2673-2756) for
// chunks a host-side ChunkedVoxelGenerator produced (generation.rs:41-67): one CTA per chunk reads the chunk's 4096
// `Voxel`s (AoS, 48 contiguous bytes per (i, j) column), classifies it and stores NonUniform chunks as planes.
// sparseness: bit 0 has_only_empty_voxels, bit 1 is_void — the generator's own answer, taken as given like the reference
__global__ void __launch_bounds__(256) k_ingest_chunks(const unsigned char* __restrict__ src, const uint8_t* __restrict__ sparseness,
                                                       uint32_t n, DevChunk* __restrict__ chunks, unsigned char* __restrict__ voxels,
                                                       uint32_t* __restrict__ slot_counter) {
    __shared__ __align__(16) uint8_t s_fl[4096];
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_first, s_slot;
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
        const uint32_t sp = sparseness[c];
        DevChunk me{};
        me.slot = 0xFFFFFFFFu;
        if (sp & 2u) {  // Void
            if (tid == 0) chunks[c] = me;
            continue;
        }
        const uint4* p = reinterpret_cast<const uint4*>(src + (size_t)c * SLOT_BYTES + (size_t)tid * 48);
        const uint4 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        const uint32_t w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
        uint8_t ty[16], sd[16], fl[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int b = 3 * k;  // Voxel { voxel_type, signed_distance, flags } (lib.rs:60-66)
            ty[k] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
            sd[k] = (uint8_t)(w[(b + 1) >> 2] >> (8 * ((b + 1) & 3)));
            fl[k] = (uint8_t)(w[(b + 2) >> 2] >> (8 * ((b + 2) & 3)));
        }
        *reinterpret_cast<uint4*>(&s_fl[tid * 16]) = pack16(fl);
        if (tid < 8) s_cnt[tid] = 0;
        if (tid == 0) s_first = (uint32_t)ty[0] | ((uint32_t)fl[0] << 8);
        __syncthreads();
        const uint32_t first = s_first;
        bool same = true, valid = true;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            same = same && ty[k] == (uint8_t)first && fl[k] == (uint8_t)(first >> 8) && sd[k] == 0x80u;
            // the `Voxel` constructors' invariant (lib.rs:300-348, 451-461): EMPTY ⇔ the distance code is not negative;
            // the occupied ranges and the mesher read emptiness from the sign
            valid = valid && ((fl[k] & 1u) != 0) == ((sd[k] & 0x80u) == 0);
        }
        if (!valid) atomicOr(slot_counter + 1, 1u);
        const int uniform = __syncthreads_and(same) && !(sp & 1u);
        if (uniform) {
            // stored as one voxel that assumes full adjacency; fixed later if a neighbour disagrees (object.rs:1940-1952)
            me.kind = 1;
            me.u_type = (uint8_t)first;
            me.u_sd = (int8_t)-128;
            me.u_flags = (uint8_t)(first >> 8) | 0xFCu;
            if (tid == 0) chunks[c] = me;
            __syncthreads();
            continue;
        }
        uint8_t nf[16];
        const uint32_t empty_mask = refresh_column_flags(s_fl, ti, tj, nf);
        count_face_empties(empty_mask, ti, tj, s_cnt);
        if (tid == 0) s_slot = atomicAdd(slot_counter, 1u);
        __syncthreads();
        me.kind = 2;
        me.slot = s_slot;
        if (sp & 1u) {
            me.flags = 1u << 6;  // HAS_ONLY_EMPTY_VOXELS, all faces Empty
            for (int q = 0; q < 6; ++q) me.face[q] = 0;
        } else {
            for (int q = 0; q < 6; ++q) me.face[q] = s_cnt[q] == 256u ? 0 : (s_cnt[q] == 0u ? 1 : 2);
        }
        unsigned char* slot = voxels + (size_t)me.slot * SLOT_BYTES;
        *reinterpret_cast<uint4*>(slot + PLANE_SD + tid * 16) = pack16(sd);
        *reinterpret_cast<uint4*>(slot + PLANE_TYPE + tid * 16) = pack16(ty);
        *reinterpret_cast<uint4*>(slot + PLANE_FLAGS + tid * 16) = pack16(nf);
        if (tid == 0) chunks[c] = me;
        __syncthreads();
    }
}
cudaError_t launch_ingest_chunks(const unsigned char* src, const uint8_t* sparseness, uint32_t n, DevChunk* chunks,
                                 unsigned char* voxels, uint32_t* slot_counter, uint32_t grid, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_ingest_chunks<<<grid, 256, 0, st>>>(src, sparseness, n, chunks, voxels, slot_counter);
    return cudaGetLastError();
}

cudaError_t launch_extract_chunks(const ExtractArgs& a, uint32_t grid, cudaStream_t st) {
    if (a.n_ext == 0) return cudaSuccess;
    k_extract_chunks<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_repack_single(const DevChunk* chunks, const uint32_t nb[3], const unsigned char* voxels, const uint32_t org[3],
                                 const uint32_t occ_lo[3], const uint32_t occ_hi[3], DevChunk* out_chunk, unsigned char* out_slot,
                                 cudaStream_t st) {
    k_repack_single<<<1, 256, 0, st>>>(chunks, make_uint3(nb[0], nb[1], nb[2]), voxels, make_uint3(org[0], org[1], org[2]),
                                       make_uint3(occ_lo[0], occ_lo[1], occ_lo[2]), make_uint3(occ_hi[0], occ_hi[1], occ_hi[2]),
                                       out_chunk, out_slot);
    return cudaGetLastError();
}

}  // namespace ivx
