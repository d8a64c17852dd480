// SDF generation kernels (sm_100a).
//
//   k_fold<false>  conservative program specialisation per super-block
//   k_fold<true>   exact per-chunk specialisation: reproduces every cull /
//                  apply decision of compute_signed_distances_for_block
//                  (generation/sdf/atomic.rs:633-875) from the chunk AABB and
//                  the 14 sampled voxels its predicates read (atomic.rs:1661-1797)
//   k_eval         evaluates the specialised program over the 4096 voxels of a
//                  chunk, quantises, classifies, assigns voxel types and
//                  in-chunk adjacency flags, and writes the three 4 KiB planes
//                  (generation.rs:293-371, voxel_type.rs:54-168,
//                  object.rs:1890-1964, 2673-2756)
//
// The reference runs the whole post-order program on every chunk, touching a
// 16 KiB array per node even when the node is culled. Here the decisions are
// made first (warp per chunk, 14 lanes = 14 sampled voxels), constant
// sub-trees are folded, dropped operands are never evaluated, and only the
// surviving instructions are run on all voxels — with identical results.
#include <algorithm>

#include "common.cuh"
#include "packed.cuh"
#include "kernels.h"

namespace ivx {


// ---------------------------------------------------------------------------
// Noise frame of a MultifractalNoise node for one block
// (atomic.rs:1430-1451, 1520-1537)
struct NoiseFrame {
    f3 o, dxn, dyn, dzn;  // origin_for_noise, d*_for_noise
    float freq;           // unscaled_frequency
    bool rotated;
};
__device__ __forceinline__ NoiseFrame make_noise_frame(const ivx_node& n, f3 block_lo) {
    const float* M = n.transform_to_node_space;
    f3 origin = transform_point(M, block_lo);
    f3 dx = mk3(M[0], M[1], M[2]), dy = mk3(M[4], M[5], M[6]), dz = mk3(M[8], M[9], M[10]);
    float inverse_scale = norm3(dx);
    float scale = 1.0f / inverse_scale;
    NoiseFrame f;
    f.freq = n.p[0] * inverse_scale;
    f.o = scale * origin;
    f.dxn = scale * dx;
    f.dyn = scale * dy;
    f.dzn = scale * dz;
    f.rotated = fabsf(dx.x * inverse_scale - 1.0f) > 1e-6f || fabsf(dy.y * inverse_scale - 1.0f) > 1e-6f;
    return f;
}
// fbm_3d_offset(pos.z,1,pos.y,1,pos.x,1): simdnoise x := our z
__device__ __forceinline__ float noise_at(const ivx_node& n, float freq, f3 pos) {
    return fbm3(pos.z * freq, pos.y * freq, pos.x * freq, n.p[1], n.p[2], n.octaves, (int32_t)n.seed);
}
// Noise coordinates of voxel (i, j, k) of the block for the un-rotated block
// call (atomic.rs:1487-1502): simdnoise walks x (our k) as start + lane for one
// 8-wide vector then += 8, and y / z (our j / i) by repeated += 1.0.
__device__ __forceinline__ float block_noise_x(float start, int k) {
    return k < 8 ? start + (float)k : (start + (float)(k - 8)) + 8.0f;
}
__device__ __forceinline__ float noise_for_voxel(const ivx_node& n, const NoiseFrame& f, int i, int j, int k) {
    if (f.rotated) {
        f3 pos = (f.o + (float)i * f.dxn) + (float)j * f.dyn;
        for (int t = 0; t < k; ++t) pos = pos + f.dzn;
        return noise_at(n, f.freq, pos);
    }
    float zc = accumulate_ones(f.o.x, i), yc = accumulate_ones(f.o.y, j), xc = block_noise_x(f.o.z, k);
    return fbm3(xc * f.freq, yc * f.freq, zc * f.freq, n.p[1], n.p[2], n.octaves, (int32_t)n.seed);
}

// The 26 block test samples (atomic.rs:1683-1797): position of sample t and the
// lane (0..13) that holds the voxel whose value the reference reads for it.
__device__ __forceinline__ f3 test_position(int t, f3 o, f3 dx, f3 dy, f3 dz) {
    const float s = 15.0f, h = 7.5f;
    switch (t) {
        case 0: return o;
        case 1: return o + s * dx;
        case 2: return o + s * dy;
        case 3: return o + s * dz;
        case 4: return o + s * (dx + dy);
        case 5: return o + s * (dx + dz);
        case 6: return o + s * (dy + dz);
        case 7: return o + s * ((dx + dy) + dz);
        case 8: return o + h * dx;
        case 9: return (o + h * dx) + s * dy;
        case 10: return (o + h * dx) + s * dz;
        case 11: return (o + h * dx) + s * (dy + dz);
        case 12: return o + h * dy;
        case 13: return (o + h * dy) + s * dx;
        case 14: return (o + h * dy) + s * dz;
        case 15: return (o + h * dy) + s * (dx + dz);
        case 16: return o + h * dz;
        case 17: return (o + h * dz) + s * dx;
        case 18: return (o + h * dz) + s * dy;
        case 19: return (o + h * dz) + s * (dx + dy);
        case 20: return (o + h * dy) + h * dz;
        case 21: return ((o + s * dx) + h * dy) + h * dz;
        case 22: return (o + h * dx) + h * dz;
        case 23: return ((o + s * dy) + h * dx) + h * dz;
        case 24: return (o + h * dx) + h * dy;
        default: return ((o + s * dz) + h * dx) + h * dy;
    }
}
__constant__ int8_t c_test_lane[26] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 2, 3, 6, 0, 1, 3, 5, 0, 1, 2, 4, 8, 9, 10, 11, 12, 13};
// voxel (i,j,k) sampled by lane l < 14
__constant__ int8_t c_sample_ijk[14][3] = {{0, 0, 0},   {15, 0, 0},  {0, 15, 0},  {0, 0, 15}, {15, 15, 0},
                                           {15, 0, 15}, {0, 15, 15}, {15, 15, 15}, {0, 8, 8},  {15, 8, 8},
                                           {8, 0, 8},   {8, 15, 8},  {8, 8, 0},    {8, 8, 15}};

// ---------------------------------------------------------------------------
constexpr int FOLD_WARPS = 4;
constexpr int FOLD_TILE = 256;   // instructions classified per phase-1 tile
#ifndef IVX_FOLD_MAX_DEPTH
#define IVX_FOLD_MAX_DEPTH 64
#endif
constexpr int FOLD_MAX_DEPTH = IVX_FOLD_MAX_DEPTH;

struct FoldWarpSmem {
    uint8_t cls[FOLD_TILE];
    uint32_t seg_start[FOLD_MAX_DEPTH];
    float cval[FOLD_MAX_DEPTH];
    float ilo[FOLD_MAX_DEPTH];  // rigorous value range of the slot over the whole block
    float ihi[FOLD_MAX_DEPTH];
    uint8_t is_const[FOLD_MAX_DEPTH];
    float val[FOLD_MAX_DEPTH][32];  // EXACT only: sampled values per stack slot
};

// |simplex3| <= 1 over all gradient assignments (the 32.694… normalisation of
// simdnoise; checked by dense sampling in tests/test_noise_bound.py), so
// |fbm * noise_scale| <= amplitude. 0.5 % + 1e-3 absorbs f32 rounding.
__device__ __forceinline__ float noise_bound(const ivx_node& n) { return 1.005f * fabsf(n.p[3]) + 1e-3f; }

// ---------------------------------------------------------------------------
// Rigorous range of the noise over a block (warp-cooperative; every lane calls).
//
// simplex3(p) = 32.694 * sum over the 4 corners Q of p's simplex of t^4 (g_Q . d), d = p - Q, t = max(0.6 - |d|^2, 0).
// For a box B of noise-space points, every lattice point Q that can reach B (t > 0 somewhere in B) is
// enumerated, its own gradient is taken from the same hash simplex3 uses, and t^4 (g . d) is bounded over
// d in B - Q with interval arithmetic. A lattice point is a corner for some points of B and not for
// others, so each term enters as [min(lo, 0), max(hi, 0)]. The sum of those intervals contains simplex3(p)
// for every p in B. At the usual frequencies a block touches one or two simplex cells, and the range is a
// small fraction of [-1, 1]; when the box spans too many cells the octave falls back to [-1, 1].
constexpr int NOISE_RANGE_MAX_POINTS = 1024;
constexpr float NOISE_RANGE_MAX_EXTENT = 0.75f;

__device__ __forceinline__ void simplex3_range(float blo[3], float bhi[3], int32_t seed, int lane, float& out_lo,
                                               float& out_hi) {
    out_lo = -1.0f;
    out_hi = 1.0f;
    const float R = 0.7747f;  // sqrt(0.6) rounded up: beyond this distance t = 0
    const float G3 = 1.0f / 6.0f;
    float elo[3], ehi[3];
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        elo[c] = blo[c] - R;
        ehi[c] = bhi[c] + R;
        ok = ok && (fabsf(elo[c]) < 1.0e6f) && (fabsf(ehi[c]) < 1.0e6f) && (blo[c] <= bhi[c]);
    }
    if (!ok) return;  // also taken for NaN
    // a box wider than this spans several simplex cells: the bound is no better than [-1, 1]
    if (fmaxf(bhi[0] - blo[0], fmaxf(bhi[1] - blo[1], bhi[2] - blo[2])) > NOISE_RANGE_MAX_EXTENT) return;
    // Lattice points by planes S = I + J + K: the unskewed position is P = (I, J, K) - S/6, so
    // S = 2 (P.x + P.y + P.z), and on a plane I - S/6 and J - S/6 must lie in the box's x / y range.
    const float sum_lo = (elo[0] + elo[1]) + elo[2], sum_hi = (ehi[0] + ehi[1]) + ehi[2];
    const int s0 = (int)floorf(2.0f * sum_lo) - 1;
    const int ns = (int)ceilf(2.0f * sum_hi) + 1 - s0 + 1;
    const int na = (int)ceilf(ehi[0] - elo[0]) + 2, nb = (int)ceilf(ehi[1] - elo[1]) + 2;
    const int total = ns * na * nb;
    if (total > NOISE_RANGE_MAX_POINTS) return;
    float acc_lo = 0.0f, acc_hi = 0.0f;
    for (int q = lane; q < total; q += 32) {
        const int S = s0 + q / (na * nb);
        const float sg = (float)S * G3;
        const int I = (int)floorf(elo[0] + sg) + (q / nb) % na, J = (int)floorf(elo[1] + sg) + q % nb;
        const int K = S - I - J;
        const float g = G3 * (float)(I + J + K);
        const float P[3] = {(float)I - g, (float)J - g, (float)K - g};
        float dlo[3], dhi[3], minsq = 0.0f, maxsq = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float e = 4e-6f * (1.0f + fabsf(P[c]) + fabsf(blo[c]) + fabsf(bhi[c]));
            dlo[c] = (blo[c] - P[c]) - e;
            dhi[c] = (bhi[c] - P[c]) + e;
            const float a2 = dlo[c] * dlo[c], b2 = dhi[c] * dhi[c];
            maxsq += fmaxf(a2, b2);
            minsq += (dlo[c] <= 0.0f && dhi[c] >= 0.0f) ? 0.0f : fminf(a2, b2);
        }
        const float thi = fmaxf(0.6f - minsq * 0.99999f, 0.0f) * 1.00001f;
        if (thi <= 0.0f) continue;
        const float tlo = fmaxf(0.6f - maxsq * 1.00001f, 0.0f) * 0.99999f;
        // the lattice point's gradient: grad3d_dot of common.cuh
        uint32_t h = ((uint32_t)I * 1619u) ^ (uint32_t)seed;
        h = ((uint32_t)J * 31337u) ^ h;
        h = ((uint32_t)K * 6791u) ^ h;
        h = ((h * h) * 60493u) * h;
        h = (uint32_t)((int32_t)h >> 13) ^ h;
        const uint32_t h13 = h & 13u;
        const int ua = (h13 < 8u) ? 0 : 1;
        const int va = (h13 < 2u) ? 1 : ((h13 == 12u) ? 0 : 2);
        float ulo = dlo[ua], uhi = dhi[ua], vlo = dlo[va], vhi = dhi[va];
        if (h & 1u) { const float t = ulo; ulo = -uhi; uhi = -t; }
        if (h & 2u) { const float t = vlo; vlo = -vhi; vhi = -t; }
        const float glo = ulo + vlo, ghi = uhi + vhi;
        float t4lo = tlo * tlo, t4hi = thi * thi;
        t4lo = t4lo * t4lo * 0.9999f;
        t4hi = t4hi * t4hi * 1.0001f;
        const float clo = fminf(fminf(t4lo * glo, t4hi * glo), 0.0f);
        const float chi = fmaxf(fmaxf(t4lo * ghi, t4hi * ghi), 0.0f);
        acc_lo += clo;
        acc_hi += chi;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        acc_lo += __shfl_xor_sync(0xffffffffu, acc_lo, d);
        acc_hi += __shfl_xor_sync(0xffffffffu, acc_hi, d);
    }
    const float S = 32.69428253173828125f * 1.0001f;
    out_lo = fmaxf(acc_lo * S - 1e-5f, -1.0f);
    out_hi = fminf(acc_hi * S + 1e-5f, 1.0f);
    if (!(out_lo <= out_hi)) {  // cannot happen; stay safe
        out_lo = -1.0f;
        out_hi = 1.0f;
    }
}

// Range of `fbm * noise_scale` (the term a MultifractalNoise node adds, atomic.rs:1423-1507) over the voxels
// of the block [lo, lo + bs)^3 in root space. Always within [-A, A], A = noise_bound(n).
__device__ __forceinline__ void noise_term_range(const ivx_node& n, f3 lo, float bs, int lane, float& out_lo,
                                                 float& out_hi) {
    const float A = 1.005f * fabsf(n.p[3]) + 1e-3f;
    out_lo = -A;
    out_hi = A;
    const NoiseFrame f = make_noise_frame(n, lo);
    // positions walked by the block: o + i dxn + j dyn + k dzn, 0 <= i, j, k <= bs - 1 (voxel samples)
    const float o3[3] = {f.o.x, f.o.y, f.o.z};
    const float dx3[3] = {f.dxn.x, f.dxn.y, f.dxn.z}, dy3[3] = {f.dyn.x, f.dyn.y, f.dyn.z},
                dz3[3] = {f.dzn.x, f.dzn.y, f.dzn.z};
    float blo[3], bhi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = bs * dx3[c], b = bs * dy3[c], d = bs * dz3[c];
        float l = o3[c] + (fminf(a, 0.0f) + fminf(b, 0.0f) + fminf(d, 0.0f));
        float h = o3[c] + (fmaxf(a, 0.0f) + fmaxf(b, 0.0f) + fmaxf(d, 0.0f));
        // f32 rounding of the incremental position walks
        const float e = 1e-5f * (fabsf(l) + fabsf(h) + bs) + 1e-4f;
        l = (l - e) * f.freq;
        h = (h + e) * f.freq;
        // simdnoise x, y, z = our z, y, x
        blo[2 - c] = fminf(l, h);
        bhi[2 - c] = fmaxf(l, h);
    }
    const float lac = n.p[1], gain = n.p[2];
    const uint32_t n_oct = max(n.octaves & 0xFFu, 1u);
    float amp = 1.0f, sum_lo = 0.0f, sum_hi = 0.0f, sum_abs = 0.0f;
    for (uint32_t o = 0; o < n_oct; ++o) {
        if (o > 0) {
            amp = amp * gain;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float l = blo[c] * lac, h = bhi[c] * lac;
                const float e = 1e-6f * (fabsf(l) + fabsf(h));
                blo[c] = fminf(l, h) - e;
                bhi[c] = fmaxf(l, h) + e;
            }
        }
        float slo, shi;
        simplex3_range(blo, bhi, (int32_t)n.seed, lane, slo, shi);
        const float x = amp * slo, y = amp * shi;
        sum_lo += fminf(x, y);
        sum_hi += fmaxf(x, y);
        sum_abs += fabsf(amp);
    }
    const float ns = n.p[4];
    const float x = sum_lo * ns, y = sum_hi * ns;
    const float e = 1e-5f * sum_abs * fabsf(ns) + 1e-4f;
    out_lo = fmaxf(fminf(x, y) - e, -A);
    out_hi = fminf(fmaxf(x, y) + e, A);
    if (!(out_lo <= out_hi)) {
        out_lo = -A;
        out_hi = A;
    }
}

__device__ __forceinline__ float slack(float a, float b) { return 1e-3f + 2e-5f * (fabsf(a) + fabsf(b)); }

// moves `len` instructions from out[src..] down to out[dst..] (dst < src), warp-cooperatively
__device__ __forceinline__ void shift_down(Instr* out, uint32_t dst, uint32_t src, uint32_t len, int lane) {
    for (uint32_t off = 0; off < len; off += 32) {
        Instr v{};
        const bool in = off + lane < len;
        if (in) v = out[src + off + lane];
        __syncwarp();
        if (in) out[dst + off + lane] = v;
        __syncwarp();
    }
}

// Evaluates out[begin, end) — the specialised program of one operand — at voxel (si, sj, sk) of the
// block. Pruning is value-preserving, so this is exactly the value the reference holds for that voxel
// at this point of its sweep. `stk` is this lane's column of the warp's shared value stack.
__device__ float eval_segment_at_voxel(const ivx_node* __restrict__ nodes, const Instr* seg, uint32_t begin, uint32_t end,
                                       f3 lo, int si, int sj, int sk, float* stk /* stride 32 */) {
    int sp = 0;
    for (uint32_t q = begin; q < end; ++q) {
        const Instr in = seg[q];
        const uint32_t op = in.op_node >> 28;
        const ivx_node& n = nodes[in.op_node & 0x0FFFFFFFu];
        if (op == OP_CONST) {
            stk[32 * sp++] = in.value;
        } else if (op == OP_LEAF) {
            const float* M = n.transform_to_node_space;
            f3 origin = transform_point(M, lo);
            f3 dx = mk3(M[0], M[1], M[2]), dy = mk3(M[4], M[5], M[6]), dz = mk3(M[8], M[9], M[10]);
            f3 pos = (origin + (float)si * dx) + (float)sj * dy;
            for (int t = 0; t < sk; ++t) pos = pos + dz;
            stk[32 * sp++] = sd_leaf(n.kind, n.p, pos);
        } else if (op == OP_SCALE) {
            stk[32 * (sp - 1)] = stk[32 * (sp - 1)] * n.p[0];
        } else if (op == OP_NEG) {
            stk[32 * (sp - 1)] = -stk[32 * (sp - 1)];
        } else if (op == OP_NOISE) {
            NoiseFrame f = make_noise_frame(n, lo);
            stk[32 * (sp - 1)] = stk[32 * (sp - 1)] + noise_for_voxel(n, f, si, sj, sk) * n.p[4];
        } else {
            const float r = op_combine(n.kind, stk[32 * (sp - 2)], stk[32 * (sp - 1)], n.p[0], n.p[1]);
            sp -= 1;
            stk[32 * (sp - 1)] = r;
        }
    }
    return stk[0];
}

// Program specialisation for one block (a super-block of chunks, or one chunk).
//
// EXACT = true (block = one 16³ chunk): makes exactly the decisions
// compute_signed_distances_for_block makes for this chunk — leaf fill / evaluate
// from the chunk AABB (atomic.rs:654-668), noise / combine apply-or-skip from
// the AABB and the 14 sampled voxels (atomic.rs:765-869) — by evaluating the
// whole surviving program on those 14 voxels (one per lane).
// EXACT = false (super-block): only folds what holds for EVERY chunk inside.
//
// On top of the reference's decisions both modes carry a rigorous value range
// per operand (leaf SDFs are 1-Lipschitz, |noise| <= amplitude) and drop an
// operand of a smooth combine when the other one provably decides the result
// on every voxel of the block (|a - b| >= k ⇒ h = 0 ⇒ min/max returns one operand
// bit-exactly). This never changes a voxel, only skips arithmetic the reference
// spends on operands that cannot matter.
template <bool EXACT>
__global__ void __launch_bounds__(FOLD_WARPS * 32) k_fold(FoldArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t block = blockIdx.x * FOLD_WARPS + warp;
    if (block >= a.n_blocks) return;
    FoldWarpSmem& S = *reinterpret_cast<FoldWarpSmem*>(
        smem_raw + (size_t)warp * (EXACT ? sizeof(FoldWarpSmem) : offsetof(FoldWarpSmem, val)));

    // ---- block geometry ----
    f3 lo, hi;
    uint32_t parent = 0;
    uint32_t bc[3] = {0, 0, 0};
    if (a.explicit_origins) {
        lo = mk3(a.explicit_origins[3 * block], a.explicit_origins[3 * block + 1], a.explicit_origins[3 * block + 2]);
        hi = lo + mk3(16.0f, 16.0f, 16.0f);
    } else {
        bc[2] = block % a.nb[2];
        bc[1] = (block / a.nb[2]) % a.nb[1];
        bc[0] = block / (a.nb[2] * a.nb[1]);
        uint32_t px = bc[0] / a.ratio, py = bc[1] / a.ratio, pz = bc[2] / a.ratio;
        parent = a.parent_nb[0] == 0 ? 0u : (px * a.parent_nb[1] + py) * a.parent_nb[2] + pz;
        const float bs = (float)(a.block_chunks * 16u);
        // chunk_origin_in_root_space = origin as f32 - shifted_grid_center (generation.rs:314-315)
        lo = mk3((float)((bc[0] * a.block_chunks + a.first_chunk[0]) * 16u) - a.gp.shifted_center[0],
                 (float)((bc[1] * a.block_chunks + a.first_chunk[1]) * 16u) - a.gp.shifted_center[1],
                 (float)((bc[2] * a.block_chunks + a.first_chunk[2]) * 16u) - a.gp.shifted_center[2]);
        hi = lo + mk3(bs, bs, bs);
    }
    const Instr* __restrict__ pin = a.parent_instrs + a.parent_off[parent];
    const uint32_t plen = a.parent_len[parent];
    Instr* out = a.out_instrs + a.out_off[block];

    // chunks entirely beyond the grid are void without looking at the program
    // (generation.rs:299-312)
    if (EXACT && !a.explicit_origins) {
        if ((bc[0] + a.first_chunk[0]) * 16u >= a.gp.grid_shape[0] ||
            (bc[1] + a.first_chunk[1]) * 16u >= a.gp.grid_shape[1] ||
            (bc[2] + a.first_chunk[2]) * 16u >= a.gp.grid_shape[2] || a.gp.n_nodes == 0 || bc[0] < a.own_lo ||
            bc[0] >= a.own_hi) {
            if (lane == 0) {
                a.out_len[block] = 0;
                DevChunk c{};
                c.kind = 0;
                c.pre = PRE_VOID;
                a.chunks[block] = c;
            }
            return;
        }
    }

    int si = 0, sj = 0, sk = 0;
    if (EXACT) {
        int sl = lane < 14 ? lane : 0;
        si = c_sample_ijk[sl][0];
        sj = c_sample_ijk[sl][1];
        sk = c_sample_ijk[sl][2];
    }
    const bool prune = a.prune != 0;

    uint32_t olen = 0;
    int sp = 0;
    bool overflow = false;

    for (uint32_t base = 0; base < plen; base += FOLD_TILE) {
        const uint32_t tile_n = min((uint32_t)FOLD_TILE, plen - base);
        // ---- phase 1: AABB classification, one instruction per lane ----
        // bit 0-1: 1 = block outside domain_with_margin, 2 = inside the leaf's interior box
        // bit 2 (conservative leaf): some chunk may still be filled +margin; bit 3: ... -margin
        for (uint32_t t = lane; t < tile_n; t += 32) {
            Instr in = pin[base + t];
            uint32_t op = in.op_node >> 28;
            uint8_t cls = 0;
            if (op == OP_LEAF || op == OP_NOISE || op == OP_COMBINE) {
                const ivx_node& n = a.nodes[in.op_node & 0x0FFFFFFFu];
                f3 blo, bhi;
                aabb_of_transformed(n.transform_to_node_space, lo, hi, blo, bhi);
                if (EXACT) {
                    if (box_lies_outside(n.domain_lo, n.domain_hi, blo, bhi))
                        cls = 1;
                    else if (op == OP_LEAF && sym_box_contains(leaf_interior_half_extents(n), blo, bhi))
                        cls = 2;
                } else {
                    // conservative: must hold for every chunk inside this block despite f32 rounding
                    float mx = fmaxf(fmaxf(fabsf(blo.x), fabsf(bhi.x)),
                                     fmaxf(fmaxf(fabsf(blo.y), fabsf(bhi.y)), fmaxf(fabsf(blo.z), fabsf(bhi.z))));
                    float eps = 1e-4f * (1.0f + mx);
                    bool outside = (bhi.x + eps < n.domain_lo[0]) || (bhi.y + eps < n.domain_lo[1]) ||
                                   (bhi.z + eps < n.domain_lo[2]) || (n.domain_hi[0] + eps < blo.x) ||
                                   (n.domain_hi[1] + eps < blo.y) || (n.domain_hi[2] + eps < blo.z);
                    if (outside) {
                        cls = 1;
                    } else if (op != OP_LEAF) {
                        // a noise / combine can only be skipped by chunks lying outside its domain; if the whole
                        // block sits inside the domain box every chunk applies it
                        bool within_domain = (blo.x - eps >= n.domain_lo[0]) && (blo.y - eps >= n.domain_lo[1]) &&
                                             (blo.z - eps >= n.domain_lo[2]) && (bhi.x + eps <= n.domain_hi[0]) &&
                                             (bhi.y + eps <= n.domain_hi[1]) && (bhi.z + eps <= n.domain_hi[2]);
                        if (!within_domain) cls |= 4;
                    } else {
                        f3 h = leaf_interior_half_extents(n);
                        bool inside = (blo.x - eps >= -h.x) && (blo.y - eps >= -h.y) && (blo.z - eps >= -h.z) &&
                                      (bhi.x + eps <= h.x) && (bhi.y + eps <= h.y) && (bhi.z + eps <= h.z);
                        if (inside) {
                            cls = 2;
                        } else {
                            // which fills remain possible for some chunk of the block
                            bool within_domain = (blo.x - eps >= n.domain_lo[0]) && (blo.y - eps >= n.domain_lo[1]) &&
                                                 (blo.z - eps >= n.domain_lo[2]) && (bhi.x + eps <= n.domain_hi[0]) &&
                                                 (bhi.y + eps <= n.domain_hi[1]) && (bhi.z + eps <= n.domain_hi[2]);
                            bool misses_interior = (bhi.x + eps < -h.x) || (bhi.y + eps < -h.y) || (bhi.z + eps < -h.z) ||
                                                   (h.x + eps < blo.x) || (h.y + eps < blo.y) || (h.z + eps < blo.z) ||
                                                   h.x < 0.0f || h.y < 0.0f || h.z < 0.0f;
                            if (!within_domain) cls |= 4;
                            if (!misses_interior) cls |= 8;
                        }
                    }
                }
            }
            S.cls[t] = cls;
        }
        __syncwarp();

        // ---- phase 2: sequential fold (warp-uniform control flow) ----
        for (uint32_t t = 0; t < tile_n; ++t) {
            __syncwarp();  // lane 0's stack bookkeeping of the previous step is visible to all lanes
            const Instr in = pin[base + t];
            const uint32_t op = in.op_node >> 28;
            const uint32_t ni = in.op_node & 0x0FFFFFFFu;
            const uint8_t cls = S.cls[t];
            if (op == OP_CONST || (op == OP_LEAF && (cls & 3) != 0)) {
                float c = in.value;
                if (op == OP_LEAF) {
                    float m = a.nodes[ni].margin;
                    c = (cls & 3) == 1 ? m : -m;
                }
                if (sp >= FOLD_MAX_DEPTH) { overflow = true; break; }
                if (lane == 0) {
                    S.seg_start[sp] = olen;
                    S.cval[sp] = c;
                    S.ilo[sp] = c;
                    S.ihi[sp] = c;
                    S.is_const[sp] = 1;
                    out[olen] = Instr{(uint32_t)OP_CONST << 28, c};
                }
                olen += 1;
                sp += 1;
            } else if (op == OP_LEAF) {
                if (sp >= FOLD_MAX_DEPTH) { overflow = true; break; }
                const ivx_node& n = a.nodes[ni];
                const float* M = n.transform_to_node_space;
                // value range over the block: the primitive SDFs are 1-Lipschitz in node space
                f3 blo, bhi;
                aabb_of_transformed(M, lo, hi, blo, bhi);
                const f3 ctr = 0.5f * (blo + bhi);
                // The node transform is a similarity (translations, rotations, uniform scalings: atomic.rs:495-596),
                // so every sample of the block lies within scale * half-diagonal of the block centre's image —
                // tighter than the diagonal of the rotated block's axis-aligned box for rotated leaves.
                const float m_scale = fmaxf(norm3(mk3(M[0], M[1], M[2])), fmaxf(norm3(mk3(M[4], M[5], M[6])), norm3(mk3(M[8], M[9], M[10]))));
                const float rad = fminf(norm3(0.5f * (bhi - blo)),
                                        m_scale * (0.8660254f * (hi.x - lo.x)) * 1.00001f + 1e-4f);
                const float vc = sd_leaf(n.kind, n.p, ctr);
                const float e = 1e-3f + 1e-4f * (fabsf(vc) + rad);
                float rlo = vc - rad - e, rhi = vc + rad + e;
                if (!EXACT) {
                    if (cls & 4) rhi = fmaxf(rhi, n.margin);
                    if (cls & 8) rlo = fminf(rlo, -n.margin);
                    if (cls & 4) rlo = fminf(rlo, n.margin);
                    if (cls & 8) rhi = fmaxf(rhi, -n.margin);
                }
                if (lane == 0) {
                    S.seg_start[sp] = olen;
                    S.is_const[sp] = 0;
                    S.ilo[sp] = rlo;
                    S.ihi[sp] = rhi;
                    out[olen] = in;
                }
                olen += 1;
                sp += 1;
            } else if (op == OP_SCALE || op == OP_NEG) {
                const float s = op == OP_NEG ? -1.0f : a.nodes[ni].p[0];
                const bool tc = S.is_const[sp - 1] != 0;
                const float c0 = S.cval[sp - 1], l0 = S.ilo[sp - 1], h0 = S.ihi[sp - 1];
                const uint32_t seg = S.seg_start[sp - 1];
                __syncwarp();
                if (lane == 0) {
                    if (tc) {
                        const float c = op == OP_NEG ? -c0 : c0 * s;
                        S.cval[sp - 1] = c;
                        S.ilo[sp - 1] = c;
                        S.ihi[sp - 1] = c;
                        out[seg] = Instr{(uint32_t)OP_CONST << 28, c};
                    } else {
                        const float x = l0 * s, y = h0 * s;
                        const float e = slack(x, y);
                        S.ilo[sp - 1] = fminf(x, y) - e;
                        S.ihi[sp - 1] = fmaxf(x, y) + e;
                        out[olen] = in;
                    }
                }
                if (!tc) olen += 1;
            } else if (op == OP_NOISE) {
                const ivx_node& n = a.nodes[ni];
                bool apply = true;
                if (EXACT && cls == 1) {
                    // all_modified_signed_distances_at_block_test_positions_pass_predicate (atomic.rs:1510-1572):
                    // the operand's value at the 14 sampled voxels, then the 26 test positions
                    NoiseFrame f = make_noise_frame(n, lo);
                    const float cv = eval_segment_at_voxel(a.nodes, out, S.seg_start[sp - 1], olen, lo, si, sj, sk,
                                                           &S.val[0][lane]);
                    const int tl = lane < 26 ? lane : 0;
                    float v = __shfl_sync(0xffffffffu, cv, c_test_lane[tl]);
                    f3 tp = test_position(tl, f.o, f.dxn, f.dyn, f.dzn);
                    float mv = v + noise_at(n, f.freq, tp) * n.p[4];
                    bool pass = mv >= n.margin;
                    apply = !__all_sync(0xffffffffu, pass);
                }
                if (apply) {
                    const float l0 = S.ilo[sp - 1], h0 = S.ihi[sp - 1];
                    float nlo, nhi;  // range of the added term over the block
                    // The lattice analysis only pays where its result can matter. For the program's last
                    // instruction the range feeds nothing but the root saturation test below: when the operand
                    // is further from the saturated codes than the amplitude, the plain bound decides it already.
                    const float A = noise_bound(n);
                    const bool is_root_op = (base + t + 1 == plen) && sp == 1;
                    const bool undecided_without = (l0 - A < 2.03f) && (h0 + A > -2.5601f);
                    if (is_root_op && !undecided_without) {
                        nlo = -A;
                        nhi = A;
                    } else {
                        noise_term_range(n, lo, hi.x - lo.x, lane, nlo, nhi);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        S.is_const[sp - 1] = 0;
                        S.ilo[sp - 1] = (l0 + nlo) - slack(l0, nlo);
                        S.ihi[sp - 1] = (h0 + nhi) + slack(h0, nhi);
                        out[olen] = in;
                    }
                    olen += 1;
                }
            } else {  // OP_COMBINE
                const ivx_node& n = a.nodes[ni];
                const int ia = sp - 2, ib = sp - 1;
                const bool ac = S.is_const[ia] != 0, bcst = S.is_const[ib] != 0;
                const bool both_const = ac && bcst;
                const float ca = S.cval[ia], cb = S.cval[ib];
                const float alo = S.ilo[ia], ahi = S.ihi[ia], blo2 = S.ilo[ib], bhi2 = S.ihi[ib];
                const uint32_t seg_a = S.seg_start[ia], seg_b = S.seg_start[ib];
                const float k = n.p[0];
                int decision;  // 0 apply, 1 skip (keep child 1), 2 undecided (conservative only)
                if (EXACT) {
                    decision = 0;
                    if (cls == 1) {
                        // all_block_test_positions_pass_predicate (atomic.rs:796-806): both operands at the
                        // 14 sampled voxels (one per lane)
                        const float va = eval_segment_at_voxel(a.nodes, out, seg_a, seg_b, lo, si, sj, sk, &S.val[0][lane]);
                        const float vb = eval_segment_at_voxel(a.nodes, out, seg_b, olen, lo, si, sj, sk, &S.val[0][lane]);
                        const float r = op_combine(n.kind, va, vb, n.p[0], n.p[1]);
                        const bool pass = (lane >= 14) || (r >= n.margin);
                        if (__all_sync(0xffffffffu, pass)) decision = 1;
                    }
                } else {
                    if ((cls & 5) == 0) {
                        decision = 0;  // no chunk of the block lies outside the node's domain: always applied
                    } else if (both_const) {
                        bool pass = op_combine(n.kind, ca, cb, n.p[0], n.p[1]) >= n.margin;
                        decision = !pass ? 0 : (cls == 1 ? 1 : 2);
                    } else {
                        decision = 2;
                    }
                }
                // which operand provably decides the result on every voxel (if the op is applied)
                // 1: result == a, 2: result == b, 3: result == -b
                int keep = 0;
                if (prune && decision != 1 && !both_const) {
                    const float ek = k + slack(fmaxf(fabsf(alo), fabsf(ahi)), fmaxf(fabsf(blo2), fabsf(bhi2)));
                    if (n.kind == IVX_UNION) {
                        if (blo2 - ahi >= ek) keep = 1;
                        else if (alo - bhi2 >= ek) keep = 2;
                    } else if (n.kind == IVX_SUBTRACTION) {
                        if (alo + blo2 >= ek) keep = 1;
                        else if (ahi + bhi2 <= -ek) keep = 3;
                    } else {
                        if (alo - bhi2 >= ek) keep = 1;
                        else if (blo2 - ahi >= ek) keep = 2;
                    }
                    if (keep >= 2 && decision != 0) keep = 0;  // "== b" only holds when the op is applied
                }
                __syncwarp();
                if (decision == 1 || keep == 1) {
                    olen = seg_b;  // drop child 2's instructions; slot a is unchanged
                } else if (keep >= 2) {
                    const uint32_t len_b = olen - seg_b;
                    shift_down(out, seg_a, seg_b, len_b, lane);
                    olen = seg_a + len_b;
                    if (lane == 0) {
                        if (keep == 3) {
                            if (bcst) {
                                out[seg_a] = Instr{(uint32_t)OP_CONST << 28, -cb};
                                S.cval[ia] = -cb;
                                S.ilo[ia] = -cb;
                                S.ihi[ia] = -cb;
                            } else {
                                out[olen] = Instr{(uint32_t)OP_NEG << 28, 0.0f};
                                S.ilo[ia] = -bhi2;
                                S.ihi[ia] = -blo2;
                            }
                        } else {
                            S.cval[ia] = cb;
                            S.ilo[ia] = blo2;
                            S.ihi[ia] = bhi2;
                        }
                        S.is_const[ia] = bcst ? 1 : 0;
                    }
                    if (keep == 3 && !bcst) olen += 1;
                } else if (decision == 0 && both_const) {
                    float c = op_combine(n.kind, ca, cb, n.p[0], n.p[1]);
                    if (lane == 0) {
                        S.cval[ia] = c;
                        S.ilo[ia] = c;
                        S.ihi[ia] = c;
                        out[seg_a] = Instr{(uint32_t)OP_CONST << 28, c};
                    }
                    olen = seg_a + 1;
                } else {
                    if (lane == 0) {
                        // value range of the applied smooth op: min / max of the operands moved by h^2 / (4k),
                        // h = max(k - |u - v|, 0) (generation/sdf.rs:89-102). Over the block |u - v| lies in
                        // [dmin, dmax] (from the operand ranges), so the shift lies in
                        // [(k - dmax)^2 / 4k, (k - dmin)^2 / 4k] — far tighter than [0, k/4] once the operands
                        // are known to differ, which matters with the asteroid's k of 100 voxels.
                        float rl, rh, ulo, uhi, vlo, vhi;
                        if (n.kind == IVX_UNION) { ulo = alo; uhi = ahi; vlo = blo2; vhi = bhi2; }
                        else if (n.kind == IVX_SUBTRACTION) { ulo = -ahi; uhi = -alo; vlo = blo2; vhi = bhi2; }
                        else { ulo = -ahi; uhi = -alo; vlo = -bhi2; vhi = -blo2; }
                        float q = 0.0f, qmin = 0.0f;
                        if (k > 0.0f) {
                            const float dmin = fmaxf(0.0f, fmaxf(ulo - vhi, vlo - uhi));
                            const float dmax = fmaxf(uhi - vlo, vhi - ulo);
                            const float hmax = fmaxf(k - dmin, 0.0f), hmin = fmaxf(k - dmax, 0.0f);
                            q = fminf((hmax * hmax) * (0.25f / k) * 1.0001f + 1e-5f, 0.25f * k);
                            qmin = fmaxf((hmin * hmin) * (0.25f / k) * 0.9999f - 1e-5f, 0.0f);
                        }
                        if (n.kind == IVX_UNION) {
                            rl = fminf(alo, blo2) - q;
                            rh = fminf(ahi, bhi2) - qmin;
                        } else if (n.kind == IVX_SUBTRACTION) {
                            rl = fmaxf(alo, -bhi2) + qmin;
                            rh = fmaxf(ahi, -blo2) + q;
                        } else {
                            rl = fmaxf(alo, blo2) + qmin;
                            rh = fmaxf(ahi, bhi2) + q;
                        }
                        if (decision == 2) {  // may also be skipped in some chunks: result == a there
                            rl = fminf(rl, alo);
                            rh = fmaxf(rh, ahi);
                        }
                        const float e = slack(rl, rh);
                        S.ilo[ia] = rl - e;
                        S.ihi[ia] = rh + e;
                        S.is_const[ia] = 0;
                        out[olen] = in;
                    }
                    olen += 1;
                }
                sp -= 1;
            }
        }
        __syncwarp();
        if (overflow) break;
    }
    __syncwarp();

    // ---- root saturation: the whole block quantises to one code (lib.rs:195-201, 240-243) ----
    // every code > 100 ⇒ the chunks are void; every code == -128 ⇒ maximally inside everywhere
    if (!overflow && a.saturate && sp == 1 && !S.is_const[0]) {
        const float rl = S.ilo[0], rh = S.ihi[0];
        float c = 0.0f;
        bool sat = false;
        if (rl >= 2.03f) { c = 1000.0f; sat = true; }
        else if (rh <= -2.5601f) { c = -1000.0f; sat = true; }
        if (sat) {
            __syncwarp();
            if (lane == 0) {
                out[0] = Instr{(uint32_t)OP_CONST << 28, c};
                S.is_const[0] = 1;
                S.cval[0] = c;
            }
            olen = 1;
            __syncwarp();
        }
    }

    if (lane == 0) {
        if (overflow) atomicExch(a.error_flag, 1u);
        a.out_len[block] = olen;
    }
    if (!EXACT) return;

    // ---- exact level: pre-classify the chunk ----
    __syncwarp();
    if (lane == 0) {
        DevChunk c{};
        c.pre = PRE_ACTIVE;
        if (a.chunks) {
            if (olen == 1 && S.is_const[0]) {
                const int code = sd_encode(S.cval[0]);
                const uint32_t ox = (bc[0] + a.first_chunk[0]) * 16u, oy = (bc[1] + a.first_chunk[1]) * 16u,
                               oz = (bc[2] + a.first_chunk[2]) * 16u;
                const bool full_in_grid = ox + 16u <= a.gp.grid_shape[0] && oy + 16u <= a.gp.grid_shape[1] &&
                                          oz + 16u <= a.gp.grid_shape[2];
                if (code > 100) {
                    c.pre = PRE_VOID;
                } else if (code == -128 && full_in_grid && a.gp.types.kind == 0) {
                    c.pre = PRE_UNIFORM;
                    c.kind = 1;
                    c.u_type = (uint8_t)a.gp.types.same_type;
                    c.u_sd = -128;
                    c.u_flags = 0xFC;
                    if (a.occ) {
                        const uint32_t o3[3] = {ox, oy, oz};
                        for (int d = 0; d < 3; ++d) {
                            atomicMin(&a.occ[d], o3[d]);
                            atomicMax(&a.occ[3 + d], o3[d] + 15u);
                        }
                    }
                }
            }
            a.chunks[block] = c;
        }
        if (c.pre == PRE_ACTIVE) {
            // stack depth the evaluator needs for this list
            int d = 0, md = 0;
            for (uint32_t q = 0; q < olen; ++q) {
                uint32_t o2 = out[q].op_node >> 28;
                if (o2 == OP_LEAF || o2 == OP_CONST) { d++; md = max(md, d); }
                else if (o2 == OP_COMBINE) d--;
            }
            atomicMax(a.max_depth, (uint32_t)md);
        }
    }
}

// One out-of-line copy of the 3-D simplex noise for k_eval: the kernel reaches it from three places (generic,
// rotated and compacted noise paths) and its instruction footprint, not its call overhead, is what costs
// (ncu: "no instruction" stalls from instruction-cache misses).
__device__ __noinline__ float simplex3_call(float x, float y, float z, int32_t seed) { return simplex3(x, y, z, seed); }
__device__ __noinline__ float fbm3_call(float x, float y, float z, float lacunarity, float gain, uint32_t octaves,
                                        int32_t seed) {
    const uint32_t oct = octaves & 0xFFu;
    float amp = 1.0f;
    float result = simplex3_call(x, y, z, seed);
    for (uint32_t o = 1; o < oct; ++o) {
        x = x * lacunarity;
        y = y * lacunarity;
        z = z * lacunarity;
        amp = amp * gain;
        result = simplex3_call(x, y, z, seed) * amp + result;
    }
    return result;
}

// ---- 3-D simplex noise from a per-chunk lattice table, two voxels per thread --------------------------
// simplex3 (common.cuh) spends more than half of its instructions on the per-corner gradient: the integer hash,
// the selection of two of (x, y, z) and their sign flips, all on the half-rate ALU pipe. The hash depends only on
// the lattice point, and one octave of a 16³ chunk touches few lattice points, so the compacted final-noise path
// builds a table of gradient vectors g ∈ {-1, 0, +1}³ per (chunk, octave) and evaluates a corner's gradient as
// fma(g.x, x, fma(g.y, y, g.z·z)): products with ±1 / 0 are exact and exactly one component is 0, so the result
// is RN(±u ± v), simdnoise's `xor_sign(u) + xor_sign(v)` up to the sign of an exact zero (which cannot reach a
// stored voxel code). The rest of the arithmetic runs on voxel pairs with packed FADD2 / FFMA2 (packed.cuh),
// operation for operation simplex3's.
constexpr int TAB3_CAP = 1536;           // float4 entries
constexpr float CELL3_LIMIT = 1024.0f;   // |cell| < 2^10, strides < 2^11: every integer-valued f32 term stays below 2^22

struct Tab3 {
    float base;   // MAGIC - entry of the first lattice cell
    float sa, sb; // entry strides of the x and y cell axes (z has stride 1)
    float s3;     // sa + sb + 1
    uint32_t addr;  // shared-memory byte address of entry 0, minus 16 * MAGIC_BITS
};

// gradient vector of lattice point (I, J, K): grad3d_dot's hash and selection (common.cuh)
__device__ __forceinline__ float4 grad3_vector(int32_t seed, int32_t I, int32_t J, int32_t K) {
    uint32_t hash = ((uint32_t)I * 1619u) ^ (uint32_t)seed;
    hash = ((uint32_t)J * 31337u) ^ hash;
    hash = ((uint32_t)K * 6791u) ^ hash;
    hash = ((hash * hash) * 60493u) * hash;
    hash = (uint32_t)((int32_t)hash >> 13) ^ hash;
    const uint32_t h13 = hash & 13u;
    const float su = (hash & 1u) ? -1.0f : 1.0f, sv = (hash & 2u) ? -1.0f : 1.0f;
    float4 g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (h13 < 8u) g.x = su; else g.y = su;        // u = x or y
    if (h13 < 2u) g.y = sv;                         // v = y
    else if (h13 == 12u) g.x = sv;                  // v = x (u is y here)
    else g.z = sv;                                  // v = z
    return g;
}

__device__ __forceinline__ float gdot3(const float4 g, float x, float y, float z) {
    return __fmaf_rn(g.x, x, __fmaf_rn(g.y, y, g.z * z));
}

__device__ __forceinline__ f2 simplex3_tab2(f2 x, f2 y, f2 z, const Tab3& T, f2 nz) {
    const float F3 = 1.0f / 3.0f, G3 = 1.0f / 6.0f, G33 = 3.0f / 6.0f - 1.0f;
    const f2 f = mul2(bc2(F3), add2(add2(x, y), z), nz);
    const f2 xf = floor2(add2(x, f)), yf = floor2(add2(y, f)), zf = floor2(add2(z, f));
    const f2 g = mul2(bc2(G3), add2(add2(xf, yf), zf), nz);
    const f2 x0 = sub2(x, sub2(xf, g)), y0 = sub2(y, sub2(yf, g)), z0 = sub2(z, sub2(zf, g));
    // corner steps as exact 0 / 1 floats: a = [x0 >= y0], b = [y0 >= z0], c = [x0 >= z0]
    const f2 a = make_float2(x0.x >= y0.x ? 1.0f : 0.0f, x0.y >= y0.y ? 1.0f : 0.0f);
    const f2 b = make_float2(y0.x >= z0.x ? 1.0f : 0.0f, y0.y >= z0.y ? 1.0f : 0.0f);
    const f2 c = make_float2(x0.x >= z0.x ? 1.0f : 0.0f, x0.y >= z0.y ? 1.0f : 0.0f);
    const f2 one = bc2(1.0f);
    const f2 na = sub2(one, a), nb = sub2(one, b), nc = sub2(one, c);
    const f2 i1 = mul2(a, c, nz), j1 = mul2(na, b, nz), k1 = mul2(nc, nb, nz);                 // and
    const f2 i2 = sub2(add2(a, c), i1), j2 = sub2(add2(na, b), j1), k2 = sub2(one, mul2(c, b, nz));  // or / nand

    const f2 sa = bc2(T.sa), sb = bc2(T.sb);
    const f2 e0 = fma2(xf, sa, fma2(yf, sb, add2(zf, bc2(T.base))));
    const f2 e1 = fma2(i1, sa, fma2(j1, sb, add2(k1, e0)));
    const f2 e2 = fma2(i2, sa, fma2(j2, sb, add2(k2, e0)));
    const f2 e3 = add2(e0, bc2(T.s3));
    const float4 g0a = tab_load(e0.x, T.addr), g0b = tab_load(e0.y, T.addr);
    const float4 g1a = tab_load(e1.x, T.addr), g1b = tab_load(e1.y, T.addr);
    const float4 g2a = tab_load(e2.x, T.addr), g2b = tab_load(e2.y, T.addr);
    const float4 g3a = tab_load(e3.x, T.addr), g3b = tab_load(e3.y, T.addr);

    const f2 cG3 = bc2(G3), cF3 = bc2(F3), cG33 = bc2(G33), c06 = bc2(0.6f);
    const f2 x1 = add2(sub2(x0, i1), cG3), y1 = add2(sub2(y0, j1), cG3), z1 = add2(sub2(z0, k1), cG3);
    const f2 x2 = add2(sub2(x0, i2), cF3), y2 = add2(sub2(y0, j2), cF3), z2 = add2(sub2(z0, k2), cF3);
    const f2 x3 = add2(x0, cG33), y3 = add2(y0, cG33), z3 = add2(z0, cG33);
#define IVX_T3(X, Y, Z) sub2(sub2(sub2(c06, mul2(X, X, nz)), mul2(Y, Y, nz)), mul2(Z, Z, nz))
    f2 t0 = IVX_T3(x0, y0, z0), t1 = IVX_T3(x1, y1, z1), t2 = IVX_T3(x2, y2, z2), t3 = IVX_T3(x3, y3, z3);
#undef IVX_T3
    // (t >= 0) ? t : 0, NaN → 0
    t0 = max2c(t0, 0.0f); t1 = max2c(t1, 0.0f); t2 = max2c(t2, 0.0f); t3 = max2c(t3, 0.0f);
    t0 = mul2(t0, t0, nz); t1 = mul2(t1, t1, nz); t2 = mul2(t2, t2, nz); t3 = mul2(t3, t3, nz);
    t0 = mul2(t0, t0, nz); t1 = mul2(t1, t1, nz); t2 = mul2(t2, t2, nz); t3 = mul2(t3, t3, nz);
    const f2 v0 = mul2(t0, make_float2(gdot3(g0a, x0.x, y0.x, z0.x), gdot3(g0b, x0.y, y0.y, z0.y)), nz);
    const f2 v1 = mul2(t1, make_float2(gdot3(g1a, x1.x, y1.x, z1.x), gdot3(g1b, x1.y, y1.y, z1.y)), nz);
    const f2 v2 = mul2(t2, make_float2(gdot3(g2a, x2.x, y2.x, z2.x), gdot3(g2b, x2.y, y2.y, z2.y)), nz);
    const f2 v3 = mul2(t3, make_float2(gdot3(g3a, x3.x, y3.x, z3.x), gdot3(g3b, x3.y, y3.y, z3.y)), nz);
    const f2 p1 = add2(v3, v2);
    const f2 p2 = add2(p1, v1);
    return mul2(add2(p2, v0), bc2(32.69428253173828125f), nz);
}

// One octave of a whole k-column (16 voxels of thread (i, j)) as 8 voxel pairs. Out of line on purpose: k_eval keeps
// its column of 16 distances live across the noise node; the call boundary parks them once per octave instead of
// raising the kernel's register count (a resident CTA per SM).
__device__ __noinline__ void noise_column_tab(float* acc, const float* coord, int ti, int tj, Tab3 T, float amp,
                                              bool first_octave, float neg_zero) {
    __builtin_assume(__isShared(coord));
    const f2 nz = bc2(neg_zero);
    const f2 y = bc2(coord[16 + tj]), z = bc2(coord[32 + ti]);
#pragma unroll 1
    for (int k = 0; k < 16; k += 2) {
        const f2 sv = simplex3_tab2(make_float2(coord[k], coord[k + 1]), y, z, T, nz);
        acc[k] = first_octave ? sv.x : sv.x * amp + acc[k];
        acc[k + 1] = first_octave ? sv.y : sv.y * amp + acc[k + 1];
    }
}

// Lattice cells one octave of the chunk touches, from the noise coordinates of its 8 corner voxels (coord[0..15] x,
// [16..31] y, [32..47] z): the cell is monotone in each coordinate (every step of simplex3's cell computation is a
// monotone f32 operation) and each coordinate is monotone in its voxel index, so the corners bound it. Called by
// warp 0; lane 0 writes min[3], max[3] and the usable flag to `cell`.
__device__ __forceinline__ void lattice3_bounds(const float* coord, int lane, int* cell) {
    const float F3 = 1.0f / 3.0f;
    const float cx = coord[(lane & 1) ? 15 : 0], cy = coord[16 + ((lane & 2) ? 15 : 0)], cz = coord[32 + ((lane & 4) ? 15 : 0)];
    const float ff = F3 * ((cx + cy) + cz);
    const float c3[3] = {floorf(cx + ff), floorf(cy + ff), floorf(cz + ff)};
    int ok = 1, lo3[3], hi3[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) ok &= (fabsf(c3[d]) < CELL3_LIMIT) ? 1 : 0;  // false for NaN too
#pragma unroll
    for (int d = 0; d < 3; ++d) lo3[d] = hi3[d] = ok ? (int)c3[d] : 0;
#pragma unroll
    for (int sh = 1; sh < 8; sh <<= 1) {
        ok &= __shfl_xor_sync(0xffffffffu, ok, sh);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo3[d] = min(lo3[d], __shfl_xor_sync(0xffffffffu, lo3[d], sh));
            hi3[d] = max(hi3[d], __shfl_xor_sync(0xffffffffu, hi3[d], sh));
        }
    }
    if (lane == 0) {
        const int64_t entries = ok ? (int64_t)(hi3[0] - lo3[0] + 2) * (hi3[1] - lo3[1] + 2) * (hi3[2] - lo3[2] + 2) : 0;
        cell[6] = (ok && entries <= TAB3_CAP) ? 1 : 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            cell[d] = lo3[d];
            cell[3 + d] = hi3[d];
        }
    }
}

// fills the gradient table for the cell bounds in `cell` (all threads) and returns its addressing constants
__device__ __forceinline__ Tab3 lattice3_build(const int* cell, float4* tab, int32_t seed, int tid) {
    const int cx0 = cell[0], cy0 = cell[1], cz0 = cell[2];
    const uint32_t Dy = (uint32_t)(cell[4] - cy0 + 2), Dz = (uint32_t)(cell[5] - cz0 + 2);
    const uint32_t entries = (uint32_t)(cell[3] - cx0 + 2) * Dy * Dz;
    for (uint32_t e = tid; e < entries; e += 256u) {
        const uint32_t dz_ = e % Dz, dy_ = (e / Dz) % Dy, dx_ = e / (Dz * Dy);
        tab[e] = grad3_vector(seed, cx0 + (int)dx_, cy0 + (int)dy_, cz0 + (int)dz_);
    }
    Tab3 T;
    T.sa = (float)(Dy * Dz);
    T.sb = (float)Dz;
    T.s3 = (float)(Dy * Dz + Dz + 1u);
    T.base = MAGIC - (((float)cx0 * T.sa + (float)cy0 * T.sb) + (float)cz0);
    T.addr = (uint32_t)__cvta_generic_to_shared(tab) - MAGIC_BITS * 16u;
    return T;
}

// ---------------------------------------------------------------------------
// k_eval: CTA (256 threads) per active chunk; thread (i, j) owns the 16 voxels
// of one k-column so the reference's incremental `position += dz` walk
// (atomic.rs:1614-1626) is reproduced exactly. The operand stack lives in
// shared memory ([level][k][thread], conflict-free); its top stays in
// registers.
constexpr int EVAL_THREADS = 256;

__device__ __forceinline__ float* stack_level(float* smem_stack, float* spill, int level, int smem_levels) {
    return level < smem_levels ? smem_stack + (size_t)level * 4096 : spill + (size_t)(level - smem_levels) * 4096;
}

#ifndef IVX_EVAL_CTAS
#define IVX_EVAL_CTAS 3
#endif
__global__ void __launch_bounds__(EVAL_THREADS, IVX_EVAL_CTAS) k_eval(EvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_stack = reinterpret_cast<float*>(smem_raw);
    __shared__ float s_coord[48];
    __shared__ uint32_t s_wsum[EVAL_THREADS / 32];
    __shared__ __align__(16) float4 s_tab3[TAB3_CAP];  // lattice gradient table of the current (chunk, octave)
    __shared__ int s_cell[8];                          // cell bounds min[3], max[3], usable flag
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tj = tid & 15;
    float* spill = a.spill ? a.spill + (size_t)blockIdx.x * a.spill_levels * 4096 : nullptr;

    for (uint32_t work = blockIdx.x; work < a.n_active; work += gridDim.x) {
        const uint32_t chunk = a.active ? a.active[work] : work;
        const Instr* __restrict__ prog = a.instrs + a.off[chunk];
        const uint32_t plen = a.len[chunk];

        f3 lo;
        uint32_t org[3] = {0, 0, 0};
        if (a.explicit_origins) {
            lo = mk3(a.explicit_origins[3 * chunk], a.explicit_origins[3 * chunk + 1], a.explicit_origins[3 * chunk + 2]);
        } else {
            uint32_t ck = chunk % a.nb[2], cj = (chunk / a.nb[2]) % a.nb[1], ci = chunk / (a.nb[2] * a.nb[1]);
            org[0] = (ci + a.first_i) * 16u;
            org[1] = cj * 16u;
            org[2] = ck * 16u;
            lo = mk3((float)org[0] - a.gp.shifted_center[0], (float)org[1] - a.gp.shifted_center[1],
                     (float)org[2] - a.gp.shifted_center[2]);
        }

        float top[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) top[k] = 0.02f * 127.0f;
        int sp = 0;

        for (uint32_t pc = 0; pc < plen; ++pc) {
            const Instr in = prog[pc];
            const uint32_t op = in.op_node >> 28;
            const ivx_node& n = a.nodes[in.op_node & 0x0FFFFFFFu];
            if (op == OP_LEAF || op == OP_CONST) {
                if (sp > 0) {
                    float* lv = stack_level(s_stack, spill, sp - 1, a.smem_levels);
#pragma unroll
                    for (int k = 0; k < 16; ++k) lv[k * 256 + tid] = top[k];
                }
                sp++;
                if (op == OP_CONST) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) top[k] = in.value;
                } else {
                    const float* M = n.transform_to_node_space;
                    f3 origin = transform_point(M, lo);
                    f3 dx = mk3(M[0], M[1], M[2]), dy = mk3(M[4], M[5], M[6]), dz = mk3(M[8], M[9], M[10]);
                    f3 pos = (origin + (float)ti * dx) + (float)tj * dy;
                    const uint32_t kind = n.kind;
                    const float p0 = n.p[0], p1 = n.p[1], p2 = n.p[2];
                    const float pp[3] = {p0, p1, p2};
                    // one straight-line loop per primitive: the kind test stays outside the unrolled body, so an
                    // executed leaf streams ~300 instructions instead of hopping through 1300 (i-cache)
                    if (kind == IVX_SPHERE) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            top[k] = norm3(pos) - p0;
                            pos = pos + dz;
                        }
                    } else if (kind == IVX_CAPSULE) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            top[k] = sd_leaf(IVX_CAPSULE, pp, pos);
                            pos = pos + dz;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            top[k] = sd_leaf(IVX_BOX, pp, pos);
                            pos = pos + dz;
                        }
                    }
                }
            } else if (op == OP_SCALE) {
                const float s = n.p[0];
#pragma unroll
                for (int k = 0; k < 16; ++k) top[k] = top[k] * s;
            } else if (op == OP_NEG) {
#pragma unroll
                for (int k = 0; k < 16; ++k) top[k] = -top[k];
            } else if (op == OP_NOISE) {
                NoiseFrame f = make_noise_frame(n, lo);
                // Last instruction of a voxel (not raw) evaluation: a voxel whose value cannot leave the
                // saturated code range whatever the noise adds (|noise * scale| <= A) is stored as 127 / -128
                // without evaluating the noise (lib.rs:195-201); decided per voxel, skipped per warp.
                const bool final_sat = (pc + 1 == plen) && a.raw_out == nullptr && a.saturate_final_noise;
                const float A = 1.005f * fabsf(n.p[3]) + 1e-3f;
                const float ns = n.p[4];
                const float lac = n.p[1], gain = n.p[2];
                const uint32_t oct = n.octaves;
                const int32_t seed = (int32_t)n.seed;
                if (final_sat && !f.rotated) {
                    // ---- compacted, octave-by-octave evaluation ----
                    // After octaves 0..o the remaining ones can move the result by at most
                    // |noise_scale| * sum_{i>o} gain^i (|simplex3| <= 1), so a voxel whose partial value is further
                    // than that beyond the saturated code range is finished. The surviving voxels of the whole
                    // chunk are compacted into a dense list each octave, so that lanes do not idle beside a few
                    // unfinished neighbours. The arithmetic of a surviving voxel is fbm3's, operation for operation.
                    float* res = s_stack;                                            // [k][thread] partial fBm sums
                    uint16_t* list = reinterpret_cast<uint16_t*>(s_stack + 4096);    // voxel indices still needed
                    const uint32_t n_oct = max(oct & 0xFFu, 1u);
                    float total = 1.0f;  // sum of the octave amplitudes
                    {
                        float am = 1.0f;
                        for (uint32_t o = 1; o < n_oct; ++o) {
                            am = am * gain;
                            total = total + am;
                        }
                    }
                    uint32_t mask = 0xFFFFu;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float v = top[k];
                        if (v - A >= 2.5401f) { top[k] = 1000.0f; mask &= ~(1u << k); }
                        else if (v + A <= -2.5601f) { top[k] = -1000.0f; mask &= ~(1u << k); }
                    }
                    __syncthreads();  // stack levels are free: this is the last instruction (sp == 1)
                    if (tid < 16) {
                        s_coord[tid] = block_noise_x(f.o.z, tid) * f.freq;            // simdnoise x = our k
                        s_coord[16 + tid] = accumulate_ones(f.o.y, tid) * f.freq;     // y = our j
                        s_coord[32 + tid] = accumulate_ones(f.o.x, tid) * f.freq;     // z = our i
                    }
                    float amp = 1.0f, done = 0.0f;
                    for (uint32_t o = 0; o < n_oct; ++o) {
                        if (o > 0) {
                            amp = amp * gain;
                            if (tid < 48) s_coord[tid] = s_coord[tid] * lac;
                            const float R = fabsf(ns) * (total - done) * 1.005f + (1e-3f + 1e-5f * fabsf(ns) * total);
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
                                if (mask & (1u << k)) {
                                    const float c = top[k] + res[k * 256 + tid] * ns;
                                    if (c - R >= 2.5401f) { top[k] = 1000.0f; mask &= ~(1u << k); }
                                    else if (c + R <= -2.5601f) { top[k] = -1000.0f; mask &= ~(1u << k); }
                                }
                            }
                        }
                        // CTA-wide compaction of the surviving voxels
                        const uint32_t cnt = __popc(mask);
                        uint32_t incl = cnt;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                            if ((tid & 31) >= d) incl += y;
                        }
                        if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
                        __syncthreads();
                        uint32_t base = 0, n_list = 0;
#pragma unroll
                        for (int w = 0; w < EVAL_THREADS / 32; ++w) {
                            const uint32_t ws = s_wsum[w];
                            if (w < (tid >> 5)) base += ws;
                            n_list += ws;
                        }
                        if (n_list == 0) break;  // CTA-uniform
                        {
                            uint32_t pos = base + incl - cnt, m = mask;
                            while (m) {
                                const int k = __ffs(m) - 1;
                                m &= m - 1;
                                list[pos++] = (uint16_t)(tid * 16 + k);
                            }
                        }
                        __syncthreads();
                        // (~300 surviving voxels per chunk and octave on the asteroid: one evaluation per thread, so the
                        // lattice table of the generic path below does not pay here)
                        for (uint32_t it = tid; it < n_list; it += EVAL_THREADS) {
                            const uint32_t idx = list[it];
                            const float sx = s_coord[idx & 15u], sy = s_coord[16 + ((idx >> 4) & 15u)],
                                        sz = s_coord[32 + (idx >> 8)];
                            const float sv = simplex3_call(sx, sy, sz, seed);
                            float* r = &res[(idx & 15u) * 256 + (idx >> 4)];
                            *r = o == 0 ? sv : sv * amp + *r;
                        }
                        done = done + amp;
                        __syncthreads();
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (mask & (1u << k)) top[k] = top[k] + res[k * 256 + tid] * ns;
                } else if (f.rotated) {
                    f3 pos = (f.o + (float)ti * f.dxn) + (float)tj * f.dyn;
                    for (int k = 0; k < 16; ++k) {
                        const float v = top[k];
                        bool need = true;
                        if (final_sat) {
                            if (v - A >= 2.5401f) { need = false; top[k] = 1000.0f; }
                            else if (v + A <= -2.5601f) { need = false; top[k] = -1000.0f; }
                        }
                        if (need) {
                            float nv = fbm3_call(pos.z * f.freq, pos.y * f.freq, pos.x * f.freq, lac, gain, oct, seed);
                            top[k] = v + nv * ns;
                        }
                        pos = pos + f.dzn;
                    }
                } else {
                    // ---- the whole chunk, octave by octave: fbm3 (common.cuh) with the octave loop outermost, so that
                    // one lattice gradient table serves all 4096 evaluations of an octave ----
                    const uint32_t n_oct = max(oct & 0xFFu, 1u);
                    __syncthreads();  // s_coord / s_tab3 may still be read by the previous noise node
                    if (tid < 16) {
                        s_coord[tid] = block_noise_x(f.o.z, tid) * f.freq;            // simdnoise x = our k
                        s_coord[16 + tid] = accumulate_ones(f.o.y, tid) * f.freq;     // y = our j
                        s_coord[32 + tid] = accumulate_ones(f.o.x, tid) * f.freq;     // z = our i
                    }
                    float acc[16];
                    float amp = 1.0f;
                    for (uint32_t o = 0; o < n_oct; ++o) {
                        if (o > 0) {
                            amp = amp * gain;
                            if (tid < 48) s_coord[tid] = s_coord[tid] * lac;
                        }
                        __syncthreads();
                        if (tid < 32) lattice3_bounds(s_coord, tid, s_cell);
                        __syncthreads();
                        if (s_cell[6]) {
                            const Tab3 T = lattice3_build(s_cell, s_tab3, seed, tid);
                            __syncthreads();
                            noise_column_tab(acc, s_coord, ti, tj, T, amp, o == 0, a.neg_zero);
                        } else {
                            const float yc = s_coord[16 + tj], zc = s_coord[32 + ti];
#pragma unroll 1
                            for (int k = 0; k < 16; ++k) {
                                const float sv = simplex3_call(s_coord[k], yc, zc, seed);
                                acc[k] = o == 0 ? sv : sv * amp + acc[k];
                            }
                        }
                        __syncthreads();  // before the coordinates and the table are overwritten
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) top[k] = top[k] + acc[k] * ns;
                }
            } else {  // OP_COMBINE
                const float* lv = stack_level(s_stack, spill, sp - 2, a.smem_levels);
                const uint32_t kind = n.kind;
                const float ks = n.p[0], qik = n.p[1];
                // same for the combine: kind and the hard / smooth choice are block-uniform
                if (ks == 0.0f) {
                    if (kind == IVX_UNION) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) top[k] = fminf(lv[k * 256 + tid], top[k]);
                    } else if (kind == IVX_SUBTRACTION) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) top[k] = fmaxf(lv[k * 256 + tid], -top[k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; ++k) top[k] = fmaxf(lv[k * 256 + tid], top[k]);
                    }
                } else if (kind == IVX_UNION) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) top[k] = smooth_union(lv[k * 256 + tid], top[k], ks, qik);
                } else if (kind == IVX_SUBTRACTION) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) top[k] = -smooth_union(-lv[k * 256 + tid], top[k], ks, qik);
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) top[k] = -smooth_union(-lv[k * 256 + tid], -top[k], ks, qik);
                }
                sp--;
            }
        }

        if (a.raw_out) {
            float* o = a.raw_out + (size_t)chunk * 4096 + tid * 16;
#pragma unroll
            for (int k = 0; k < 16; k += 4)
                *reinterpret_cast<float4*>(o + k) = make_float4(top[k], top[k + 1], top[k + 2], top[k + 3]);
            continue;
        }

        // ---- quantise + classify (generation.rs:330-357) ----
        const bool col_in = (org[0] + ti < a.gp.grid_shape[0]) && (org[1] + tj < a.gp.grid_shape[1]);
        uint32_t void_mask = 0;
        uint4 pk;
        {
            uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = 0u;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                int c = 127;
                const bool in_grid = col_in && (org[2] + k < a.gp.grid_shape[2]);
                if (in_grid) c = sd_encode(top[k]);
                // out-of-grid voxels do not influence is_void (generation.rs:336-340)
                if (!in_grid || c > 100) void_mask |= 1u << k;
                w[k >> 2] |= (uint32_t)(uint8_t)(int8_t)c << (8 * (k & 3));
            }
        }
        const int all_void = __syncthreads_and(void_mask == 0xFFFFu);

        if (all_void) {
            if (tid == 0) {
                DevChunk cd = a.chunks[chunk];
                cd.kind = 0;
                cd.pre = PRE_VOID;
                cd.flags = 0;
                a.chunks[chunk] = cd;
            }
        } else {
            // the chunk stays PRE_ACTIVE: k_types assigns voxel types, flags and the chunk descriptor
            unsigned char* slot = a.voxels + (size_t)a.slot_of[chunk] * SLOT_BYTES;
            *reinterpret_cast<uint4*>(slot + PLANE_SD + tid * 16) = pk;
        }
    }
}

// ---------------------------------------------------------------------------
// k_eval_blocks: compute_signed_distances_for_block_preserving_gradients<SIZE, SIZE³>
// (atomic.rs:877-998) for many small blocks (SIZE = 1 or 2): the whole program, no culling, one thread
// per voxel with its operand stack in local memory. Used by the meta-graph compiler's surface probes
// (meta.rs:2705-2748), batched over all instances of a node.
__global__ void k_eval_blocks(const ivx_node* __restrict__ nodes, uint32_t n_nodes, const float* __restrict__ origins,
                              uint32_t n_blocks, int size, float* __restrict__ out) {
    const uint32_t count = (uint32_t)(size * size * size);
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_blocks * count) return;
    const uint32_t b = gid / count, v = gid % count;
    const int i = (int)(v / (uint32_t)(size * size)), j = (int)((v / (uint32_t)size) % (uint32_t)size), k = (int)(v % (uint32_t)size);
    const f3 lo = mk3(origins[3 * b], origins[3 * b + 1], origins[3 * b + 2]);
    float stk[FOLD_MAX_DEPTH];
    int sp = 0;
    if (n_nodes == 0) {
        out[gid] = 0.02f * 127.0f;
        return;
    }
    for (uint32_t q = 0; q < n_nodes; ++q) {
        const ivx_node& n = nodes[q];
        if (n.kind <= IVX_BOX) {
            const float* M = n.transform_to_node_space;
            f3 origin = transform_point(M, lo);
            f3 dx = mk3(M[0], M[1], M[2]), dy = mk3(M[4], M[5], M[6]), dz = mk3(M[8], M[9], M[10]);
            f3 pos = (origin + (float)i * dx) + (float)j * dy;
            for (int t = 0; t < k; ++t) pos = pos + dz;
            if (sp < FOLD_MAX_DEPTH) stk[sp] = sd_leaf(n.kind, n.p, pos);
            sp++;
        } else if (n.kind == IVX_SCALING) {
            stk[sp - 1] = stk[sp - 1] * n.p[0];
        } else if (n.kind == IVX_MULTIFRACTAL_NOISE) {
            NoiseFrame f = make_noise_frame(n, lo);
            float nv;
            if (f.rotated) {
                f3 pos = (f.o + (float)i * f.dxn) + (float)j * f.dyn;
                for (int t = 0; t < k; ++t) pos = pos + f.dzn;
                nv = noise_at(n, f.freq, pos);
            } else {
                // block call with SIZE < 8: x = start + lane, y / z by repeated += 1.0
                const float zc = accumulate_ones(f.o.x, i), yc = accumulate_ones(f.o.y, j), xc = f.o.z + (float)k;
                nv = fbm3(xc * f.freq, yc * f.freq, zc * f.freq, n.p[1], n.p[2], n.octaves, (int32_t)n.seed);
            }
            stk[sp - 1] = stk[sp - 1] + nv * n.p[4];
        } else if (n.kind >= IVX_UNION) {
            const float r = op_combine(n.kind, stk[sp - 2], stk[sp - 1], n.p[0], n.p[1]);
            sp -= 1;
            stk[sp - 1] = r;
        }
    }
    out[gid] = stk[0];
}

cudaError_t launch_eval_blocks(const ivx_node* nodes, uint32_t n_nodes, const float* origins, uint32_t n_blocks, int size,
                               float* out, cudaStream_t st) {
    const uint32_t total = n_blocks * (uint32_t)(size * size * size);
    if (total == 0) return cudaSuccess;
    k_eval_blocks<<<(total + 127) / 128, 128, 0, st>>>(nodes, n_nodes, origins, n_blocks, size, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// host launch wrappers
size_t fold_smem_bytes(bool exact) {
    return (size_t)FOLD_WARPS * (exact ? sizeof(FoldWarpSmem) : offsetof(FoldWarpSmem, val));
}

cudaError_t launch_fold(bool exact, const FoldArgs& a, cudaStream_t st) {
    if (a.n_blocks == 0) return cudaSuccess;
    const uint32_t grid = (a.n_blocks + FOLD_WARPS - 1) / FOLD_WARPS;
    const size_t smem = fold_smem_bytes(exact);
    if (exact) {
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(k_fold<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_done = true;
        }
        k_fold<true><<<grid, FOLD_WARPS * 32, smem, st>>>(a);
    } else {
        k_fold<false><<<grid, FOLD_WARPS * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// shared memory of k_eval: the operand stack levels; at least 24 KiB because the compacted final-noise path
// reuses the (then empty) stack area for its partial sums and its voxel list
static size_t eval_smem_bytes(int smem_levels) {
    return std::max<size_t>((size_t)smem_levels * 4096 * sizeof(float), 24576);
}

cudaError_t launch_eval(const EvalArgs& a, uint32_t grid, cudaStream_t st) {
    if (a.n_active == 0 || grid == 0) return cudaSuccess;
    const size_t smem = eval_smem_bytes(a.smem_levels);
    cudaFuncSetAttribute(k_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_eval<<<grid, EVAL_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

int eval_max_blocks_per_sm(int smem_levels) {
    int nb = 0;
    const size_t smem = eval_smem_bytes(smem_levels);
    cudaFuncSetAttribute(k_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_eval, EVAL_THREADS, smem);
    return nb;
}

}  // namespace ivx
