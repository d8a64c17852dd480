// Voxel types and per-voxel derived flags (sm_100a).
//
//   k_types  second half of chunk generation, for every chunk k_eval left non-void:
//            voxel types (generation.rs:359-365, voxel_type.rs:54-168), uniform detection
//            (object.rs:1913-1918), IS_EMPTY + in-chunk adjacency flags (object.rs:2673-2756),
//            face distributions (object.rs:1920-1936, 2967-2981), occupied ranges
//            (object.rs:1187-1280); writes the type and flag planes.
//
// GradientNoise types are the arg-max over `n_types` evaluations of simdnoise's 4-D simplex noise
// per voxel — for a 1024³ asteroid that is ~10⁹ evaluations and the largest single cost of the whole
// path. The per-voxel arithmetic cannot be shared between voxels without changing rounding, but the
// *lattice* part can: a 16³ chunk at the usual frequencies touches only a few simplex cells, so the
// permutation-table hash chain (4 dependent byte gathers per corner, 5 corners per evaluation) and the
// gradient selection it feeds are hoisted into a per-chunk table of gradient vectors, one float4 in
// {-1, 0, +1}⁴ per lattice point. A corner's gradient dot product then is
//     g.x·x + (g.y·y + (g.z·z + g.w·w))
// which reproduces simdnoise's `a + (b + c)` of sign-selected components exactly: multiplying by ±1 is
// exact, and the one zero term only ever adds ±0. Differences are confined to the sign of an exact
// zero, which no comparison downstream can see (the result only feeds `noise > max_noise`).
// Corner ranks are computed as small floats so that almost all of the work runs on the FMA pipe
// (FADD / FMUL / FFMA / IMAD) instead of the half-rate ALU pipe (ISETP / SEL / LOP3).
#include "common.cuh"
#include "packed.cuh"
#include "kernels.h"

namespace ivx {

constexpr int TYPES_THREADS = 256;
#ifndef IVX_TYPES_UNROLL
#define IVX_TYPES_UNROLL 1
#endif
constexpr int TYPES_UNROLL = IVX_TYPES_UNROLL;  // voxel pairs in flight per thread
#ifndef IVX_TYPES_CTAS
#define IVX_TYPES_CTAS 3   // resident CTAs per SM the register budget is set for
#endif
constexpr int TAB_CAP = 1024;        // gradient-table entries (float4) per batch of types
constexpr int MAX_TYPES = 255;
constexpr float CELL_LIMIT = 4096.0f;  // |cell index| < 2^12 and byte strides <= 2^14: every address term stays below 2^26 (TypeTab2)

// Per-type constants of the table walk, every value duplicated into an f32x2 pair (packed operands are register
// pairs; a pair loaded from shared memory needs no MOVs). Table positions are carried as *byte addresses held in
// the bit pattern of a float*: an integer N < 2^24 reinterpreted as f32 is the subnormal (or smallest-binade normal)
// value N·2^-149, sums and products with small integers of such values are exact in f32 (all terms here are multiples
// of 16·2^-149 below 2^28·2^-149), and FADD / FFMA handle subnormals at full rate. The walk therefore ends with the
// shared-memory byte address of the gradient entry *as the float's bits*: no integer instruction per table load.
struct TypeTab2 {
    f2 x;       // noise x coordinate of this type (already multiplied by voxel_type_frequency)
    f2 base;    // bits: byte address of the entry of lattice cell (0,0,0,0) (may be "negative": exact all the same)
    f2 sa, sb, sc;  // bits: byte strides of axes 0..2 (axis 3 has stride 16)
    f2 s4;      // sa + sb + sc + 16: byte step to the far corner
};
__device__ __forceinline__ float addr_as_float(int32_t n) {
    // n·2^-149, exactly (|n| < 2^24): non-negative n is the bit pattern itself
    return n >= 0 ? __uint_as_float((uint32_t)n) : -__uint_as_float((uint32_t)(-n));
}
__device__ __forceinline__ float4 tab_load_bits(float e) { return lds128(__float_as_uint(e)); }

// simdnoise simplex_4d (see common.cuh simplex4_t for the direct restatement) for the voxel pair (y.x, y.y) — z and w
// are the same for both — with the gradient of each corner read from the chunk's lattice table; every packed operation
// rounds like the scalar one it stands for.
// `yzw` = y + (z + w) (the same for every type of a voxel pair), `T` = the type's constants in shared memory.
__device__ __forceinline__ f2 simplex4_tab2(f2 y, f2 z, f2 w, f2 yzw, const TypeTab2& T, f2 nz) {
    const float F4 = 0.309016994f, G4 = 0.138196601f;
    const float G24 = 2.0f * 0.138196601f, G34 = 3.0f * 0.138196601f, G44 = 4.0f * 0.138196601f;
    const f2 x = T.x;
    const f2 s = mul2(bc2(F4), add2(x, yzw), nz);
    const f2 ips = floor2(add2(x, s)), jps = floor2(add2(y, s)), kps = floor2(add2(z, s)), lps = floor2(add2(w, s));
    const f2 t = mul2(add2(ips, add2(jps, add2(kps, lps))), bc2(G4), nz);
    const f2 x0 = sub2(x, sub2(ips, t)), y0 = sub2(y, sub2(jps, t)), z0 = sub2(z, sub2(kps, t)), w0 = sub2(w, sub2(lps, t));

    const f2 pxy = gt2(x0, y0), pxz = gt2(x0, z0), pxw = gt2(x0, w0);
    const f2 pyz = gt2(y0, z0), pyw = gt2(y0, w0), pzw = gt2(z0, w0);
    const f2 rx = add2(add2(pxy, pxz), pxw);
    const f2 ry = add2(add2(sub2(bc2(1.0f), pxy), pyz), pyw);
    const f2 rz = add2(sub2(sub2(bc2(2.0f), pxz), pyz), pzw);
    const f2 rw = sub2(sub2(sub2(bc2(6.0f), rx), ry), rz);
    const f2 i1 = gtc2(rx, 2.5f), j1 = gtc2(ry, 2.5f), k1 = gtc2(rz, 2.5f), l1 = gtc2(rw, 2.5f);
    const f2 i2 = gtc2(rx, 1.5f), j2 = gtc2(ry, 1.5f), k2 = gtc2(rz, 1.5f), l2 = gtc2(rw, 1.5f);
    const f2 i3 = min2c(rx, 1.0f), j3 = min2c(ry, 1.0f), k3 = min2c(rz, 1.0f), l3 = min2c(rw, 1.0f);

    // byte addresses of the five corners' gradient entries, as float bits (see TypeTab2)
    const f2 sa = T.sa, sb = T.sb, sc = T.sc, sd = bc2(__uint_as_float(16u));
    const f2 e0 = fma2(ips, sa, fma2(jps, sb, fma2(kps, sc, fma2(lps, sd, T.base))));
    const f2 e1 = fma2(i1, sa, fma2(j1, sb, fma2(k1, sc, fma2(l1, sd, e0))));
    const f2 e2 = fma2(i2, sa, fma2(j2, sb, fma2(k2, sc, fma2(l2, sd, e0))));
    const f2 e3 = fma2(i3, sa, fma2(j3, sb, fma2(k3, sc, fma2(l3, sd, e0))));
    const f2 e4 = add2(e0, T.s4);
    const float4 g0a = tab_load_bits(e0.x), g0b = tab_load_bits(e0.y);
    const float4 g1a = tab_load_bits(e1.x), g1b = tab_load_bits(e1.y);
    const float4 g2a = tab_load_bits(e2.x), g2b = tab_load_bits(e2.y);
    const float4 g3a = tab_load_bits(e3.x), g3b = tab_load_bits(e3.y);
    const float4 g4a = tab_load_bits(e4.x), g4b = tab_load_bits(e4.y);

    const f2 c1 = bc2(G4), c2 = bc2(G24), c3 = bc2(G34), c4 = bc2(G44), one = bc2(1.0f), half = bc2(0.5f);
    const f2 x1 = add2(sub2(x0, i1), c1), y1 = add2(sub2(y0, j1), c1), z1 = add2(sub2(z0, k1), c1), w1 = add2(sub2(w0, l1), c1);
    const f2 x2 = add2(sub2(x0, i2), c2), y2 = add2(sub2(y0, j2), c2), z2 = add2(sub2(z0, k2), c2), w2 = add2(sub2(w0, l2), c2);
    const f2 x3 = add2(sub2(x0, i3), c3), y3 = add2(sub2(y0, j3), c3), z3 = add2(sub2(z0, k3), c3), w3 = add2(sub2(w0, l3), c3);
    const f2 x4 = add2(sub2(x0, one), c4), y4 = add2(sub2(y0, one), c4), z4 = add2(sub2(z0, one), c4), w4 = add2(sub2(w0, one), c4);

#define IVX_T2(X, Y, Z, W) \
    sub2(sub2(sub2(sub2(half, mul2(X, X, nz)), mul2(Y, Y, nz)), mul2(Z, Z, nz)), mul2(W, W, nz))
    f2 t0 = IVX_T2(x0, y0, z0, w0), t1 = IVX_T2(x1, y1, z1, w1), t2 = IVX_T2(x2, y2, z2, w2),
       t3 = IVX_T2(x3, y3, z3, w3), t4 = IVX_T2(x4, y4, z4, w4);
#undef IVX_T2
    t0 = max2c(t0, 0.0f); t1 = max2c(t1, 0.0f); t2 = max2c(t2, 0.0f); t3 = max2c(t3, 0.0f); t4 = max2c(t4, 0.0f);
    f2 q0 = mul2(t0, t0, nz), q1 = mul2(t1, t1, nz), q2 = mul2(t2, t2, nz), q3 = mul2(t3, t3, nz), q4 = mul2(t4, t4, nz);
    q0 = mul2(q0, q0, nz); q1 = mul2(q1, q1, nz); q2 = mul2(q2, q2, nz); q3 = mul2(q3, q3, nz); q4 = mul2(q4, q4, nz);
    const f2 n0 = mul2(q0, make_float2(gdot(g0a, x0.x, y0.x, z0.x, w0.x), gdot(g0b, x0.y, y0.y, z0.y, w0.y)), nz);
    const f2 n1 = mul2(q1, make_float2(gdot(g1a, x1.x, y1.x, z1.x, w1.x), gdot(g1b, x1.y, y1.y, z1.y, w1.y)), nz);
    const f2 n2 = mul2(q2, make_float2(gdot(g2a, x2.x, y2.x, z2.x, w2.x), gdot(g2b, x2.y, y2.y, z2.y, w2.y)), nz);
    const f2 n3 = mul2(q3, make_float2(gdot(g3a, x3.x, y3.x, z3.x, w3.x), gdot(g3b, x3.y, y3.y, z3.y, w3.y)), nz);
    const f2 n4 = mul2(q4, make_float2(gdot(g4a, x4.x, y4.x, z4.x, w4.x), gdot(g4b, x4.y, y4.y, z4.y, w4.y)), nz);
    return mul2(add2(n0, add2(n1, add2(n2, add2(n3, n4)))), bc2(62.77772078955791f), nz);
}

// the lattice cell of a noise-space point, exactly as simplex4 computes it
__device__ __forceinline__ void simplex4_cell(float x, float y, float z, float w, float c[4]) {
    const float F4 = 0.309016994f;
    const float s = F4 * (x + (y + (z + w)));
    c[0] = floorf(x + s);
    c[1] = floorf(y + s);
    c[2] = floorf(z + s);
    c[3] = floorf(w + s);
}

// simdnoise grad4 (common.cuh) as a vector: grad = g · (x, y, z, t)
__device__ __forceinline__ float4 grad4_vector(int32_t h) {
    const float sa = (h & 1) ? -1.0f : 1.0f, sb = (h & 2) ? -1.0f : 1.0f, sc = (h & 4) ? -1.0f : 1.0f;
    if (h < 8) return make_float4(sa, sb, sc, 0.0f);    // (x, y, z)
    if (h < 16) return make_float4(sa, sb, 0.0f, sc);   // (x, y, t)
    if (h < 24) return make_float4(sa, 0.0f, sb, sc);   // (x, z, t)
    return make_float4(0.0f, sa, sb, sc);               // (y, z, t)
}

// x coordinate of type t: simdnoise walks x as start + lane inside one 8-wide vector, then += 8
__device__ __forceinline__ float type_axis_coordinate(uint32_t t) {
    float xc = 0.0f + (float)(t & 7u);
    for (uint32_t v = 0; v < (t >> 3); ++v) xc = xc + 8.0f;
    return xc;
}

struct TypesSmem {
    float4 tab[TAB_CAP];
    TypeTab2 tt[TAB_CAP / 16];          // per-type constants of the current batch (a type's table has >= 2^4 entries)
    float best[16 * TYPES_THREADS];     // [k][thread]: running maximum, only between the batches of a multi-batch chunk
    uint16_t type_pair[8 * TYPES_THREADS];  // [k / 2][thread]: the winning types of a voxel pair (low byte = even k)
    int8_t sd[4096];
    int cmin[MAX_TYPES + 1][4];
    int cmax[MAX_TYPES + 1][4];
    uint16_t tab_off[MAX_TYPES + 1];    // entry offset of the type's table inside its batch
    uint8_t new_batch[MAX_TYPES + 1];
    uint8_t perm[256];
    uint32_t cnt[16];
    int fallback;
    uint8_t first_type;
};

__global__ void __launch_bounds__(TYPES_THREADS, IVX_TYPES_CTAS) k_types(TypesArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TypesSmem& S = *reinterpret_cast<TypesSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tj = tid & 15;
    S.perm[tid] = c_perm[tid];
    __syncthreads();
    const uint32_t tab_base = (uint32_t)__cvta_generic_to_shared(&S.tab[0]);
    const uint32_t n_types = a.gp.types.n_types;

    for (uint32_t work = blockIdx.x; work < a.n_active; work += gridDim.x) {
        const uint32_t chunk = a.active ? a.active[work] : work;
        DevChunk cd = a.chunks[chunk];
        if (cd.pre != PRE_ACTIVE) continue;  // k_eval found the chunk void (block-uniform branch)
        const uint32_t slot = a.slot_of[chunk];
        unsigned char* slot_ptr = a.voxels + (size_t)slot * SLOT_BYTES;

        uint32_t org[3];
        {
            const uint32_t ck = chunk % a.nb[2], cj = (chunk / a.nb[2]) % a.nb[1], ci = chunk / (a.nb[2] * a.nb[1]);
            org[0] = (ci + a.first_i) * 16u;
            org[1] = cj * 16u;
            org[2] = ck * 16u;
        }
        const f3 lo = mk3((float)org[0] - a.gp.shifted_center[0], (float)org[1] - a.gp.shifted_center[1],
                          (float)org[2] - a.gp.shifted_center[2]);

        // ---- the chunk's signed-distance codes (written by k_eval) ----
        const uint4 pk = *reinterpret_cast<const uint4*>(slot_ptr + PLANE_SD + tid * 16);
        *reinterpret_cast<uint4*>(&S.sd[tid * 16]) = pk;
        int8_t codes[16];
        {
            const uint32_t* wv = reinterpret_cast<const uint32_t*>(&pk);
#pragma unroll
            for (int k = 0; k < 16; ++k) codes[k] = (int8_t)((wv[k >> 2] >> (8 * (k & 3))) & 0xFFu);
        }
        uint32_t empty_mask = 0, m128_mask = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (codes[k] >= 0) empty_mask |= 1u << k;
            if (codes[k] == -128) m128_mask |= 1u << k;
        }
        if (tid < 16) S.cnt[tid] = (tid >= 6 && tid < 9) ? 0xFFFFFFFFu : 0u;
        const int any_nonempty = __syncthreads_or(empty_mask != 0xFFFFu);
        const int all_m128 = __syncthreads_and(m128_mask == 0xFFFFu);

        // ---- voxel types (generation.rs:359-365, voxel_type.rs) ----
        uint8_t types[16];
        if (!any_nonempty) {
#pragma unroll
            for (int k = 0; k < 16; ++k) types[k] = 255;
        } else if (a.gp.types.kind == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) types[k] = (uint8_t)a.gp.types.same_type;
        } else {
            if (a.noise_evaluations && tid == 0) atomicAdd(a.noise_evaluations, 4096ull * n_types);
            // gradient_4d_offset(0, n, o.z, 16, o.y, 16, o.x, 16): x = type axis,
            // y / z / w = our k / j / i, each walked by repeated += 1.0
            const float ft = a.gp.types.voxel_type_frequency, fn = a.gp.types.noise_frequency;
            const int32_t seed = (int32_t)a.gp.types.seed;
            const float wc = accumulate_ones(lo.x, ti) * fn;
            const float zc = accumulate_ones(lo.y, tj) * fn;
            const f2 nz = bc2(a.neg_zero), Z2 = bc2(zc), W2 = bc2(wc), ZW2 = bc2(zc + wc);

            // ---- lattice cells touched by the chunk, per type: the cell of a voxel is monotone in each
            // voxel index (every step of the computation is a monotone f32 operation), so the 8 corner
            // voxels bound it ----
            for (uint32_t q = tid; q < n_types * 4u; q += TYPES_THREADS) {
                S.cmin[q >> 2][q & 3] = INT_MAX;
                S.cmax[q >> 2][q & 3] = INT_MIN;
            }
            if (tid == 0) S.fallback = 0;
            __syncthreads();
            for (uint32_t q = tid; q < n_types * 8u; q += TYPES_THREADS) {
                const uint32_t t = q >> 3, cv = q & 7u;
                const float x = type_axis_coordinate(t) * ft;
                const float w = accumulate_ones(lo.x, (cv & 4u) ? 15 : 0) * fn;
                const float z = accumulate_ones(lo.y, (cv & 2u) ? 15 : 0) * fn;
                const float y = accumulate_ones(lo.z, (cv & 1u) ? 15 : 0) * fn;
                float c[4];
                simplex4_cell(x, y, z, w, c);
                bool ok = true;
#pragma unroll
                for (int d = 0; d < 4; ++d) ok = ok && (fabsf(c[d]) < CELL_LIMIT);  // false for NaN too
                if (!ok) {
                    S.fallback = 1;
                } else {
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        atomicMin(&S.cmin[t][d], (int)c[d]);
                        atomicMax(&S.cmax[t][d], (int)c[d]);
                    }
                }
            }
            __syncthreads();
            // batches of types whose tables fit TAB_CAP together (sequential, n_types <= 255)
            if (tid == 0 && !S.fallback) {
                uint32_t run = 0;
                for (uint32_t t = 0; t < n_types; ++t) {
                    uint64_t e = 1;
                    for (int d = 0; d < 4; ++d) e *= (uint64_t)((int64_t)S.cmax[t][d] - S.cmin[t][d] + 2);
                    if (e > (uint64_t)TAB_CAP) {
                        S.fallback = 1;
                        break;
                    }
                    const bool nb = (t == 0) || (run + (uint32_t)e > (uint32_t)TAB_CAP);
                    if (nb) run = 0;
                    S.new_batch[t] = nb ? 1 : 0;
                    S.tab_off[t] = (uint16_t)run;
                    run += (uint32_t)e;
                }
            }
            __syncthreads();

            if (S.fallback) {
                // cells too spread out (very high frequencies) or out of the exact-integer range:
                // direct evaluation with the permutation table, as in common.cuh
                float yacc = lo.z;
                for (int k = 0; k < 16; ++k) {
                    const float yc = yacc * fn;
                    float best = 0.0f;
                    uint32_t best_t = 0;
                    for (uint32_t t = 0; t < n_types; ++t) {
                        const float nv = simplex4_t(type_axis_coordinate(t) * ft, yc, zc, wc, seed, S.perm);
                        if (t == 0 || nv > best) {
                            best = nv;
                            best_t = t;
                        }
                    }
                    types[k] = (uint8_t)best_t;
                    yacc = yacc + 1.0f;
                }
            } else {
                uint32_t t0 = 0;
                while (t0 < n_types) {
                    uint32_t t1 = t0 + 1;
                    while (t1 < n_types && !S.new_batch[t1]) ++t1;
                    // ---- gradient table of the batch ----
                    {
                        uint32_t n_entries = S.tab_off[t1 - 1];
                        {
                            uint32_t e = 1;
                            for (int d = 0; d < 4; ++d) e *= (uint32_t)(S.cmax[t1 - 1][d] - S.cmin[t1 - 1][d] + 2);
                            n_entries += e;
                        }
                        uint32_t t = t0;
                        for (uint32_t e = tid; e < n_entries; e += TYPES_THREADS) {
                            while (t + 1 < t1 && e >= S.tab_off[t + 1]) ++t;
                            uint32_t r = e - S.tab_off[t];
                            const uint32_t Dd = (uint32_t)(S.cmax[t][3] - S.cmin[t][3] + 2);
                            const uint32_t Dc = (uint32_t)(S.cmax[t][2] - S.cmin[t][2] + 2);
                            const uint32_t Db = (uint32_t)(S.cmax[t][1] - S.cmin[t][1] + 2);
                            const uint32_t d3 = r % Dd; r /= Dd;
                            const uint32_t d2 = r % Dc; r /= Dc;
                            const uint32_t d1 = r % Db; r /= Db;
                            const uint32_t d0 = r;
                            const int32_t I = S.cmin[t][0] + (int32_t)d0, J = S.cmin[t][1] + (int32_t)d1,
                                          K = S.cmin[t][2] + (int32_t)d2, L = S.cmin[t][3] + (int32_t)d3;
                            int32_t g = S.perm[L & 255];
                            g = S.perm[((K & 255) + g) & 255];
                            g = S.perm[((J & 255) + g) & 255];
                            g = S.perm[((I & 255) + g) & 255];
                            S.tab[e] = grad4_vector((seed ^ g) & 31);
                        }
                    }
                    __syncthreads();
                    // ---- per-type constants (one thread per type of the batch) ----
                    for (uint32_t t = t0 + tid; t < t1; t += TYPES_THREADS) {
                        const int32_t Dd = S.cmax[t][3] - S.cmin[t][3] + 2, Dc = S.cmax[t][2] - S.cmin[t][2] + 2,
                                      Db = S.cmax[t][1] - S.cmin[t][1] + 2;
                        const int32_t sc = 16 * Dd, sb = sc * Dc, sa = sb * Db;
                        // address of cell (0,0,0,0): |cmin| < 2^12 and strides <= 2^14 bytes keep every term below 2^26
                        const float base =
                            ((addr_as_float((int32_t)tab_base + 16 * (int32_t)S.tab_off[t]) -
                              (float)S.cmin[t][0] * addr_as_float(sa)) - (float)S.cmin[t][1] * addr_as_float(sb)) -
                            ((float)S.cmin[t][2] * addr_as_float(sc) + (float)S.cmin[t][3] * addr_as_float(16));
                        TypeTab2 T;
                        T.x = bc2(type_axis_coordinate(t) * ft);
                        T.base = bc2(base);
                        T.sa = bc2(addr_as_float(sa));
                        T.sb = bc2(addr_as_float(sb));
                        T.sc = bc2(addr_as_float(sc));
                        T.s4 = bc2(addr_as_float(sa + sb + sc + 16));
                        S.tt[t - t0] = T;
                    }
                    __syncthreads();
                    // ---- evaluate the batch's types over this thread's k-column: voxel pair outermost, types
                    // innermost, the running maximum of the pair in registers ----
                    const bool first_batch = t0 == 0, last_batch = t1 == n_types;
                    float yacc = lo.z;
#pragma unroll 1
                    for (int k = 0; k < 16; k += 2) {
                        const float yacc1 = yacc + 1.0f;
                        const f2 Y2 = make_float2(yacc * fn, yacc1 * fn);
                        const f2 YZW = add2(Y2, ZW2);
                        const int bi = k * TYPES_THREADS + tid;
                        f2 best;
                        uint32_t bt;
                        if (first_batch) {
                            best = make_float2(0.0f, 0.0f);
                            bt = 0u;
                        } else {
                            best = make_float2(S.best[bi], S.best[bi + TYPES_THREADS]);
                            bt = S.type_pair[(k >> 1) * TYPES_THREADS + tid];
                        }
#pragma unroll 1
                        for (uint32_t t = t0; t < t1; ++t) {
                            const f2 nv = simplex4_tab2(Y2, Z2, W2, YZW, S.tt[t - t0], nz);
                            // voxel_type.rs:154-165: the first type's noise starts the maximum, later ones need `>`
                            if (t == 0u || nv.x > best.x) {
                                best.x = nv.x;
                                bt = (bt & 0xFF00u) | t;
                            }
                            if (t == 0u || nv.y > best.y) {
                                best.y = nv.y;
                                bt = (bt & 0x00FFu) | (t << 8);
                            }
                        }
                        S.type_pair[(k >> 1) * TYPES_THREADS + tid] = (uint16_t)bt;
                        if (!last_batch) {
                            S.best[bi] = best.x;
                            S.best[bi + TYPES_THREADS] = best.y;
                        }
                        yacc = yacc1 + 1.0f;
                    }
                    __syncthreads();  // the table is rebuilt by the next batch
                    t0 = t1;
                }
#pragma unroll
                for (int k = 0; k < 16; k += 2) {
                    const uint32_t bt = S.type_pair[(k >> 1) * TYPES_THREADS + tid];
                    types[k] = (uint8_t)(bt & 0xFFu);
                    types[k + 1] = (uint8_t)(bt >> 8);
                }
            }
        }

        // uniform ⇔ every voxel is maximally inside with the same type (object.rs:1913-1918)
        if (all_m128) {
            if (tid == 0) S.first_type = types[0];
            __syncthreads();
            bool same_cta = true;
#pragma unroll
            for (int k = 0; k < 16; ++k) same_cta = same_cta && (types[k] == S.first_type);
            const int uniform = __syncthreads_and(same_cta);
            if (uniform) {
                if (tid == 0) {
                    cd.kind = 1;
                    cd.pre = PRE_UNIFORM;
                    cd.u_type = S.first_type;
                    cd.u_sd = -128;
                    cd.u_flags = 0xFC;
                    cd.flags = 0;
                    a.chunks[chunk] = cd;
                    if (a.occ) {
                        for (int d = 0; d < 3; ++d) {
                            atomicMin(&a.occ[d], org[d]);
                            atomicMax(&a.occ[3 + d], org[d] + 15u);
                        }
                    }
                }
                __syncthreads();
                continue;
            }
        }

        // ---- flags: IS_EMPTY + in-chunk adjacency (object.rs:2673-2756 on fresh voxels) ----
        uint8_t flags[16];
        {
            auto nonempty_at = [&](int i, int j, int k) -> bool { return S.sd[vidx(i, j, k)] < 0; };
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                uint8_t f = 0;
                if (codes[k] >= 0) {
                    f = 1;  // IS_EMPTY; empty voxels carry no in-chunk adjacency bits
                } else {
                    if (ti > 0 && nonempty_at(ti - 1, tj, k)) f |= 1u << 2;
                    if (tj > 0 && nonempty_at(ti, tj - 1, k)) f |= 1u << 3;
                    if (k > 0 && codes[k - 1] < 0) f |= 1u << 4;
                    if (ti < 15 && nonempty_at(ti + 1, tj, k)) f |= 1u << 5;
                    if (tj < 15 && nonempty_at(ti, tj + 1, k)) f |= 1u << 6;
                    if (k < 15 && codes[k + 1] < 0) f |= 1u << 7;
                }
                flags[k] = f;
            }
        }

        // ---- face empty counts → FaceVoxelDistribution (object.rs:1920-1936, 2967-2981) ----
        {
            const uint32_t ne = __popc(empty_mask);
            if (ti == 0) atomicAdd(&S.cnt[0], ne);
            if (ti == 15) atomicAdd(&S.cnt[1], ne);
            if (tj == 0) atomicAdd(&S.cnt[2], ne);
            if (tj == 15) atomicAdd(&S.cnt[3], ne);
            if (empty_mask & 1u) atomicAdd(&S.cnt[4], 1u);
            if (empty_mask & 0x8000u) atomicAdd(&S.cnt[5], 1u);
            // bounding range of non-empty voxels (object.rs:1187-1280)
            const uint32_t nonempty = (~empty_mask) & 0xFFFFu;
            if (nonempty) {
                atomicMin(&S.cnt[6], (uint32_t)ti);
                atomicMin(&S.cnt[7], (uint32_t)tj);
                atomicMin(&S.cnt[8], (uint32_t)(__ffs(nonempty) - 1));
                atomicMax(&S.cnt[9], (uint32_t)ti);
                atomicMax(&S.cnt[10], (uint32_t)tj);
                atomicMax(&S.cnt[11], (uint32_t)(31 - __clz(nonempty)));
            }
        }
        __syncthreads();

        // ---- store the type and flag planes: 16 B per thread per plane, fully coalesced ----
        {
            uint4 pt, pf;
            uint32_t* wt = reinterpret_cast<uint32_t*>(&pt);
            uint32_t* wf = reinterpret_cast<uint32_t*>(&pf);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                wt[q] = (uint32_t)types[4 * q] | ((uint32_t)types[4 * q + 1] << 8) | ((uint32_t)types[4 * q + 2] << 16) |
                        ((uint32_t)types[4 * q + 3] << 24);
                wf[q] = (uint32_t)flags[4 * q] | ((uint32_t)flags[4 * q + 1] << 8) | ((uint32_t)flags[4 * q + 2] << 16) |
                        ((uint32_t)flags[4 * q + 3] << 24);
            }
            *reinterpret_cast<uint4*>(slot_ptr + PLANE_TYPE + tid * 16) = pt;
            *reinterpret_cast<uint4*>(slot_ptr + PLANE_FLAGS + tid * 16) = pf;
        }
        if (tid == 0) {
            cd.kind = 2;
            cd.pre = PRE_ACTIVE;
            cd.slot = slot;
            if (!any_nonempty) {
                for (int q = 0; q < 6; ++q) cd.face[q] = 0;
                cd.flags = 1u << 6;  // HAS_ONLY_EMPTY_VOXELS
            } else {
                for (int q = 0; q < 6; ++q) cd.face[q] = S.cnt[q] == 256u ? 0 : (S.cnt[q] == 0u ? 1 : 2);
                cd.flags = 0;
                if (a.occ) {
                    for (int d = 0; d < 3; ++d) {
                        atomicMin(&a.occ[d], org[d] + S.cnt[6 + d]);
                        atomicMax(&a.occ[3 + d], org[d] + S.cnt[9 + d]);
                    }
                }
            }
            a.chunks[chunk] = cd;
        }
        __syncthreads();
    }
}

size_t types_smem_bytes() { return sizeof(TypesSmem); }

int types_max_blocks_per_sm() {
    int nb = 0;
    cudaFuncSetAttribute(k_types, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TypesSmem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_types, TYPES_THREADS, sizeof(TypesSmem));
    return nb;
}

cudaError_t launch_types(const TypesArgs& a, uint32_t grid, cudaStream_t st) {
    if (a.n_active == 0 || grid == 0) return cudaSuccess;
    cudaFuncSetAttribute(k_types, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TypesSmem));
    k_types<<<grid, TYPES_THREADS, sizeof(TypesSmem), st>>>(a);
    return cudaGetLastError();
}

}  // namespace ivx
