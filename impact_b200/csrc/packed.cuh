// Packed f32x2 arithmetic (sm_100a FADD2 / FFMA2) and the lattice-table helpers shared by the noise kernels.
#pragma once
#include "common.cuh"

namespace ivx {

constexpr float MAGIC = 12582912.0f;      // 1.5 * 2^23: (small integer + MAGIC) keeps the integer in the low mantissa bits
constexpr uint32_t MAGIC_BITS = 0x4B400000u;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// ---- two voxels per thread on the packed f32x2 pipe -------------------------------------------------
// sm_100a issues FADD2 / FFMA2 (two independent, individually rounded f32 operations per lane) at the
// scalar FP32 rate per *operation* but half the rate per *instruction*; k_types is bound by issue
// slots, not by the FMA pipe (tools/microbench/f32x2.cu, profiles/README.md), so the same arithmetic
// on (k, k+1) voxel pairs needs ~1/3 fewer slots. Every packed operation rounds exactly like the scalar
// one it replaces. ptxas 12.9 contracts `mul.rn.f32x2` followed by `add.rn.f32x2` into FFMA2 even
// under --fmad=false, which would change results; packed products therefore go through FFMA2 with an
// addend of -0.0 that is only known at run time (x·y + (-0) == RN(x·y) for every x·y, signed zeros
// included), so there is no multiply for ptxas to contract.
typedef float2 f2;
__device__ __forceinline__ f2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b, f2 nz) { return __ffma2_rn(a, b, nz); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 gt2(f2 a, f2 b) { return make_float2(a.x > b.x ? 1.0f : 0.0f, a.y > b.y ? 1.0f : 0.0f); }
__device__ __forceinline__ f2 gtc2(f2 a, float c) { return make_float2(a.x > c ? 1.0f : 0.0f, a.y > c ? 1.0f : 0.0f); }
__device__ __forceinline__ f2 floor2(f2 a) { return make_float2(floorf(a.x), floorf(a.y)); }
__device__ __forceinline__ f2 min2c(f2 a, float c) { return make_float2(fminf(a.x, c), fminf(a.y, c)); }
__device__ __forceinline__ f2 max2c(f2 a, float c) { return make_float2(fmaxf(a.x, c), fmaxf(a.y, c)); }
// entry address of a MAGIC-biased entry number. Inline PTX on purpose: nvcc 12.9 drops the `* 16` for the
// low half of a float2 returned by the f32x2 builtins when this is written in C++.
__device__ __forceinline__ float4 tab_load(float e, uint32_t addr) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(r) : "r"(__float_as_uint(e)), "r"(addr));
    return lds128(r);
}
__device__ __forceinline__ float gdot(const float4 g, float x, float y, float z, float w) {
    return __fmaf_rn(g.x, x, __fmaf_rn(g.y, y, __fmaf_rn(g.z, z, g.w * w)));
}


}  // namespace ivx
