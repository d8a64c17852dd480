// Microbenchmark: how the FMA pipe (packed FFMA2 / FADD2, scalar FFMA), the ALU pipe
// (FMNMX, FSET) and LDS share issue slots on sm_100a. Inline PTX fixes the instruction mix; 8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipemix pipemix.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 8;
typedef unsigned long long u64;
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, int warps_limit, int iters) {
    __shared__ float4 tab[1024];
    for (int q = threadIdx.x; q < 1024; q += 256) tab[q] = make_float4(q, 1, 2, 3);
    __syncthreads();
    if ((int)(threadIdx.x >> 5) >= warps_limit) return;  // warps_limit > 8: all 8 warps, several CTAs per SM (see run)
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(tab);
    u64 a[CH], b[CH], c[CH], d[CH];
    float s[CH], t[CH], u[CH], v[CH];
    float4 w[CH];
    for (int i = 0; i < CH; ++i) {
        float2 x = make_float2(seed + i + threadIdx.x, seed - i), y = make_float2(1.0000001f, 0.9999999f), z = make_float2(1e-7f * (i + 1), -1e-7f);
        a[i] = *reinterpret_cast<u64*>(&x); b[i] = *reinterpret_cast<u64*>(&y); c[i] = *reinterpret_cast<u64*>(&z); d[i] = a[i];
        s[i] = seed * i; t[i] = 0.999f; u[i] = seed + i; v[i] = seed - i; w[i] = make_float4(0, 0, 0, 0);
    }
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(c[(i+1)%CH]));
            }
            if (MODE == 1) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c[i]));
            }
            if (MODE == 2) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(t[i]), "f"(u[(i+1)%CH]));
            }
            if (MODE == 3) {
                asm volatile("min.f32 %0, %0, %1;" : "+f"(u[i]) : "f"(v[(i+1)%CH]));
            }
            if (MODE == 4) {
                asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(u[(i+2)%CH]));
            }
            if (MODE == 5) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c[i]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(u[i]) : "f"(v[(i+1)%CH]));
            }
            if (MODE == 6) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c[i]));
                asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(u[(i+2)%CH]));
            }
            if (MODE == 7) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c[i]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(c[i]) : "l"(d[(i+2)%CH]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(u[i]) : "f"(v[(i+1)%CH]));
            }
            if (MODE == 8) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(c[(i+1)%CH]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(u[i]) : "f"(v[(i+1)%CH]));
            }
            if (MODE == 9) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c[i]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(t[i]), "f"(u[(i+1)%CH]));
            }
            if (MODE == 10) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c[i]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(t[i]), "f"(u[(i+1)%CH]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(t[i]) : "f"(u[i]), "f"(s[(i+3)%CH]));
            }
            if (MODE == 11) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(t[i]), "f"(u[(i+1)%CH]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(u[i]) : "f"(v[(i+1)%CH]));
            }
            if (MODE == 12) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(t[i]), "f"(u[(i+1)%CH]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(t[i]) : "f"(u[i]), "f"(s[(i+3)%CH]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(u[i]) : "f"(v[(i+1)%CH]));
            }
            if (MODE == 13) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c[i]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(c[i]) : "l"(d[(i+2)%CH]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(c[(i+1)%CH]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(a[(i+5)%CH]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[i]) : "l"(c[i]), "l"(b[(i+3)%CH]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(t[i]), "f"(u[(i+1)%CH]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(t[i]) : "f"(u[i]), "f"(s[(i+3)%CH]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(u[i]) : "f"(v[(i+1)%CH]));
                asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(u[(i+2)%CH]));
            }
            if (MODE == 14) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c[i]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(c[i]) : "l"(d[(i+2)%CH]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(c[(i+1)%CH]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(a[(i+5)%CH]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[i]) : "l"(c[i]), "l"(b[(i+3)%CH]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(t[i]), "f"(u[(i+1)%CH]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(t[i]) : "f"(u[i]), "f"(s[(i+3)%CH]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(u[i]) : "f"(v[(i+1)%CH]));
                asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(u[(i+2)%CH]));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w[i].x), "=f"(w[i].y), "=f"(w[i].z), "=f"(w[i].w) : "r"(saddr + ((__float_as_uint(u[i]) & 0x3ffu) << 4)));
            }

        }
    }
    float r = 0;
    for (int i = 0; i < CH; ++i) {
        float2 x = *reinterpret_cast<float2*>(&a[i]), y = *reinterpret_cast<float2*>(&b[i]), z = *reinterpret_cast<float2*>(&c[i]), q = *reinterpret_cast<float2*>(&d[i]);
        r += x.x + x.y + y.x + y.y + z.x + z.y + q.x + q.y + s[i] + t[i] + u[i] + v[i] + w[i].x + w[i].y + w[i].z + w[i].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, float* out, int warps, int n_instr) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int ITERS = 1 << 18;
    k<MODE><<<148 * (warps > 8 ? warps / 8 : 1), 256>>>(out, 1.0f, warps, ITERS);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * (warps > 8 ? warps / 8 : 1), 256>>>(out, 1.0f, warps, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups = (warps / 4.0) * ITERS * CH;  // per sub-partition
    const double cyc = ms * 1e-3 * 1.965e9 / groups;
    printf("%-44s warps/SMSP %d  %8.4f ms  %6.2f cycles/group  (%d instr: IPC %.2f)\n", name, warps / 4, ms, cyc, n_instr, n_instr / cyc);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 256 * 4 * 4);
    k<2><<<148, 256>>>(out, 1.0f, 8, 1 << 22);  // warm the clocks up
    cudaDeviceSynchronize();
    for (int warps : {4, 8, 24}) {
        run<0>("FFMA2 (3 distinct pairs)", out, warps, 1);
        run<1>("FADD2", out, warps, 1);
        run<2>("FFMA", out, warps, 1);
        run<3>("FMNMX", out, warps, 1);
        run<4>("FSET", out, warps, 1);
        run<5>("FADD2 + FMNMX", out, warps, 2);
        run<6>("FADD2 + FSET", out, warps, 2);
        run<7>("2 FADD2 + FMNMX", out, warps, 3);
        run<8>("FFMA2 + FMNMX", out, warps, 2);
        run<9>("FADD2 + FFMA", out, warps, 2);
        run<10>("FADD2 + 2 FFMA", out, warps, 3);
        run<11>("FFMA + FMNMX", out, warps, 2);
        run<12>("2 FFMA + FMNMX", out, warps, 3);
        run<13>("k_types mix: 5 packed + 2 FFMA + 2 ALU", out, warps, 9);
        run<14>("k_types mix + LDS", out, warps, 10);

    }
    return 0;
}
