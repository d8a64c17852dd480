// Microbenchmark: issue rate of packed f32x2 arithmetic (FFMA2 / FADD2 / FMUL2) against scalar FFMA on sm_100a,
// alone and mixed with ALU-pipe work. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, float negzero) {
    const float2 nz = make_float2(negzero, negzero);
    float2 a[CHAINS];
    float s[CHAINS * 2];
    unsigned u[CHAINS];
    for (int c = 0; c < CHAINS; ++c) {
        a[c] = make_float2(seed + c + threadIdx.x, seed - c);
        s[2 * c] = a[c].x; s[2 * c + 1] = a[c].y;
        u[c] = threadIdx.x * 7 + c;
    }
    const float2 m = make_float2(1.0000001f, 0.9999999f), b = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) {           // scalar FFMA: 2 per chain
                s[2 * c] = __fmaf_rn(s[2 * c], m.x, b.x);
                s[2 * c + 1] = __fmaf_rn(s[2 * c + 1], m.y, b.y);
            } else if (MODE == 1) {    // FFMA2: 1 per chain (same flops as mode 0)
                a[c] = __ffma2_rn(a[c], m, b);
            } else if (MODE == 2) {    // FMUL2 + FADD2
                a[c] = __fadd2_rn(__ffma2_rn(a[c], m, nz), b);  // exact product (runtime -0 addend: ptxas cannot contract), then add
            } else if (MODE == 3) {    // scalar FMUL + FADD (2+2 per chain)
                s[2 * c] = __fadd_rn(__fmul_rn(s[2 * c], m.x), b.x);
                s[2 * c + 1] = __fadd_rn(__fmul_rn(s[2 * c + 1], m.y), b.y);
            } else if (MODE == 4) {    // FFMA2 + one ALU op (LOP3) per chain
                a[c] = __ffma2_rn(a[c], m, b);
                u[c] = (u[c] ^ (u[c] >> 3)) & 0x7fffffffu;
            } else if (MODE == 6) {    // FADD2 only
                a[c] = __fadd2_rn(a[c], b);
            } else if (MODE == 7) {    // scalar FADD x2
                s[2 * c] = __fadd_rn(s[2 * c], b.x);
                s[2 * c + 1] = __fadd_rn(s[2 * c + 1], b.y);
            } else if (MODE == 5) {    // 2 scalar FFMA + one ALU op
                s[2 * c] = __fmaf_rn(s[2 * c], m.x, b.x);
                s[2 * c + 1] = __fmaf_rn(s[2 * c + 1], m.y, b.y);
                u[c] = (u[c] ^ (u[c] >> 3)) & 0x7fffffffu;
            }
        }
    }
    float r = 0;
    for (int c = 0; c < CHAINS; ++c) r += a[c].x + a[c].y + s[2 * c] + s[2 * c + 1] + (float)u[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* out, double flop_pairs_per_iter) {
    const int grid = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 1.0f, -0.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(out, 1.0f, -0.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double lane_ops = (double)grid * 256 * ITERS * CHAINS * flop_pairs_per_iter;  // scalar-equivalent f32 ops
    const double per_clk_sm = lane_ops / (ms * 1e-3) / 148 / 1.965e9;
    printf("%-34s %8.3f ms  %7.1f f32-ops/clk/SM (at 1965 MHz)\n", name, ms, per_clk_sm);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA x2", out, 2);
    run<1>("FFMA2", out, 2);
    run<2>("FFMA2(-0)+FADD2", out, 4);
    run<3>("scalar FMUL+FADD x2", out, 4);
    run<4>("FFMA2 + LOP-ish", out, 2);
    run<5>("scalar FFMA x2 + LOP-ish", out, 2);
    run<6>("FADD2", out, 2);
    run<7>("scalar FADD x2", out, 2);
    return 0;
}
