// Micro-benchmark (diagnosis only): results per clock per SM of the special-function unit for the approximations a fused
// SiLU can be built from. Every warp of a full-occupancy grid runs long chains of independent operations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/ubench/mufu_rate.cu -o /tmp/mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ void op(float& x, uint32_t& u) {
    if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x));
    if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
    if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x));
    if (OP == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x));
    if (OP == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u));
    if (OP == 5) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u));
    if (OP == 6) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u));
    if (OP == 7) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
    if (OP == 8) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x));
}

template <int OP>
__global__ void __launch_bounds__(256) rate_k(float* out, long long* cyc, int iters) {
    float x[8];
    uint32_t u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        x[j] = 0.001f * (threadIdx.x + j);
        u[j] = 0x3c003c00u + threadIdx.x + j;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) op<OP>(x[j], u[j]);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + __uint_as_float(u[j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
static void run(const char* name, int per_op) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaMalloc(&cyc, 8);
    const int iters = 2000, blocks_per_sm = 4;
    rate_k<OP><<<148 * blocks_per_sm, 256>>>(out, cyc, iters);
    rate_k<OP><<<148 * blocks_per_sm, 256>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double results = (double)blocks_per_sm * 256 * iters * 8 * per_op;
    printf("%-28s %8.2f results / clk / SM  (%lld cycles)\n", name, results / (double)h, h);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("tanh.approx.f32", 1);
    run<1>("ex2.approx.ftz.f32", 1);
    run<2>("rcp.approx.ftz.f32", 1);
    run<3>("rsqrt.approx.ftz.f32", 1);
    run<4>("tanh.approx.bf16x2", 2);
    run<5>("ex2.approx.ftz.bf16x2", 2);
    run<6>("tanh.approx.f16x2", 2);
    run<7>("ex2.approx.f16x2", 2);
    run<8>("fma.rn.f32", 1);
    return 0;
}
