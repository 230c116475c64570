// HBM-bound kernels of the WFC-guidance path. See kernels.cuh for the contract of each entry.
#include "kernels.cuh"
#include "tapgemm.cuh"  // set_error, num_sms
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace wfd {

#define WFD_CHECK_LAUNCH(name)                                          \
    do {                                                                \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) {                                       \
            set_error("%s launch: %s", name, cudaGetErrorString(e__));  \
            return -7;                                                  \
        }                                                               \
        ++g_launch_count;                                               \
        if (debug_sync_enabled()) {                                     \
            int rc__ = debug_sync_check(name, s);                       \
            if (rc__ < 0) return rc__;                                  \
        }                                                               \
    } while (0)

#define WFD_DISPATCH_BF16(flag, ...)      \
    do {                                  \
        if (flag) {                       \
            constexpr bool BF16 = true;   \
            __VA_ARGS__;                  \
        } else {                          \
            constexpr bool BF16 = false;  \
            __VA_ARGS__;                  \
        }                                 \
    } while (0)

// ------------------------------------------------------------------------------------------------ 16-bit helpers
template <bool BF16>
__device__ __forceinline__ float u16_to_f(uint32_t v) {
    if (BF16) return __uint_as_float(v << 16);
    return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xFFFF)));
}
template <bool BF16>
__device__ __forceinline__ uint16_t f_to_u16(float f) {
    if (BF16) {
        __nv_bfloat16 h = __float2bfloat16_rn(f);
        return *reinterpret_cast<uint16_t*>(&h);
    }
    __half h = __float2half_rn(f);
    return *reinterpret_cast<uint16_t*>(&h);
}
template <bool BF16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (BF16) {
            f[2 * j] = __uint_as_float(w[j] << 16);
            f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
        } else {
            const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
            f[2 * j] = h.x;
            f[2 * j + 1] = h.y;
        }
    }
}
template <bool BF16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (BF16) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&h);
        } else {
            __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------ GroupNorm
constexpr int GN_THREADS = 384;
constexpr int GN_MAXG = 32;

struct GnMap {  // thread -> (vector column(s), row lane) mapping shared by the four GN kernels
    int nv, cpt, nvt, R;
};
__host__ __device__ inline GnMap gn_map(int C) {
    GnMap m;
    m.nv = C / 8;
    m.cpt = (m.nv + GN_THREADS - 1) / GN_THREADS;
    m.nvt = (m.nv + m.cpt - 1) / m.cpt;
    m.R = GN_THREADS / m.nvt;
    return m;
}
__device__ __forceinline__ unsigned long long to_fx(float v, float scale) {
    return static_cast<unsigned long long>(__float2ll_rn(v * scale));
}
__device__ __forceinline__ void gn_finalize(const acc_t* sums, int b, int g, int G, double n, float eps, float& mean,
                                            float& rstd) {
    const double s = static_cast<double>(sums[(b * G + g) * 2]) * (1.0 / WFD_FX_STATS);
    const double ss = static_cast<double>(sums[(b * G + g) * 2 + 1]) * (1.0 / WFD_FX_STATS);
    const double m = s / n;
    double var = ss / n - m * m;
    if (var < 0) var = 0;
    mean = static_cast<float>(m);
    rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

template <bool BF16>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_k(const uint4* __restrict__ x, acc_t* __restrict__ sums, int HW,
                                                          int C, int G, int ppb) {
    __shared__ unsigned long long sh[2 * GN_MAXG];
    const GnMap m = gn_map(C);
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
    const int tcol = threadIdx.x % m.nvt, trow = threadIdx.x / m.nvt;
    const int cpg = C / G;
    if (threadIdx.x < 2 * G) sh[threadIdx.x] = 0ull;
    __syncthreads();
    if (trow < m.R) {
        for (int k = 0; k < m.cpt; ++k) {
            const int col = tcol + k * m.nvt;
            if (col >= m.nv) break;
            // every 8-channel vector goes to fixed point on its own, so the sums do not depend on how pixels are split
            // over threads, blocks or row bands (integer addition is associative)
            unsigned long long is = 0ull, iss = 0ull;
            const uint4* px = x + (static_cast<long long>(b) * HW) * m.nv + col;
            constexpr int U = 4;
            for (int p = p0 + trow; p < p1; p += U * m.R) {
                uint4 raw[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (p + u * m.R < p1) raw[u] = __ldg(px + static_cast<long long>(p + u * m.R) * m.nv);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (p + u * m.R >= p1) break;
                    float f[8];
                    unpack8<BF16>(raw[u], f);
                    float s = 0.f, ss = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        s += f[j];
                        ss += f[j] * f[j];
                    }
                    is += to_fx(s, WFD_FX_STATS);
                    iss += to_fx(ss, WFD_FX_STATS);
                }
            }
            const int g = (col * 8) / cpg;
            atomicAdd(&sh[2 * g], is);
            atomicAdd(&sh[2 * g + 1], iss);
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * G)
        atomicAdd(reinterpret_cast<unsigned long long*>(sums) + b * G * 2 + threadIdx.x, sh[threadIdx.x]);
}

// sigmoid through the hardware tanh (one MUFU): 0.5 * tanh(0.5 u) + 0.5, abs error ~2.5e-4 — below the 16-bit output
// rounding of every consumer
__device__ __forceinline__ float sigmoidf_(float u) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
    return fmaf(0.5f, t, 0.5f);
}

template <bool BF16, int U>
__global__ void __launch_bounds__(GN_THREADS, U <= 2 ? 3 : 2)
gn_apply_k(const uint4* __restrict__ x, const acc_t* __restrict__ sums, const float* __restrict__ gamma,
           const float* __restrict__ beta, uint4* __restrict__ y, int HW, int HW_stat, int C, int G, float eps, int silu,
           int ppb) {
    const GnMap m = gn_map(C);
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
    const int tcol = threadIdx.x % m.nvt, trow = threadIdx.x / m.nvt;
    const int cpg = C / G;
    // mean / rstd come out of double-precision math on the integer sums: once per group and block, not once per thread
    __shared__ float s_mean[GN_MAXG], s_rstd[GN_MAXG];
    if (threadIdx.x < G)
        gn_finalize(sums, b, threadIdx.x, G, static_cast<double>(cpg) * HW_stat, eps, s_mean[threadIdx.x], s_rstd[threadIdx.x]);
    __syncthreads();
    if (trow >= m.R) return;
    for (int k = 0; k < m.cpt; ++k) {
        const int col = tcol + k * m.nvt;
        if (col >= m.nv) break;
        const float mean = s_mean[(col * 8) / cpg], rstd = s_rstd[(col * 8) / cpg];
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float ga = __ldg(gamma + col * 8 + j);
            sc[j] = rstd * ga;
            sh[j] = __ldg(beta + col * 8 + j) - mean * rstd * ga;
        }
        const long long base = (static_cast<long long>(b) * HW) * m.nv + col;
        // U independent 16-byte loads in flight per thread
        for (int p = p0 + trow; p < p1; p += U * m.R) {
            uint4 raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (p + u * m.R < p1) raw[u] = __ldg(x + base + static_cast<long long>(p + u * m.R) * m.nv);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (p + u * m.R >= p1) break;
                float f[8];
                unpack8<BF16>(raw[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v = f[j] * sc[j] + sh[j];
                    f[j] = silu ? v * sigmoidf_(v) : v;
                }
                y[base + static_cast<long long>(p + u * m.R) * m.nv] = pack8<BF16>(f);
            }
        }
    }
}

template <bool BF16, int U>
__global__ void __launch_bounds__(GN_THREADS, U <= 2 ? 3 : 2)
gn_bwd_reduce_k(const uint4* __restrict__ dy, const uint4* __restrict__ x, const acc_t* __restrict__ sums,
                const float* __restrict__ gamma, const float* __restrict__ beta, acc_t* __restrict__ red, int HW,
                int HW_stat, int C, int G, float eps, int silu, int ppb) {
    __shared__ unsigned long long sh[2 * GN_MAXG];
    const GnMap m = gn_map(C);
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
    const int tcol = threadIdx.x % m.nvt, trow = threadIdx.x / m.nvt;
    const int cpg = C / G;
    __shared__ float s_mean[GN_MAXG], s_rstd[GN_MAXG];
    if (threadIdx.x < 2 * G) sh[threadIdx.x] = 0ull;
    if (threadIdx.x < G)
        gn_finalize(sums, b, threadIdx.x, G, static_cast<double>(cpg) * HW_stat, eps, s_mean[threadIdx.x], s_rstd[threadIdx.x]);
    __syncthreads();
    if (trow < m.R) {
        for (int k = 0; k < m.cpt; ++k) {
            const int col = tcol + k * m.nvt;
            if (col >= m.nv) break;
            const int g = (col * 8) / cpg;
            const float mean = s_mean[g], rstd = s_rstd[g];
            float ga[8], be[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ga[j] = __ldg(gamma + col * 8 + j);
                be[j] = __ldg(beta + col * 8 + j);
            }
            // every 8-channel vector goes to fixed point on its own (as in gn_stats_k): the sums then do not depend on how the
            // pixels are split over threads, blocks or row bands, nor on the batch size that sizes the grid
            unsigned long long i1 = 0ull, i2 = 0ull;
            const long long base = (static_cast<long long>(b) * HW) * m.nv + col;
            for (int p = p0 + trow; p < p1; p += U * m.R) {
                uint4 rx[U], rd[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (p + u * m.R < p1) {
                        rx[u] = __ldg(x + base + static_cast<long long>(p + u * m.R) * m.nv);
                        rd[u] = __ldg(dy + base + static_cast<long long>(p + u * m.R) * m.nv);
                    }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (p + u * m.R >= p1) break;
                    float fx[8], fd[8];
                    unpack8<BF16>(rx[u], fx);
                    unpack8<BF16>(rd[u], fd);
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float xh = (fx[j] - mean) * rstd;
                        float du = fd[j];
                        if (silu) {
                            const float v = xh * ga[j] + be[j];
                            const float sg = sigmoidf_(v);
                            du *= sg * (1.f + v * (1.f - sg));
                        }
                        const float t = du * ga[j];
                        s1 += t;
                        s2 += t * xh;
                    }
                    i1 += to_fx(s1, WFD_FX_GRAD);
                    i2 += to_fx(s2, WFD_FX_GRAD);
                }
            }
            atomicAdd(&sh[2 * g], i1);
            atomicAdd(&sh[2 * g + 1], i2);
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * G)
        atomicAdd(reinterpret_cast<unsigned long long*>(red) + b * G * 2 + threadIdx.x, sh[threadIdx.x]);
}

template <bool BF16>
__global__ void __launch_bounds__(GN_THREADS, 2)
gn_bwd_apply_k(const uint4* __restrict__ dy, const uint4* __restrict__ x, const acc_t* __restrict__ sums,
               const float* __restrict__ gamma, const float* __restrict__ beta, const acc_t* __restrict__ red,
               const uint4* __restrict__ add, uint4* __restrict__ dx, int HW, int HW_stat, int C, int G, float eps,
               int silu, int ppb) {
    const GnMap m = gn_map(C);
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
    const int tcol = threadIdx.x % m.nvt, trow = threadIdx.x / m.nvt;
    const int cpg = C / G;
    __shared__ float s_mean[GN_MAXG], s_rstd[GN_MAXG], s_r1[GN_MAXG], s_r2[GN_MAXG];
    if (threadIdx.x < G) {
        const int g = threadIdx.x;
        const double n = static_cast<double>(cpg) * HW_stat;
        gn_finalize(sums, b, g, G, n, eps, s_mean[g], s_rstd[g]);
        s_r1[g] = static_cast<float>(static_cast<double>(red[(b * G + g) * 2]) * (1.0 / WFD_FX_GRAD) / n);
        s_r2[g] = static_cast<float>(static_cast<double>(red[(b * G + g) * 2 + 1]) * (1.0 / WFD_FX_GRAD) / n);
    }
    __syncthreads();
    if (trow >= m.R) return;
    for (int k = 0; k < m.cpt; ++k) {
        const int col = tcol + k * m.nvt;
        if (col >= m.nv) break;
        const int g = (col * 8) / cpg;
        const float mean = s_mean[g], rstd = s_rstd[g], r1 = s_r1[g], r2 = s_r2[g];
        float ga[8], be[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ga[j] = __ldg(gamma + col * 8 + j);
            be[j] = __ldg(beta + col * 8 + j);
        }
        const long long base = (static_cast<long long>(b) * HW) * m.nv + col;
        constexpr int U = 2;
        for (int p = p0 + trow; p < p1; p += U * m.R) {
            uint4 rx[U], rd[U], ra[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (p + u * m.R < p1) {
                    const long long idx = base + static_cast<long long>(p + u * m.R) * m.nv;
                    rx[u] = __ldg(x + idx);
                    rd[u] = __ldg(dy + idx);
                    if (add) ra[u] = __ldg(add + idx);
                }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (p + u * m.R >= p1) break;
                const long long idx = base + static_cast<long long>(p + u * m.R) * m.nv;
                float fx[8], fd[8], fa[8];
                unpack8<BF16>(rx[u], fx);
                unpack8<BF16>(rd[u], fd);
                if (add) unpack8<BF16>(ra[u], fa);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (fx[j] - mean) * rstd;
                    float du = fd[j];
                    if (silu) {
                        const float v = xh * ga[j] + be[j];
                        const float sg = sigmoidf_(v);
                        du *= sg * (1.f + v * (1.f - sg));
                    }
                    float v = rstd * (du * ga[j] - r1 - xh * r2);
                    if (add) v += fa[j];
                    fd[j] = v;
                }
                dx[idx] = pack8<BF16>(fd);
            }
        }
    }
}

static int gn_check(int C, int G) {
    if (C % 8 != 0 || G <= 0 || G > GN_MAXG || C % G != 0 || (C / G) % 8 != 0) {
        set_error("GroupNorm kernels need C %% 8 == 0 and (C/G) %% 8 == 0 (C=%d G=%d)", C, G);
        return -3;
    }
    return 0;
}
// One wave: the grid is sized to the number of blocks that are resident at once (blocks_per_sm comes from the occupancy
// API for the kernel at hand). A grid of 1.33 waves - what "4 per SM" gave for kernels that fit 3 - runs as long as 2.
template <typename K>
static int gn_blocks_per_sm(K kernel) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, GN_THREADS, 0) != cudaSuccess || n < 1) n = 2;
    return n;
}
static void gn_grid(int B, int HW, int C, int blocks_per_sm, dim3& grid, int& ppb) {
    const GnMap m = gn_map(C);
    int blocks_x = (num_sms() * blocks_per_sm) / B;
    if (blocks_x < 1) blocks_x = 1;
    int min_ppb = m.R * 8;
    ppb = (HW + blocks_x - 1) / blocks_x;
    if (ppb < min_ppb) ppb = min_ppb;
    if (ppb > HW) ppb = HW;
    blocks_x = (HW + ppb - 1) / ppb;
    grid = dim3(blocks_x, B);
}

int gn_stats(const void* x, acc_t* sums, int B, int HW, int C, int G, int bf16, cudaStream_t s) {
    ProfScope prof_scope("gn_stats", s);
    if (int r = gn_check(C, G)) return r;
    dim3 grid;
    int ppb;
    static const int bps = gn_blocks_per_sm(gn_stats_k<true>);
    gn_grid(B, HW, C, bps, grid, ppb);
    WFD_DISPATCH_BF16(bf16, (gn_stats_k<BF16><<<grid, GN_THREADS, 0, s>>>(static_cast<const uint4*>(x), sums, HW, C, G, ppb)));
    WFD_CHECK_LAUNCH("gn_stats");
    return 0;
}
int gn_apply(const void* x, const acc_t* sums, const float* gamma, const float* beta, void* y, int B, int HW, int C,
             int G, float eps, int silu, int bf16, cudaStream_t s, int HW_stat) {
    char prof_name[64];  // the profile keeps the shape: bench.py derives the bytes each launch moves from it
    snprintf(prof_name, sizeof(prof_name), "gn_apply:C=%d,HW=%d,B=%d", C, HW, B);
    ProfScope prof_scope(prof_name, s);
    if (int r = gn_check(C, G)) return r;
    dim3 grid;
    int ppb;
    // loads in flight per thread: WFD_GN_APPLY_U = 2 (3 blocks per SM), 4 or 6 (2 blocks per SM)
    static int u_sel = -1;
    if (u_sel < 0) {
        const char* e = getenv("WFD_GN_APPLY_U");
        u_sel = e ? atoi(e) : 4;  // measured on B200: 1.045 / 0.993 / 0.990 ms per step for 2 / 4 / 6
    }
    if (u_sel == 4) {
        static const int bps = gn_blocks_per_sm(gn_apply_k<true, 4>);
        gn_grid(B, HW, C, bps, grid, ppb);
        WFD_DISPATCH_BF16(bf16, (gn_apply_k<BF16, 4><<<grid, GN_THREADS, 0, s>>>(static_cast<const uint4*>(x), sums, gamma, beta,
                                                                                static_cast<uint4*>(y), HW, HW_stat > 0 ? HW_stat : HW, C, G, eps, silu, ppb)));
    } else if (u_sel == 6) {
        static const int bps = gn_blocks_per_sm(gn_apply_k<true, 6>);
        gn_grid(B, HW, C, bps, grid, ppb);
        WFD_DISPATCH_BF16(bf16, (gn_apply_k<BF16, 6><<<grid, GN_THREADS, 0, s>>>(static_cast<const uint4*>(x), sums, gamma, beta,
                                                                                static_cast<uint4*>(y), HW, HW_stat > 0 ? HW_stat : HW, C, G, eps, silu, ppb)));
    } else {
        static const int bps = gn_blocks_per_sm(gn_apply_k<true, 2>);
        gn_grid(B, HW, C, bps, grid, ppb);
        WFD_DISPATCH_BF16(bf16, (gn_apply_k<BF16, 2><<<grid, GN_THREADS, 0, s>>>(static_cast<const uint4*>(x), sums, gamma, beta,
                                                                                static_cast<uint4*>(y), HW, HW_stat > 0 ? HW_stat : HW, C, G, eps, silu, ppb)));
    }
    WFD_CHECK_LAUNCH("gn_apply");
    return 0;
}
int gn_bwd_reduce(const void* dy, const void* x, const acc_t* sums, const float* gamma, const float* beta, acc_t* red,
                  int B, int HW, int C, int G, float eps, int silu, int bf16, cudaStream_t s, int HW_stat) {
    char prof_name[64];
    snprintf(prof_name, sizeof(prof_name), "gn_bwd_reduce:C=%d,HW=%d,B=%d", C, HW, B);
    ProfScope prof_scope(prof_name, s);
    if (int r = gn_check(C, G)) return r;
    dim3 grid;
    int ppb;
    static int u_sel = -1;  // (x, dy) row pairs in flight per thread: WFD_GN_REDUCE_U = 2 (3 blocks per SM) or 4 (2 per SM)
    if (u_sel < 0) {
        const char* e = getenv("WFD_GN_REDUCE_U");
        u_sel = e ? atoi(e) : 4;  // measured on B200: 0.903 / 0.853 ms per step for 2 / 4
    }
    if (u_sel == 4) {
        static const int bps = gn_blocks_per_sm(gn_bwd_reduce_k<true, 4>);
        gn_grid(B, HW, C, bps, grid, ppb);
        WFD_DISPATCH_BF16(bf16, (gn_bwd_reduce_k<BF16, 4><<<grid, GN_THREADS, 0, s>>>(
                                    static_cast<const uint4*>(dy), static_cast<const uint4*>(x), sums, gamma, beta, red, HW,
                                    HW_stat > 0 ? HW_stat : HW, C, G, eps, silu, ppb)));
    } else {
        static const int bps = gn_blocks_per_sm(gn_bwd_reduce_k<true, 2>);
        gn_grid(B, HW, C, bps, grid, ppb);
        WFD_DISPATCH_BF16(bf16, (gn_bwd_reduce_k<BF16, 2><<<grid, GN_THREADS, 0, s>>>(
                                    static_cast<const uint4*>(dy), static_cast<const uint4*>(x), sums, gamma, beta, red, HW,
                                    HW_stat > 0 ? HW_stat : HW, C, G, eps, silu, ppb)));
    }
    WFD_CHECK_LAUNCH("gn_bwd_reduce");
    return 0;
}
int gn_bwd_apply(const void* dy, const void* x, const acc_t* sums, const float* gamma, const float* beta,
                 const acc_t* red, const void* add, void* dx, int B, int HW, int C, int G, float eps, int silu,
                 int bf16, cudaStream_t s, int HW_stat) {
    char prof_name[64];
    snprintf(prof_name, sizeof(prof_name), "gn_bwd_apply:C=%d,HW=%d,B=%d,add=%d", C, HW, B, add ? 1 : 0);
    ProfScope prof_scope(prof_name, s);
    if (int r = gn_check(C, G)) return r;
    dim3 grid;
    int ppb;
    static const int bps = gn_blocks_per_sm(gn_bwd_apply_k<true>);
    gn_grid(B, HW, C, bps, grid, ppb);
    WFD_DISPATCH_BF16(bf16, (gn_bwd_apply_k<BF16><<<grid, GN_THREADS, 0, s>>>(
                                static_cast<const uint4*>(dy), static_cast<const uint4*>(x), sums, gamma, beta, red,
                                static_cast<const uint4*>(add), static_cast<uint4*>(dx), HW, HW_stat > 0 ? HW_stat : HW,
                                C, G, eps, silu, ppb)));
    WFD_CHECK_LAUNCH("gn_bwd_apply");
    return 0;
}

// GroupNorm(+SiLU) backward: reduction (unless the producing contraction fused it) then apply. Statistics are per
// sample, so the batch can be walked in L2-sized blocks of samples (WFD_GN_L2_BLOCK_MB > 0) to make the second read of
// (dy, x) hit L2; measured on B200 the extra small launches cost more than the saved HBM reads (17.8 vs 16.5 ms per
// step at batch 8), so the default is one block.
int gn_bwd(const void* dy, const void* x, const acc_t* sums, const float* gamma, const float* beta, acc_t* red,
           const void* add, void* dx, int B, int HW, int C, int G, float eps, int silu, int bf16, int do_reduce,
           cudaStream_t s) {
    const size_t per_sample = static_cast<size_t>(HW) * C * 2;
    static int block_mb = -1;
    if (block_mb < 0) {
        const char* e = getenv("WFD_GN_L2_BLOCK_MB");
        block_mb = e ? atoi(e) : 0;
    }
    int nb = B;
    if (block_mb > 0) {
        nb = static_cast<int>((static_cast<size_t>(block_mb) << 20) / (2 * per_sample));
        if (nb < 1) nb = 1;
        if (nb > B) nb = B;
    }
    for (int b0 = 0; b0 < B; b0 += nb) {
        const int n = (B - b0 < nb) ? B - b0 : nb;
        const size_t off = static_cast<size_t>(b0) * per_sample;
        const uint8_t* dyb = static_cast<const uint8_t*>(dy) + off;
        const uint8_t* xb = static_cast<const uint8_t*>(x) + off;
        const uint8_t* addb = add ? static_cast<const uint8_t*>(add) + off : nullptr;
        uint8_t* dxb = static_cast<uint8_t*>(dx) + off;
        const acc_t* sb = sums + static_cast<size_t>(b0) * G * 2;
        acc_t* rb = red + static_cast<size_t>(b0) * G * 2;
        if (do_reduce) {
            int rc = gn_bwd_reduce(dyb, xb, sb, gamma, beta, rb, n, HW, C, G, eps, silu, bf16, s);
            if (rc < 0) return rc;
        }
        int rc = gn_bwd_apply(dyb, xb, sb, gamma, beta, rb, addb, dxb, n, HW, C, G, eps, silu, bf16, s);
        if (rc < 0) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ latent side
// One block per (row segment of LI_PIX pixels, sample). u = pq(z*in_scale) is staged with a one-pixel halo; a thread then
// owns FOUR output channels and LI_PIX / 2 pixels: per (tap, latent channel) one 16-byte weight load (transposed layout
// [ky][kx][ci][co]) feeds 4 x 8 accumulators, and every pixel leaves as one 8-byte store.
constexpr int LI_PIX = 16;
template <bool BF16>
__global__ void __launch_bounds__(256)
latent_in_k(const void* __restrict__ z, int z_f16, float in_scale, const float* __restrict__ pq_w,
            const float* __restrict__ pq_b, const float* __restrict__ wt, const float* __restrict__ bias,
            uint16_t* __restrict__ out, int h, int wd, int Cm, int y_off, int h_total) {
    // rows [y_off, y_off + h) of a z plane of h_total rows are produced (row bands); out is the local band
    __shared__ float u[3][LI_PIX + 2][4];
    const int b = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * LI_PIX;
    for (int i = threadIdx.x; i < 3 * (LI_PIX + 2); i += blockDim.x) {
        const int r = i / (LI_PIX + 2), c = i % (LI_PIX + 2);
        const int yy = y + y_off + r - 1, xx = x0 + c - 1;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (yy >= 0 && yy < h_total && xx >= 0 && xx < wd) {
            float zin[4];
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                const long long idx = ((static_cast<long long>(b) * 4 + ci) * h_total + yy) * wd + xx;
                zin[ci] = (z_f16 ? __half2float(reinterpret_cast<const __half*>(z)[idx])
                                 : reinterpret_cast<const float*>(z)[idx]) * in_scale;
            }
#pragma unroll
            for (int co = 0; co < 4; ++co) {
                float a = pq_b[co];
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) a += pq_w[co * 4 + ci] * zin[ci];
                v[co] = a;
            }
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) u[r][c][ci] = v[ci];
    }
    __syncthreads();
    const int quads = Cm / 4;
    constexpr int PH = LI_PIX / 2;
    for (int item = threadIdx.x; item < 2 * quads; item += blockDim.x) {
        const int q = item % quads, half = item / quads;
        const int co = q * 4, px0 = half * PH;
        const float4 bs = __ldg(reinterpret_cast<const float4*>(bias + co));
        float acc[PH][4];
#pragma unroll
        for (int p = 0; p < PH; ++p) {
            acc[p][0] = bs.x; acc[p][1] = bs.y; acc[p][2] = bs.z; acc[p][3] = bs.w;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(wt + ((ky * 3 + kx) * 4 + ci) * Cm + co));
#pragma unroll
                    for (int p = 0; p < PH; ++p) {
                        const float uv = u[ky][px0 + p + kx][ci];
                        acc[p][0] = fmaf(w4.x, uv, acc[p][0]);
                        acc[p][1] = fmaf(w4.y, uv, acc[p][1]);
                        acc[p][2] = fmaf(w4.z, uv, acc[p][2]);
                        acc[p][3] = fmaf(w4.w, uv, acc[p][3]);
                    }
                }
#pragma unroll
        for (int p = 0; p < PH; ++p) {
            if (x0 + px0 + p >= wd) break;
            uint2 o;
            o.x = static_cast<uint32_t>(f_to_u16<BF16>(acc[p][0])) | (static_cast<uint32_t>(f_to_u16<BF16>(acc[p][1])) << 16);
            o.y = static_cast<uint32_t>(f_to_u16<BF16>(acc[p][2])) | (static_cast<uint32_t>(f_to_u16<BF16>(acc[p][3])) << 16);
            *reinterpret_cast<uint2*>(out + ((static_cast<long long>(b) * h + y) * wd + x0 + px0 + p) * Cm + co) = o;
        }
    }
}

// `wt` is the transposed conv_in weight [ky][kx][ci][co] (the layout latent_out reads too)
int latent_in(const void* z, int z_f16, float in_scale, const float* pq_w, const float* pq_b, const float* wt,
              const float* b, void* out, int B, int h, int wd, int Cm, int bf16, cudaStream_t s, int y_off,
              int h_total) {
    ProfScope prof_scope("latent_in", s);
    if (h_total <= 0) h_total = h;
    if (Cm % 4) {
        set_error("latent_in: channel count %d must be a multiple of 4", Cm);
        return -3;
    }
    dim3 grid((wd + LI_PIX - 1) / LI_PIX, h, B);
    WFD_DISPATCH_BF16(bf16, (latent_in_k<BF16><<<grid, 256, 0, s>>>(z, z_f16, in_scale, pq_w, pq_b, wt, b,
                                                                   static_cast<uint16_t*>(out), h, wd, Cm, y_off, h_total)));
    WFD_CHECK_LAUNCH("latent_in");
    return 0;
}

// conv_in^T + post_quant_conv^T: one warp per latent pixel, the nine incoming-gradient rows read as 16-byte vectors (8
// channels per lane), the transposed conv_in weights [tap][ci][co] staged once per block in shared memory; blocks walk the
// pixels grid-stride. d_u[ci] = sum_{tap,co} dy[q - off(tap), co] * w[co, ci, tap]; dz = in_scale * pq_w^T d_u.
template <bool BF16>
__global__ void __launch_bounds__(256)
latent_out_k(const uint16_t* __restrict__ dy, float in_scale, const float* __restrict__ scale_b,
             const float* __restrict__ pq_w, const float* __restrict__ wt, float* __restrict__ dz, int B, int h, int wd,
             int Cm, int y_off, int h_total) {
    // row bands: dy has valid halo rows -1 and h wherever those rows are inside the h_total-row plane
    extern __shared__ float w_s[];  // [9][4][Cm]
    for (int i = threadIdx.x; i < 36 * Cm; i += blockDim.x) w_s[i] = __ldg(wt + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int total = B * h * wd;
    const int nv = Cm / 8;
    for (int pix = blockIdx.x * wpb + wib; pix < total; pix += gridDim.x * wpb) {
        const int b = pix / (h * wd);
        const int pr = pix - b * h * wd;
        const int y = pr / wd, x = pr - y * wd;
        float du[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ky = 0; ky < 3; ++ky) {
            const int ys = y - (ky - 1);
            if (ys + y_off < 0 || ys + y_off >= h_total) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int xs = x - (kx - 1);
                if (xs < 0 || xs >= wd) continue;
                const uint4* row = reinterpret_cast<const uint4*>(dy + ((static_cast<long long>(b) * h + ys) * wd + xs) * Cm);
                const float* wk = w_s + (ky * 3 + kx) * 4 * Cm;
                for (int v = lane; v < nv; v += 32) {
                    float g[8];
                    unpack8<BF16>(__ldg(row + v), g);
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const float4 wa = *reinterpret_cast<const float4*>(wk + ci * Cm + v * 8);
                        const float4 wb = *reinterpret_cast<const float4*>(wk + ci * Cm + v * 8 + 4);
                        du[ci] += g[0] * wa.x + g[1] * wa.y + g[2] * wa.z + g[3] * wa.w + g[4] * wb.x + g[5] * wb.y +
                                  g[6] * wb.z + g[7] * wb.w;
                    }
                }
            }
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) du[ci] = warp_sum(du[ci]);
        if (lane < 4) {
            float a = 0.f;
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) a += pq_w[ci * 4 + lane] * du[ci];
            a *= in_scale;
            if (scale_b) a *= scale_b[b];
            dz[((static_cast<long long>(b) * 4 + lane) * h + y) * wd + x] = a;
        }
    }
}

int latent_out(const void* dy, float in_scale, const float* scale_b, const float* pq_w, const float* w, float* dz,
               int B, int h, int wd, int Cm, int bf16, cudaStream_t s, int y_off, int h_total) {
    ProfScope prof_scope("latent_out", s);
    if (h_total <= 0) h_total = h;
    if (Cm % 8 || 36 * static_cast<size_t>(Cm) * sizeof(float) > 200 * 1024) {
        set_error("latent_out: channel count %d must be a multiple of 8 and its weights must fit shared memory", Cm);
        return -3;
    }
    const size_t smem = 36 * static_cast<size_t>(Cm) * sizeof(float);
    const long long warps = static_cast<long long>(B) * h * wd;
    const int threads = 256;
    long long blocks = (warps + threads / 32 - 1) / (threads / 32);
    const long long cap = static_cast<long long>(num_sms()) * 2;   // two resident blocks per SM share the staged weights
    if (blocks > cap) blocks = cap;
    static bool attr_done[2] = {false, false};
    if (!attr_done[bf16 ? 1 : 0]) {
        cudaError_t e = bf16 ? cudaFuncSetAttribute(latent_out_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                             : cudaFuncSetAttribute(latent_out_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            set_error("latent_out: %s", cudaGetErrorString(e));
            return -8;
        }
        attr_done[bf16 ? 1 : 0] = true;
    }
    WFD_DISPATCH_BF16(bf16, (latent_out_k<BF16><<<static_cast<unsigned>(blocks), threads, smem, s>>>(
                                static_cast<const uint16_t*>(dy), in_scale, scale_b, pq_w, w, dz, B, h, wd, Cm, y_off, h_total)));
    WFD_CHECK_LAUNCH("latent_out");
    return 0;
}

// Second half of conv_in^T when its contraction over the Cm channels ran on the tensor cores (decoder.cu): `du` holds
// [B*h*wd][ld] fp32 with column tap*4 + ci = sum_co dy[pixel, co] * w[co, ci, tap]; a latent pixel q gathers the nine taps
// from its neighbours q - off(tap), applies post_quant_conv^T and the two scales. One thread per latent pixel.
__global__ void __launch_bounds__(256)
latent_gather_k(const float* __restrict__ du, int ld, float in_scale, const float* __restrict__ scale_b,
                const float* __restrict__ pq_w, float* __restrict__ dz, int B, int h, int wd) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= B * h * wd) return;
    const int b = pix / (h * wd);
    const int pr = pix - b * h * wd;
    const int y = pr / wd, x = pr - y * wd;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // same order of the nine taps as latent_out_k
    for (int ky = 0; ky < 3; ++ky) {
        const int ys = y - (ky - 1);
        if (ys < 0 || ys >= h) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int xs = x - (kx - 1);
            if (xs < 0 || xs >= wd) continue;
            const float4 v = __ldg(reinterpret_cast<const float4*>(du + (static_cast<long long>(b) * h * wd + ys * wd + xs) * ld) + (ky * 3 + kx));
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
        }
    }
    const float sc = in_scale * (scale_b ? scale_b[b] : 1.f);
#pragma unroll
    for (int cl = 0; cl < 4; ++cl) {
        float a = 0.f;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) a += pq_w[ci * 4 + cl] * acc[ci];
        dz[((static_cast<long long>(b) * 4 + cl) * h + y) * wd + x] = a * sc;
    }
}
int latent_gather(const float* du, int ld, float in_scale, const float* scale_b, const float* pq_w, float* dz, int B, int h,
                  int wd, cudaStream_t s) {
    ProfScope prof_scope("latent_gather", s);
    const int total = B * h * wd;
    latent_gather_k<<<(total + 255) / 256, 256, 0, s>>>(du, ld, in_scale, scale_b, pq_w, dz, B, h, wd);
    WFD_CHECK_LAUNCH("latent_gather");
    return 0;
}

// ------------------------------------------------------------------------------------------------ wave softmax
constexpr int WS_MAXV = 20;  // vectors of 8 per lane -> rows up to 5120 channels stay in registers
template <bool BF16, bool SOFTMAX>
__global__ void __launch_bounds__(256)
wave_softmax_k(const uint4* __restrict__ logits, uint4* __restrict__ wave, float* __restrict__ rowsum, long long rows,
               int T_pad, int T) {
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = T_pad / 8;
    const uint4* src = logits + row * nv;
    uint4 raw[WS_MAXV];
#pragma unroll
    for (int k = 0; k < WS_MAXV; ++k) {
        const int v = lane + 32 * k;
        if (v < nv) raw[k] = __ldg(src + v);
    }
    float inv = 1.f, mx = 0.f;
    if (SOFTMAX) {
        mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < WS_MAXV; ++k) {
            if (lane + 32 * k < nv) {
                float f[8];
                unpack8<BF16>(raw[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
            }
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < WS_MAXV; ++k) {
            if (lane + 32 * k < nv) {
                float f[8];
                unpack8<BF16>(raw[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) sum += __expf(f[j] - mx);
            }
        }
        sum = warp_sum(sum);
        inv = 1.f / sum;
    }
    float st = 0.f;
#pragma unroll
    for (int k = 0; k < WS_MAXV; ++k) {
        const int v = lane + 32 * k;
        if (v < nv) {
            float f[8];
            unpack8<BF16>(raw[k], f);
            uint4 o = raw[k];
            if (SOFTMAX) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __expf(f[j] - mx) * inv;
                o = pack8<BF16>(f);
                wave[row * nv + v] = o;
                unpack8<BF16>(o, f);  // sum the rounded values, the ones the contraction will read
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (v * 8 + j < T) st += f[j];
        }
    }
    st = warp_sum(st);
    if (lane == 0) rowsum[row] = st;
}

static int wave_rows_launch(const void* in, void* wave, float* rowsum, long long rows, int T_pad, int T, int bf16,
                            bool softmax, cudaStream_t s) {
    ProfScope prof_scope("wave_rows_launch", s);
    if (T_pad % 8 != 0 || T_pad / 8 > 32 * WS_MAXV || T > T_pad) {
        set_error("wave softmax: T_pad=%d must be a multiple of 8 and <= %d", T_pad, 8 * 32 * WS_MAXV);
        return -3;
    }
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((rows * 32 + threads - 1) / threads);
    if (softmax)
        WFD_DISPATCH_BF16(bf16, (wave_softmax_k<BF16, true><<<blocks, threads, 0, s>>>(
                                    static_cast<const uint4*>(in), static_cast<uint4*>(wave), rowsum, rows, T_pad, T)));
    else
        WFD_DISPATCH_BF16(bf16, (wave_softmax_k<BF16, false><<<blocks, threads, 0, s>>>(
                                    static_cast<const uint4*>(in), static_cast<uint4*>(wave), rowsum, rows, T_pad, T)));
    WFD_CHECK_LAUNCH("wave_softmax");
    return 0;
}
int wave_softmax(const void* logits, void* wave, float* rowsum, long long rows, int T_pad, int T, int bf16,
                 cudaStream_t s) {
    return wave_rows_launch(logits, wave, rowsum, rows, T_pad, T, bf16, true, s);
}
int wave_rowsum(const void* wave, float* rowsum, long long rows, int T_pad, int T, int bf16, cudaStream_t s) {
    return wave_rows_launch(wave, nullptr, rowsum, rows, T_pad, T, bf16, false, s);
}

// ------------------------------------------------------------------------------------------------ WFC finishing passes
// row bands: S points at the band's first row and has valid halo rows -1 and H_loc wherever they are inside the map
__device__ __forceinline__ float neighbour_sum(const float* __restrict__ S, int y, int x, int W, int y_off,
                                               int H_total) {
    float c = 0.f;
    if (x + 1 < W) c += S[y * W + x + 1];
    if (x > 0) c += S[y * W + x - 1];
    if (y + y_off + 1 < H_total) c += S[(y + 1) * W + x];
    if (y + y_off > 0) c += S[(y - 1) * W + x];
    return c;
}

__global__ void __launch_bounds__(1024)
wfc_loss_finish_k(const float* __restrict__ rowsum, const acc_t* __restrict__ rowdot, float* __restrict__ loss, int H,
                  int W, float coef, int y_off, int H_total) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    const float* S = rowsum + static_cast<long long>(b) * H * W;
    const acc_t* D = rowdot + static_cast<long long>(b) * H * W;
    double acc = 0.0;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int y = i / W, x = i - y * W;
        acc += static_cast<double>(S[i]) * neighbour_sum(S, y, x, W, y_off, H_total) -
               static_cast<double>(D[i]) * (1.0 / WFD_FX_GRAD);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) loss[b] = static_cast<float>(0.5 * coef * v);
    }
}
int wfc_loss_finish(const float* rowsum, const acc_t* rowdot, float* loss, int B, int H, int W, float coef,
                    cudaStream_t s, int y_off, int H_total) {
    ProfScope prof_scope("wfc_loss_finish", s);
    if (H_total <= 0) H_total = H;
    wfc_loss_finish_k<<<B, 1024, 0, s>>>(rowsum, rowdot, loss, H, W, coef, y_off, H_total);
    WFD_CHECK_LAUNCH("wfc_loss_finish");
    return 0;
}

template <bool BF16>
__global__ void __launch_bounds__(256)
wfc_grad_finish_k(const uint4* __restrict__ wave, const uint4* __restrict__ a, const float* __restrict__ rowsum,
                  const acc_t* __restrict__ rowdot, const float* __restrict__ dloss, uint4* __restrict__ dout, int B,
                  int H, int W, int T_pad, int T, float coef, int through_softmax, int y_off, int H_total) {
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long rows = static_cast<long long>(B) * H * W;
    if (row >= rows) return;
    const int b = static_cast<int>(row / (H * W));
    const int pr = static_cast<int>(row - static_cast<long long>(b) * H * W);
    const int y = pr / W, x = pr - y * W;
    const float* S = rowsum + static_cast<long long>(b) * H * W;
    const float c = neighbour_sum(S, y, x, W, y_off, H_total);
    const float r = S[pr] * c - static_cast<float>(static_cast<double>(rowdot[row]) * (1.0 / WFD_FX_GRAD));
    const float sc = coef * (dloss ? dloss[b] : 1.f);
    const int nv = T_pad / 8;
    for (int v = lane; v < nv; v += 32) {
        float fa[8], fp[8];
        unpack8<BF16>(__ldg(a + row * nv + v), fa);
        if (through_softmax) unpack8<BF16>(__ldg(wave + row * nv + v), fp);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float g = (v * 8 + j < T ? c : 0.f) - fa[j];
            fa[j] = through_softmax ? sc * fp[j] * (g - r) : sc * g;
        }
        dout[row * nv + v] = pack8<BF16>(fa);
    }
}
int wfc_grad_finish(const void* wave, const void* a, const float* rowsum, const acc_t* rowdot, const float* dloss,
                    void* dout, int B, int H, int W, int T_pad, int T, float coef, int through_softmax, int bf16,
                    cudaStream_t s, int y_off, int H_total) {
    ProfScope prof_scope("wfc_grad_finish", s);
    if (H_total <= 0) H_total = H;
    const long long rows = static_cast<long long>(B) * H * W;
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((rows * 32 + threads - 1) / threads);
    WFD_DISPATCH_BF16(bf16, (wfc_grad_finish_k<BF16><<<blocks, threads, 0, s>>>(
                                static_cast<const uint4*>(wave), static_cast<const uint4*>(a), rowsum, rowdot, dloss,
                                static_cast<uint4*>(dout), B, H, W, T_pad, T, coef, through_softmax, y_off, H_total)));
    WFD_CHECK_LAUNCH("wfc_grad_finish");
    return 0;
}

// ------------------------------------------------------------------------------------------------ attention helpers
template <bool BF16>
__global__ void __launch_bounds__(256)
attn_softmax_k(const float* __restrict__ scores, uint16_t* __restrict__ probs, int n) {
    __shared__ float red[2][8];
    const long long row = blockIdx.x;
    const float4* src = reinterpret_cast<const float4*>(scores + row * n);
    const int nv = n / 4;
    float m = -INFINITY, l = 0.f;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        const float4 f = __ldg(src + v);
        const float vm = fmaxf(fmaxf(f.x, f.y), fmaxf(f.z, f.w));
        if (vm > m) {
            l *= __expf(m - vm);
            m = vm;
        }
        l += __expf(f.x - m) + __expf(f.y - m) + __expf(f.z - m) + __expf(f.w - m);
    }
    const float wm = warp_max(m);
    l *= (m == -INFINITY) ? 0.f : __expf(m - wm);
    l = warp_sum(l);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = wm;
        red[1][warp] = l;
    }
    __syncthreads();
    float M = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) M = fmaxf(M, red[0][i]);
    float L = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) L += red[1][i] * ((red[0][i] == -INFINITY) ? 0.f : __expf(red[0][i] - M));
    const float inv = 1.f / L;
    uint2* dst = reinterpret_cast<uint2*>(probs + row * n);
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        const float4 f = __ldg(src + v);
        uint2 o;
        o.x = static_cast<uint32_t>(f_to_u16<BF16>(__expf(f.x - M) * inv)) |
              (static_cast<uint32_t>(f_to_u16<BF16>(__expf(f.y - M) * inv)) << 16);
        o.y = static_cast<uint32_t>(f_to_u16<BF16>(__expf(f.z - M) * inv)) |
              (static_cast<uint32_t>(f_to_u16<BF16>(__expf(f.w - M) * inv)) << 16);
        dst[v] = o;
    }
}
int attn_softmax(const float* scores, void* probs, long long rows, int n, int bf16, cudaStream_t s) {
    ProfScope prof_scope("attn_softmax", s);
    if (n % 4 != 0) {
        set_error("attn_softmax: n=%d must be a multiple of 4", n);
        return -3;
    }
    WFD_DISPATCH_BF16(bf16, (attn_softmax_k<BF16><<<static_cast<unsigned>(rows), 256, 0, s>>>(
                                scores, static_cast<uint16_t*>(probs), n)));
    WFD_CHECK_LAUNCH("attn_softmax");
    return 0;
}

template <bool BF16>
__global__ void __launch_bounds__(256)
attn_softmax_bwd_k(const uint16_t* __restrict__ probs, const float* __restrict__ dp, uint16_t* __restrict__ ds, int n,
                   float scale) {
    __shared__ float red[8];
    const long long row = blockIdx.x;
    const float4* dsrc = reinterpret_cast<const float4*>(dp + row * n);
    const uint2* psrc = reinterpret_cast<const uint2*>(probs + row * n);
    const int nv = n / 4;
    float acc = 0.f;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        const float4 d = __ldg(dsrc + v);
        const uint2 p = __ldg(psrc + v);
        acc += d.x * u16_to_f<BF16>(p.x & 0xFFFF) + d.y * u16_to_f<BF16>(p.x >> 16) +
               d.z * u16_to_f<BF16>(p.y & 0xFFFF) + d.w * u16_to_f<BF16>(p.y >> 16);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    uint2* dst = reinterpret_cast<uint2*>(ds + row * n);
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        const float4 d = __ldg(dsrc + v);
        const uint2 p = __ldg(psrc + v);
        const float o0 = u16_to_f<BF16>(p.x & 0xFFFF) * (d.x - tot) * scale;
        const float o1 = u16_to_f<BF16>(p.x >> 16) * (d.y - tot) * scale;
        const float o2 = u16_to_f<BF16>(p.y & 0xFFFF) * (d.z - tot) * scale;
        const float o3 = u16_to_f<BF16>(p.y >> 16) * (d.w - tot) * scale;
        uint2 o;
        o.x = static_cast<uint32_t>(f_to_u16<BF16>(o0)) | (static_cast<uint32_t>(f_to_u16<BF16>(o1)) << 16);
        o.y = static_cast<uint32_t>(f_to_u16<BF16>(o2)) | (static_cast<uint32_t>(f_to_u16<BF16>(o3)) << 16);
        dst[v] = o;
    }
}
int attn_softmax_bwd(const void* probs, const float* dp, void* ds, long long rows, int n, float scale, int bf16,
                     cudaStream_t s) {
    ProfScope prof_scope("attn_softmax_bwd", s);
    if (n % 4 != 0) {
        set_error("attn_softmax_bwd: n=%d must be a multiple of 4", n);
        return -3;
    }
    WFD_DISPATCH_BF16(bf16, (attn_softmax_bwd_k<BF16><<<static_cast<unsigned>(rows), 256, 0, s>>>(
                                static_cast<const uint16_t*>(probs), dp, static_cast<uint16_t*>(ds), n, scale)));
    WFD_CHECK_LAUNCH("attn_softmax_bwd");
    return 0;
}

// Row softmax, between the two passes of the score product: the {max, sum exp} partials of a row (one per N tile and
// epilogue half, empty ones are {-inf, 0}) are merged in index order into {row max, 1 / row sum}. One thread per row.
__global__ void __launch_bounds__(256)
rowstats_combine_k(const float2* __restrict__ part, float2* __restrict__ out, long long rows, int n_part) {
    const long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float2* pr = part + row * n_part;
    float M = -INFINITY;
    for (int i = 0; i < n_part; ++i) M = fmaxf(M, __ldg(pr + i).x);
    float L = 0.f;
    for (int i = 0; i < n_part; ++i) {
        const float2 v = __ldg(pr + i);
        if (v.y > 0.f) L += v.y * __expf(v.x - M);
    }
    out[row] = make_float2(M, 1.f / L);
}
int rowstats_combine(const void* part, void* out, long long rows, int n_part, cudaStream_t s) {
    ProfScope prof_scope("rowstats_combine", s);
    rowstats_combine_k<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, s>>>(static_cast<const float2*>(part),
                                                                               static_cast<float2*>(out), rows, n_part);
    WFD_CHECK_LAUNCH("rowstats_combine");
    return 0;
}

// D[row] = sum_c a[row, c] * b[row, c] in fp32 over two 16-bit [rows, C] matrices (C a multiple of 8): one warp per row.
// Softmax backward needs sum_n dP[row, n] P[row, n] = sum_c dO[row, c] O[row, c] (O = P V, dP = dO V^T).
template <bool BF16>
__global__ void __launch_bounds__(256)
rowdot16_k(const uint4* __restrict__ a, const uint4* __restrict__ b, float* __restrict__ d, long long rows, int nv) {
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int v = lane; v < nv; v += 32) {
        float fa[8], fb[8];
        unpack8<BF16>(__ldg(a + row * nv + v), fa);
        unpack8<BF16>(__ldg(b + row * nv + v), fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(fa[j], fb[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) d[row] = acc;
}
int rowdot16(const void* a, const void* b, float* d, long long rows, int C, int bf16, cudaStream_t s) {
    ProfScope prof_scope("rowdot16", s);
    if (C % 8) {
        set_error("rowdot16: C=%d must be a multiple of 8", C);
        return -3;
    }
    const unsigned blocks = static_cast<unsigned>((rows * 32 + 255) / 256);
    WFD_DISPATCH_BF16(bf16, (rowdot16_k<BF16><<<blocks, 256, 0, s>>>(static_cast<const uint4*>(a), static_cast<const uint4*>(b), d,
                                                                     rows, C / 8)));
    WFD_CHECK_LAUNCH("rowdot16");
    return 0;
}

__global__ void __launch_bounds__(256)
transpose16_k(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int R, int C, long long ld_in,
              long long zstride_in) {
    __shared__ uint16_t tile[64][66];
    const int z = blockIdx.z;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const uint16_t* src = in + z * zstride_in;
    uint16_t* dst = out + static_cast<long long>(z) * R * C;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
    for (int i = ty; i < 64; i += 4) {
        const int r = r0 + i, c = c0 + tx;
        if (r < R && c < C) tile[i][tx] = src[static_cast<long long>(r) * ld_in + c];
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 4) {
        const int c = c0 + i, r = r0 + tx;
        if (r < R && c < C) dst[static_cast<long long>(c) * R + r] = tile[tx][i];
    }
}
// 64x64 tiles moved as 32-bit pairs: coalesced 128-byte rows on both sides, two 16-bit halves recombined from two
// shared-memory rows on the way out. Needs R, C multiples of 64 and 4-byte aligned rows.
__global__ void __launch_bounds__(256)
transpose16_pair_k(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int R, int C, long long ld_in32,
                   long long zstride_in32) {
    __shared__ uint32_t tile[64][33];
    const int z = blockIdx.z;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const uint32_t* src = in + z * zstride_in32;
    uint32_t* dst = out + static_cast<long long>(z) * C * (R / 2);  // out is [C][R] 16-bit = [C][R/2] 32-bit
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = w * 8 + i;
        tile[r][lane] = __ldg(src + static_cast<long long>(r0 + r) * ld_in32 + c0 / 2 + lane);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = w * 8 + i;  // input column = output row
        const uint32_t a = tile[2 * lane][c >> 1], b = tile[2 * lane + 1][c >> 1];
        const uint32_t v = (c & 1) ? ((a >> 16) | (b & 0xFFFF0000u)) : ((a & 0xFFFFu) | (b << 16));
        dst[static_cast<long long>(c0 + c) * (R / 2) + r0 / 2 + lane] = v;
    }
}
int transpose16(const void* in, void* out, int Z, int R, int C, long long ld_in, long long zstride_in, cudaStream_t s) {
    ProfScope prof_scope("transpose16", s);
    if (R % 64 == 0 && C % 64 == 0 && ld_in % 2 == 0 && zstride_in % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0) {
        dim3 grid2(C / 64, R / 64, Z);
        transpose16_pair_k<<<grid2, 256, 0, s>>>(static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), R, C,
                                                 ld_in / 2, zstride_in / 2);
        WFD_CHECK_LAUNCH("transpose16");
        return 0;
    }
    dim3 grid((C + 63) / 64, (R + 63) / 64, Z);
    transpose16_k<<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), R, C, ld_in,
                                       zstride_in);
    WFD_CHECK_LAUNCH("transpose16");
    return 0;
}

// ------------------------------------------------------------------------------------------------ argmax
struct Best {
    float v;
    int i;
    bool nan;
};
__device__ __forceinline__ void best_take(Best& b, float v, int i) {
    // scanned in increasing index order: strict '>' keeps the first maximum, the first NaN wins over everything
    if (b.nan) return;
    if (v != v) {
        b.v = v; b.i = i; b.nan = true;
    } else if (b.i < 0 || v > b.v) {
        b.v = v; b.i = i;
    }
}
__device__ __forceinline__ bool best_better(const Best& a, const Best& b) {  // is a better than b
    if (a.i < 0) return false;
    if (b.i < 0) return true;
    if (a.nan != b.nan) return a.nan;
    if (a.nan) return a.i < b.i;
    if (a.v != b.v) return a.v > b.v;
    return a.i < b.i;
}
template <int DT>
__global__ void __launch_bounds__(256) argmax_rows_k(const void* __restrict__ x, long long rows, int n,
                                                     long long* __restrict__ out) {
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    Best b;
    b.v = 0.f; b.i = -1; b.nan = false;
    const bool vec_ok = (n % (DT == 0 ? 4 : 8) == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (DT == 0) {
        const float* src = reinterpret_cast<const float*>(x) + row * n;
        const int nv = vec_ok ? n / 4 : 0;  // rows that do not start 16-byte aligned take the scalar loop
        for (int v = lane; v < nv; v += 32) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(src) + v);
            best_take(b, f.x, v * 4); best_take(b, f.y, v * 4 + 1); best_take(b, f.z, v * 4 + 2); best_take(b, f.w, v * 4 + 3);
        }
        for (int i = nv * 4 + lane; i < n; i += 32) best_take(b, src[i], i);
    } else {
        const uint16_t* src = reinterpret_cast<const uint16_t*>(x) + row * n;
        const int nv = vec_ok ? n / 8 : 0;
        for (int v = lane; v < nv; v += 32) {
            float f[8];
            if (DT == 2) unpack8<true>(__ldg(reinterpret_cast<const uint4*>(src) + v), f);
            else unpack8<false>(__ldg(reinterpret_cast<const uint4*>(src) + v), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) best_take(b, f[j], v * 8 + j);
        }
        for (int i = nv * 8 + lane; i < n; i += 32)
            best_take(b, DT == 2 ? u16_to_f<true>(src[i]) : u16_to_f<false>(src[i]), i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best c;
        c.v = __shfl_xor_sync(0xffffffffu, b.v, o);
        c.i = __shfl_xor_sync(0xffffffffu, b.i, o);
        c.nan = __shfl_xor_sync(0xffffffffu, static_cast<int>(b.nan), o) != 0;
        if (best_better(c, b)) b = c;
    }
    if (lane == 0) out[row] = b.i < 0 ? 0 : b.i;
}
int argmax_rows(const void* x, int dtype, long long rows, int n, long long* out, cudaStream_t s) {
    ProfScope prof_scope("argmax_rows", s);
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((rows * 32 + threads - 1) / threads);
    if (rows == 0) return 0;
    if (dtype == 0) argmax_rows_k<0><<<blocks, threads, 0, s>>>(x, rows, n, out);
    else if (dtype == 1) argmax_rows_k<1><<<blocks, threads, 0, s>>>(x, rows, n, out);
    else if (dtype == 2) argmax_rows_k<2><<<blocks, threads, 0, s>>>(x, rows, n, out);
    else {
        set_error("argmax_rows: unknown dtype %d", dtype);
        return -3;
    }
    WFD_CHECK_LAUNCH("argmax_rows");
    return 0;
}

// ------------------------------------------------------------------------------------------------ rule violations
__global__ void __launch_bounds__(256)
count_violations_k(const long long* __restrict__ tiles, const uint8_t* __restrict__ mat_x,
                   const uint8_t* __restrict__ mat_y, int T, int H, int W, unsigned long long* __restrict__ out) {
    __shared__ unsigned int sh[2];
    const int b = blockIdx.y;
    if (threadIdx.x < 2) sh[threadIdx.x] = 0u;
    __syncthreads();
    const long long* tb = tiles + static_cast<long long>(b) * H * W;
    unsigned int bad_x = 0u, bad_y = 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        const long long t = tb[i];
        const bool t_ok = t >= 0 && t < T;
        if (x + 1 < W) {
            const long long r = tb[i + 1];
            const bool ok = t_ok && r >= 0 && r < T && mat_x[t * T + r] != 0;
            bad_x += ok ? 0u : 1u;
        }
        if (y + 1 < H) {
            const long long d = tb[i + W];
            const bool ok = t_ok && d >= 0 && d < T && mat_y[t * T + d] != 0;
            bad_y += ok ? 0u : 1u;
        }
    }
    atomicAdd(&sh[0], bad_x);
    atomicAdd(&sh[1], bad_y);
    __syncthreads();
    if (threadIdx.x < 2 && sh[threadIdx.x] != 0u)
        atomicAdd(out + static_cast<long long>(b) * 2 + threadIdx.x, static_cast<unsigned long long>(sh[threadIdx.x]));
}
int count_violations(const long long* tiles, const uint8_t* mat_x, const uint8_t* mat_y, int T, int B, int H, int W,
                     long long* out, cudaStream_t s) {
    ProfScope prof_scope("count_violations", s);
    if (B < 1 || H < 1 || W < 1 || T < 1) {
        set_error("count_violations: bad shape B=%d H=%d W=%d T=%d", B, H, W, T);
        return -3;
    }
    cudaError_t e = cudaMemsetAsync(out, 0, static_cast<size_t>(B) * 2 * sizeof(long long), s);
    if (e != cudaSuccess) {
        set_error("count_violations: %s", cudaGetErrorString(e));
        return -8;
    }
    int bx = (H * W + 255) / 256;
    if (bx > 64) bx = 64;
    count_violations_k<<<dim3(bx, B), 256, 0, s>>>(tiles, mat_x, mat_y, T, H, W, reinterpret_cast<unsigned long long*>(out));
    WFD_CHECK_LAUNCH("count_violations");
    return 0;
}

// ------------------------------------------------------------------------------------------------ encoder ends
// conv_in of the tile-VAE encoder on a ONE-HOT input (an integer tile map): a grouped 3x3 conv whose input has a single
// non-zero channel per pixel is a gather - every neighbour (dy, dx) of an output pixel adds the cout_g weights
// W[g*cout_g + j, tile % cin_g, tap] of its tile's group g = tile / cin_g (reference wf_vae.py:90, img2img :68-74).
// table: [C_in][9][cout_g] fp32. One block per output pixel; out nhwc16 [B, H*W, C0].
template <bool BF16>
__global__ void __launch_bounds__(128)
onehot_conv_in_k(const long long* __restrict__ tiles, const float* __restrict__ table, const float* __restrict__ bias,
                 uint16_t* __restrict__ out, int H, int W, int C_in, int C0, int cin_g, int cout_g) {
    extern __shared__ float acc[];  // [C0]
    const long long pix = blockIdx.x;
    const int b = static_cast<int>(pix / (H * W));
    const int pr = static_cast<int>(pix - static_cast<long long>(b) * H * W);
    const int y = pr / W, x = pr - y * W;
    for (int c = threadIdx.x; c < C0; c += blockDim.x) acc[c] = bias[c];
    __syncthreads();
    for (int tap = 0; tap < 9; ++tap) {  // taps one after the other: two neighbours may hit the same group
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const long long t = tiles[(static_cast<long long>(b) * H + yy) * W + xx];
            if (t >= 0 && t < C_in && static_cast<int>(threadIdx.x) < cout_g) {
                const int g = static_cast<int>(t) / cin_g;
                acc[g * cout_g + threadIdx.x] += table[(static_cast<long long>(t) * 9 + tap) * cout_g + threadIdx.x];
            }
        }
        __syncthreads();
    }
    uint16_t* dst = out + pix * C0;
    for (int c = threadIdx.x * 2; c < C0; c += blockDim.x * 2)
        *reinterpret_cast<uint32_t*>(dst + c) = static_cast<uint32_t>(f_to_u16<BF16>(acc[c])) |
                                                (static_cast<uint32_t>(f_to_u16<BF16>(acc[c + 1])) << 16);
}
int onehot_conv_in(const long long* tiles, const float* table, const float* bias, void* out, int B, int H, int W,
                   int C_in, int C0, int cin_g, int cout_g, int bf16, cudaStream_t s) {
    ProfScope prof_scope("onehot_conv_in", s);
    if (cout_g > 128 || C0 % 2 || C0 * sizeof(float) > 48 * 1024) {
        set_error("onehot_conv_in: unsupported shape (C0=%d cout_g=%d)", C0, cout_g);
        return -3;
    }
    const unsigned blocks = static_cast<unsigned>(static_cast<long long>(B) * H * W);
    WFD_DISPATCH_BF16(bf16, (onehot_conv_in_k<BF16><<<blocks, 128, C0 * sizeof(float), s>>>(
                                tiles, table, bias, static_cast<uint16_t*>(out), H, W, C_in, C0, cin_g, cout_g)));
    WFD_CHECK_LAUNCH("onehot_conv_in");
    return 0;
}
// moments [B, HW, ld] fp32 (channels-last, the first C columns valid) -> NCHW fp32 [B, C, HW]
__global__ void moments_out_k(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int ld, long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i % HW);
        const long long bc = i / HW;
        const int c = static_cast<int>(bc % C);
        const long long b = bc / C;
        out[i] = in[(b * HW + p) * ld + c];
    }
}
static unsigned ew_blocks_decl(long long n, int threads);
int moments_out(const float* in, float* out, int B, int C, int HW, int ld, cudaStream_t s) {
    ProfScope prof_scope("moments_out", s);
    const long long n = static_cast<long long>(B) * C * HW;
    moments_out_k<<<ew_blocks_decl(n, 256), 256, 0, s>>>(in, out, C, HW, ld, n);
    WFD_CHECK_LAUNCH("moments_out");
    return 0;
}

// ------------------------------------------------------------------------------------------------ guidance-loop glue
// The elementwise work around the guidance step when scheduler.step is affine in the latents, prev = a x + b eps
// (DDIM eta = 0, PNDM/PLMS; reference pipeline_wfd.py:601-603, 611-614 and the final-stage twin :630-637):
//   z    = a * lat + b * (eps_u + g * (eps_t - eps_u))        classifier-free combine + scheduler.step in one pass
//   lat -= a * dz                                              chain factor of the step applied to d loss / d z, latent update
__global__ void guidance_affine_k(const float* __restrict__ lat, const float* __restrict__ eps_u,
                                  const float* __restrict__ eps_t, float g, float a, float b, float* __restrict__ z,
                                  long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float e = eps_u[i];
        if (eps_t) e = e + g * (eps_t[i] - e);
        z[i] = a * lat[i] + b * e;
    }
}
__global__ void guidance_update_k(float* __restrict__ lat, const float* __restrict__ dz, float a, long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        lat[i] = lat[i] - a * dz[i];
}
static unsigned ew_blocks_decl(long long n, int threads);
int guidance_affine(const float* lat, const float* eps_u, const float* eps_t, float g, float a, float b, float* z,
                    long long n, cudaStream_t s) {
    ProfScope prof_scope("guidance_affine", s);
    guidance_affine_k<<<ew_blocks_decl(n, 256), 256, 0, s>>>(lat, eps_u, eps_t, g, a, b, z, n);
    WFD_CHECK_LAUNCH("guidance_affine");
    return 0;
}
int guidance_update(float* lat, const float* dz, float a, long long n, cudaStream_t s) {
    ProfScope prof_scope("guidance_update", s);
    guidance_update_k<<<ew_blocks_decl(n, 256), 256, 0, s>>>(lat, dz, a, n);
    WFD_CHECK_LAUNCH("guidance_update");
    return 0;
}

static unsigned ew_blocks_decl(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}

// ------------------------------------------------------------------------------------------------ tile post-processing
// What the reference's callers do on the host after the argmax (wfd/webui/ui.py:228-243): optional mirror symmetry
// (scmap/data_processing.py:191-196: [tiles | mapping[tiles reversed along x]]) and the id chain
// gen_to_sc[shrink_to_gen[tile]] (scmap/map_display.py:55, data_processing.py:206-211). Integer gathers, bit-exact.
__global__ void __launch_bounds__(256)
tiles_finish_k(const long long* __restrict__ tiles, int B, int H, int W, const int* __restrict__ symm, int n_symm,
               const int* __restrict__ shrink_to_gen, int n_shrink, const int* __restrict__ gen_to_sc, int n_gen,
               long long* __restrict__ tiles_out, uint16_t* __restrict__ sc_out, unsigned long long* __restrict__ bad) {
    const int Wo = symm ? 2 * W : W;
    const long long total = static_cast<long long>(B) * H * Wo;
    unsigned int my_bad = 0u;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % Wo);
        const long long row = i / Wo;  // b * H + y
        long long t;
        bool ok = true;
        if (x < W) {
            t = tiles[row * W + x];
        } else {
            const long long src = tiles[row * W + (2 * W - 1 - x)];  // mirrored column
            ok = src >= 0 && src < n_symm;
            t = ok ? symm[src] : -1;
        }
        if (tiles_out) tiles_out[i] = t;
        int sc = 0xFFFF;
        if (ok && t >= 0 && t < n_shrink) {
            const int g = shrink_to_gen ? shrink_to_gen[t] : static_cast<int>(t);
            if (g >= 0 && g < n_gen) sc = gen_to_sc[g];
            else ok = false;
        } else {
            ok = false;
        }
        if (!ok) ++my_bad;
        if (sc_out) sc_out[i] = static_cast<uint16_t>(sc);
    }
    if (my_bad) atomicAdd(bad, static_cast<unsigned long long>(my_bad));
}
int tiles_finish(const long long* tiles, int B, int H, int W, const int* symm, int n_symm, const int* shrink_to_gen,
                 int n_shrink, const int* gen_to_sc, int n_gen, long long* tiles_out, uint16_t* sc_out, long long* bad,
                 cudaStream_t s) {
    ProfScope prof_scope("tiles_finish", s);
    if (B < 1 || H < 1 || W < 1 || !gen_to_sc || n_gen < 1 || !bad) {
        set_error("tiles_finish: bad arguments (B=%d H=%d W=%d n_gen=%d)", B, H, W, n_gen);
        return -3;
    }
    if (!shrink_to_gen) n_shrink = n_gen;
    cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(long long), s);
    if (e != cudaSuccess) {
        set_error("tiles_finish: %s", cudaGetErrorString(e));
        return -8;
    }
    const long long total = static_cast<long long>(B) * H * (symm ? 2 * W : W);
    tiles_finish_k<<<ew_blocks_decl(total, 256), 256, 0, s>>>(tiles, B, H, W, symm, n_symm, shrink_to_gen, n_shrink, gen_to_sc,
                                                             n_gen, tiles_out, sc_out, reinterpret_cast<unsigned long long*>(bad));
    WFD_CHECK_LAUNCH("tiles_finish");
    return 0;
}

// ------------------------------------------------------------------------------------------------ layout conversion
template <bool BF16>
__global__ void __launch_bounds__(256)
nchw_to_nhwc16_k(const void* __restrict__ in, int in_f16, uint16_t* __restrict__ out, int C, int HW) {
    __shared__ float tile[64][65];
    const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int i = ty; i < 64; i += 4) {
        const int c = c0 + i, p = p0 + tx;
        if (c < C && p < HW) {
            const long long idx = (static_cast<long long>(b) * C + c) * HW + p;
            tile[i][tx] = in_f16 ? __half2float(reinterpret_cast<const __half*>(in)[idx])
                                 : reinterpret_cast<const float*>(in)[idx];
        }
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 4) {
        const int p = p0 + i, c = c0 + tx;
        if (c < C && p < HW) out[(static_cast<long long>(b) * HW + p) * C + c] = f_to_u16<BF16>(tile[tx][i]);
    }
}
int nchw_to_nhwc16(const void* in, int in_f16, void* out, int B, int C, int HW, int bf16, cudaStream_t s) {
    ProfScope prof_scope("nchw_to_nhwc16", s);
    dim3 grid((HW + 63) / 64, (C + 63) / 64, B);
    WFD_DISPATCH_BF16(bf16, (nchw_to_nhwc16_k<BF16><<<grid, 256, 0, s>>>(in, in_f16, static_cast<uint16_t*>(out), C, HW)));
    WFD_CHECK_LAUNCH("nchw_to_nhwc16");
    return 0;
}
template <bool BF16>
__global__ void __launch_bounds__(256)
nhwc16_to_nchw_k(const uint16_t* __restrict__ in, void* __restrict__ out, int out_f16, int C, int HW) {
    __shared__ float tile[64][65];
    const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int i = ty; i < 64; i += 4) {
        const int p = p0 + i, c = c0 + tx;
        if (c < C && p < HW) tile[i][tx] = u16_to_f<BF16>(in[(static_cast<long long>(b) * HW + p) * C + c]);
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 4) {
        const int c = c0 + i, p = p0 + tx;
        if (c < C && p < HW) {
            const long long idx = (static_cast<long long>(b) * C + c) * HW + p;
            if (out_f16) reinterpret_cast<__half*>(out)[idx] = __float2half_rn(tile[tx][i]);
            else reinterpret_cast<float*>(out)[idx] = tile[tx][i];
        }
    }
}
int nhwc16_to_nchw(const void* in, void* out, int out_f16, int B, int C, int HW, int bf16, cudaStream_t s) {
    ProfScope prof_scope("nhwc16_to_nchw", s);
    dim3 grid((HW + 63) / 64, (C + 63) / 64, B);
    WFD_DISPATCH_BF16(bf16, (nhwc16_to_nchw_k<BF16><<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(in), out, out_f16, C, HW)));
    WFD_CHECK_LAUNCH("nhwc16_to_nchw");
    return 0;
}

template <bool BF16>
__global__ void cvt16_to_f32_k(const uint16_t* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = u16_to_f<BF16>(in[i]);
}
template <bool BF16>
__global__ void cvt_f32_to_16_k(const float* __restrict__ in, uint16_t* __restrict__ out, long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = f_to_u16<BF16>(in[i]);
}
template <bool BF16>
__global__ void add16_k(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, long long nv) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float fa[8], fb[8];
        unpack8<BF16>(__ldg(a + i), fa);
        unpack8<BF16>(__ldg(b + i), fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] += fb[j];
        out[i] = pack8<BF16>(fa);
    }
}
static unsigned ew_blocks(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}
int cvt16_to_f32(const void* in, float* out, long long n, int bf16, cudaStream_t s) {
    ProfScope prof_scope("cvt16_to_f32", s);
    WFD_DISPATCH_BF16(bf16, (cvt16_to_f32_k<BF16><<<ew_blocks(n, 256), 256, 0, s>>>(static_cast<const uint16_t*>(in), out, n)));
    WFD_CHECK_LAUNCH("cvt16_to_f32");
    return 0;
}
int cvt_f32_to_16(const float* in, void* out, long long n, int bf16, cudaStream_t s) {
    ProfScope prof_scope("cvt_f32_to_16", s);
    WFD_DISPATCH_BF16(bf16, (cvt_f32_to_16_k<BF16><<<ew_blocks(n, 256), 256, 0, s>>>(in, static_cast<uint16_t*>(out), n)));
    WFD_CHECK_LAUNCH("cvt_f32_to_16");
    return 0;
}
template <bool BF16>
__global__ void cvt_f32_to_16_2d_k(const float* __restrict__ in, long long ld_in, uint16_t* __restrict__ out,
                                   long long ld_out, long long rows, int cols) {
    const long long n = rows * cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / cols;
        const int c = static_cast<int>(i - r * cols);
        out[r * ld_out + c] = f_to_u16<BF16>(in[r * ld_in + c]);
    }
}
int cvt_f32_to_16_2d(const float* in, long long ld_in, void* out, long long ld_out, long long rows, int cols, int bf16,
                     cudaStream_t s) {
    ProfScope prof_scope("cvt_f32_to_16_2d", s);
    WFD_DISPATCH_BF16(bf16, (cvt_f32_to_16_2d_k<BF16><<<ew_blocks(rows * cols, 256), 256, 0, s>>>(
                                in, ld_in, static_cast<uint16_t*>(out), ld_out, rows, cols)));
    WFD_CHECK_LAUNCH("cvt_f32_to_16_2d");
    return 0;
}
int add16(const void* a, const void* b, void* out, long long n, int bf16, cudaStream_t s) {
    ProfScope prof_scope("add16", s);
    if (n % 8 != 0) {
        set_error("add16: n must be a multiple of 8");
        return -3;
    }
    WFD_DISPATCH_BF16(bf16, (add16_k<BF16><<<ew_blocks(n / 8, 256), 256, 0, s>>>(
                                static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), n / 8)));
    WFD_CHECK_LAUNCH("add16");
    return 0;
}

}  // namespace wfd
