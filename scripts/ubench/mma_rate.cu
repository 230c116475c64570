// Micro-benchmark (diagnosis only, not part of the library): cycles per tcgen05.mma (kind::f16, K = 16) as a function of
// the N extent and the CTA group, issued back to back by one thread with operands resident in shared memory (no TMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I wavefunctiondiffusion_b200/csrc scripts/ubench/mma_rate.cu -o /tmp/mma_rate
#include "ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace wfd;

template <int CG>
__global__ void __launch_bounds__(384, 1) mma_rate_k(int N, int n_mma, int reps, int a_stride16, long long* out, int commit_every, int fill) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2;  // receives the intermediate commits (never waited on)
    __shared__ uint64_t bar3, bar4;
    __shared__ uint64_t spin_bar;           // fill 13 / 14: never completes; warps 4-11 poll it like idle epilogue warps poll theirs
    __shared__ volatile int spin_done;
    __shared__ uint64_t rfull[4], rempty[4];  // fill == 11 / 12: a ring between the MMA warp and a helper warp (the producer's role)
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 208 * 1024 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = fill ? (0x3f803f80u ^ ((i * 2654435761u) & 0x007f007fu)) : 0u;  // bf16 values near 1
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        mbar_init(&bar3, 1);
        mbar_init(&spin_bar, 1);
        spin_done = 0;
        mbar_init(&bar4, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CG == 2) { tmem_alloc_2sm(&tmem_slot, 512); tmem_relinquish_2sm(); }
        else { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    // shfl from lane 0 marks the value warp-uniform for the compiler (as the library kernel does): without it the pattern
    // loop below is compiled into an ELECT / R2UR.BROADCAST loop per MMA and measures that loop, not the tensor pipe
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = make_idesc_f16(128 * CG, (uint32_t)N, 1u);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 80 * 1024);
        long long best = 1ll << 60;
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            if (fill >= 2) {
                // The patch schedule's issue sequence, written like the library kernel: the loop state lives in warp-uniform
                // control flow (uniform registers), only the MMA runs and the commit sit inside the elected region.
                // Per tile: 3 vertical offsets x 3 horizontal x n_k K steps (commit_every carries n_k).
                // fill: 2 = the kernel's pattern (unaligned starts, SBO 1280); 3 = aligned starts, SBO 1024; 4 = 2 + a
                // tcgen05 fence per vertical offset; 6 = aligned starts, SBO 1280; 7 = unaligned starts, SBO 1024;
                // 8 = one fixed odd 128-byte row for all nine taps, SBO 1280; 9 = 2 without the per-offset commits
                const bool aligned = (fill == 3 || fill == 6);
                const uint32_t a_hi = (uint32_t)(make_kmajor_desc_sbo(0, (fill == 3 || fill == 7) ? 1024u : 1280u) >> 32);
                const uint32_t b_hi = (uint32_t)(make_kmajor_desc(0, 128) >> 32);
                const uint32_t a_lo0 = (uint32_t)make_kmajor_desc(a_base, 128), b_lo0 = (uint32_t)make_kmajor_desc(b_base, 128);
                const uint32_t bbox16 = (uint32_t)(N / CG) * 8u;
                const uint32_t dx1 = (aligned || fill == 8) ? 0u : 8u, dx2 = (aligned || fill == 8) ? 0u : 16u;
                const int n_tiles = n_mma / 27;
                uint32_t a_st = a_lo0;
                int st = 0, acc = 0;
                // fill 10: three more commits per tile (the kernel signals weight stage, activation stage and accumulator);
                // fill 11 / 12: before a tile the MMA warp waits for a stage that a helper warp hands back once the MMAs that
                // read it have COMPLETED (ring of 3 / 2 stages): the producer round trip of the real kernel without any TMA
                const int ring = fill == 12 ? 2 : 3;
                int rst = 0;
                uint32_t rph = 0;  // n_tiles is a multiple of 2 * ring (checked on the host): every rep starts on phase 0
                for (int tile = 0; tile < n_tiles; ++tile) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                    if (fill == 11 || fill == 12) {
                        mbar_wait(&rfull[rst], rph);
                        tc_fence_after();
                    }
                    for (int d = 0; d < 3; ++d) {
                        if (fill == 4) tc_fence_after();
                        const uint32_t ad = a_st + (aligned ? d * 64u : (fill == 8 ? 8u : d * 80u)), bd = b_lo0 + d * 3u * bbox16;
                        if (elect_one()) {
                            if (CG == 2) {
                                umma_f16_run_2sm(d_tmem, ad, a_hi, bd, b_hi, idesc, d > 0 ? 1u : 0u, 2u, commit_every);
                                umma_f16_run_2sm(d_tmem, ad + dx1, a_hi, bd + bbox16, b_hi, idesc, 1u, 2u, commit_every);
                                umma_f16_run_2sm(d_tmem, ad + dx2, a_hi, bd + 2u * bbox16, b_hi, idesc, 1u, 2u, commit_every);
                                if (fill != 9) umma_commit_2sm(&bar2);
                            } else {
                                umma_f16_run(d_tmem, ad, a_hi, bd, b_hi, idesc, d > 0 ? 1u : 0u, 2u, commit_every);
                                umma_f16_run(d_tmem, ad + dx1, a_hi, bd + bbox16, b_hi, idesc, 1u, 2u, commit_every);
                                umma_f16_run(d_tmem, ad + dx2, a_hi, bd + 2u * bbox16, b_hi, idesc, 1u, 2u, commit_every);
                                if (fill != 9) umma_commit(&bar2);
                            }
                        }
                        __syncwarp();
                    }
                    if (fill >= 10 && elect_one()) {
                        if (CG == 2) {
                            if (fill == 10) umma_commit_2sm(&bar3); else umma_commit_2sm(&rempty[rst]);
                            umma_commit_2sm(&bar4);
                        } else {
                            if (fill == 10) umma_commit(&bar3); else umma_commit(&rempty[rst]);
                            umma_commit(&bar4);
                        }
                    }
                    __syncwarp();
                    if (++rst == ring) { rst = 0; rph ^= 1; }
                    a_st += 1472u;  // 23552-byte stages
                    if (++st == 3) { st = 0; a_st = a_lo0; }
                    acc ^= 1;
                }
                if (elect_one()) {
                    if (CG == 2) umma_commit_2sm(&bar);
                    else umma_commit(&bar);
                }
            } else if (elect_one()) {
                int since = 0;
                for (int i = 0; i < n_mma; ++i) {
                    // operand addresses walk through a 48 KB window like the stages of the real kernel
                    const uint32_t off = (uint32_t)((i * a_stride16) & 0xBFF) << 4;
                    const uint64_t adesc = make_kmajor_desc(a_base + off, 128);
                    const uint64_t bdesc = make_kmajor_desc(b_base + (off & 0x1FFF), 128);
                    if (CG == 2) umma_f16_2sm(tmem_base, adesc, bdesc, idesc, i > 0 ? 1u : 0u);
                    else umma_f16(tmem_base, adesc, bdesc, idesc, i > 0 ? 1u : 0u);
                    if (++since == commit_every && i + 1 < n_mma) {  // (no runtime modulo: a division costs ~200 cycles here)
                        since = 0;
                        if (CG == 2) umma_commit_2sm(&bar2);
                        else umma_commit(&bar2);
                    }
                }
                if (CG == 2) umma_commit_2sm(&bar);
                else umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, phase);
            phase ^= 1;
            const long long dt = clock64() - t0;
            if (dt < best) best = dt;
        }
        spin_done = 1;
        if (lane == 0 && blockIdx.x == 0) out[0] = best;
    } else if (warp >= 4 && (fill == 13 || fill == 14)) {
        // fill 13: eight warps per CTA poll a barrier with try_wait (may suspend), fill 14: with test_wait (pure spin)
        while (!spin_done) {
            if (fill == 13) mbar_try_wait(&spin_bar, 0);
            else mbar_test_wait(&spin_bar, 0);
        }
    } else if (warp == 2 && rank == 0 && (fill == 11 || fill == 12)) {
        const int ring = fill == 12 ? 2 : 3;
        const int n_tiles = n_mma / 27;
        int rst = 0;
        uint32_t rph = 0;
        for (int r = 0; r < reps; ++r)
            for (int tile = 0; tile < n_tiles; ++tile) {
                mbar_wait(&rempty[rst], rph ^ 1);   // free once the MMAs of the tile that used it have completed
                if (lane == 0) mbar_arrive(&rfull[rst]);
                __syncwarp();
                if (++rst == ring) { rst = 0; rph ^= 1; }
            }
    } else if (CG == 2 && warp == 0 && rank == 1) {
        // the peer's barrier receives the multicast commit too: consume the phases so the barrier never overflows
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            mbar_wait(&bar, phase);
            phase ^= 1;
        }
        spin_done = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

template <int CG>
static long long run(int N, int n_mma, int a_stride16, int grid, int commit_every = 0, int fill = 0) {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(mma_rate_k<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = 210 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_rate_k<CG>, N, n_mma, 20, a_stride16, d, commit_every, fill);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long h = -1;
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    else cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return h;
}

int main() {
    const int ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
    printf("cycles per tcgen05.mma kind::f16 K=16 (best of 20 batches), operands in shared memory\n");
    printf("%5s %5s %8s | %10s %10s | %10s %10s\n", "cg", "N", "n_mma", "1 CTA", "full grid", "cyc/mma", "cyc/mma(g)");
    for (int cg = 1; cg <= 2; ++cg)
        for (int N : ns) {
            if (cg == 2 && N % 16) continue;
            for (int n_mma : {32, 256}) {
                const long long a = cg == 1 ? run<1>(N, n_mma, 2, 1) : run<2>(N, n_mma, 2, 2);
                const long long b = cg == 1 ? run<1>(N, n_mma, 2, 148) : run<2>(N, n_mma, 2, 148);
                printf("%5d %5d %8d | %10lld %10lld | %10.1f %10.1f\n", cg, N, n_mma, a, b, (double)a / n_mma, (double)b / n_mma);
            }
        }
    // A rows read with a large stride between consecutive MMAs (different swizzle atoms) vs the same 32 bytes
    for (int s : {0, 2, 64, 130})
        printf("cg=2 N=48 a_stride16=%d: %.1f cyc/mma\n", s, (double)run<2>(48, 256, s, 148) / 256);
    // patch-schedule pattern: commit_every carries n_k (MMAs per run); cycles per ISSUED MMA = total / (n_mma / 27 * 9 * n_k)
    for (int fill : {2, 13, 14})
        for (int nk : {3, 4})
            printf("pattern fill=%d n_k=%d: cg=2 N=48 %.1f cyc/mma | cg=1 N=48 %.1f | cg=2 N=96 %.1f | cg=2 N=192 %.1f\n", fill, nk,
                   (double)run<2>(48, 324, 2, 148, nk, fill) / (108 * nk), (double)run<1>(48, 324, 2, 148, nk, fill) / (108 * nk),
                   (double)run<2>(96, 324, 2, 148, nk, fill) / (108 * nk), (double)run<2>(192, 324, 2, 148, nk, fill) / (108 * nk));
    for (int fill : {0})
        for (int ce : {0, 9})
            printf("cg=2 N=48 fill=%d commit_every=%d: %.1f cyc/mma | cg=1: %.1f | cg=2 N=192: %.1f\n", fill, ce,
                   (double)run<2>(48, 270, 2, 148, ce, fill) / 270, (double)run<1>(48, 270, 2, 148, ce, fill) / 270,
                   (double)run<2>(192, 270, 2, 148, ce, fill) / 270);
    return 0;
}
