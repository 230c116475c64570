// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA/TMEM).
// Everything here is Blackwell-only; there is deliberately no fallback path.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace wfd {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, px;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spins on the barrier; a wait that lasts longer than ~2 s of SM clocks is a protocol bug, so it traps (the launch then
// fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFu) == 0 && clock64() - t0 > 4000000000LL) {
#ifdef WFD_WATCHDOG_PRINTF
            // a printf is a call: it makes the compiler save/restore convergence barriers (BMOV) around every wait of
            // the hot loops, which slowed the TMA producer warps measurably - compile it in only to chase a hang
            printf("wfd: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
#endif
            asm volatile("trap;");
        }
    }
}

// Non-suspending variant: mbarrier.test_wait returns at once, where try_wait may put the thread to sleep for a
// system-dependent time when the phase is not complete yet. The single-thread roles of the contraction kernel (TMA
// producers, MMA issuer) hand short tiles to each other through chains of barriers; a sleep on every hand-over that is a
// few cycles early costs more than the MMAs of a narrow tile (scripts/diag_g64_layout.py, profiles/r02_summary.md).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
#ifdef WFD_NO_SPIN_WAIT
    mbar_wait(bar, parity);
#else
    if (mbar_test_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_test_wait(bar, parity)) {
        if ((++spins & 0xFFFFu) == 0 && clock64() - t0 > 4000000000LL) asm volatile("trap;");
    }
#endif
}

// accumulator wait of the epilogue warps: spins too unless -DWFD_EPI_TRY_WAIT (A/B knob)
__device__ __forceinline__ void mbar_wait_epi(uint64_t* bar, uint32_t parity) {
#ifdef WFD_EPI_TRY_WAIT
    mbar_wait(bar, parity);
#else
    mbar_wait_spin(bar, parity);
#endif
}

// ---------------------------------------------------------------- 256-bit global accesses (sm_100: LDG/STG.256)
// One instruction moves a thread's whole 32-byte piece; a row-per-thread epilogue is bound by the number of LSU
// instructions and line accesses, which this halves against two 16-byte accesses.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// 16-byte shared-memory accesses by shared-window address (a pointer derived from the aligned dynamic shared memory base
// is a generic pointer to the compiler: it would emit generic LD / ST)
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t (&v)[4]) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t (&v)[4]) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// L2 prefetch of a tensor box (no shared-memory destination, no barrier): hides DRAM latency of boxes needed a few tiles ahead
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same offset in CTA 0 of the pair (bit 24 selects the peer; cute Sm100MmaPeerBitMask)
__device__ __forceinline__ uint32_t leader_addr(const void* p) { return smem_u32(p) & 0xFEFFFFFFu; }
// arrive on the LEADER CTA's barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
// TMA loads issued by either CTA of a pair into its OWN shared memory; the bytes are counted on the leader's barrier
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_addr(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 256 x N x 16 MMA over the CTA pair: each CTA contributes its 128 rows of A and N/2 rows of B from its own shared
// memory (same offsets in both CTAs); issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 or bf16 operands, fp32 accumulate), issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// NK back-to-back MMAs over consecutive 16-element K steps of one operand pair, issued by ONE thread: the descriptors
// advance by a_step / 2 (16-byte units) inside the asm block, so the issuing thread spends two 32-bit adds per MMA instead
// of rebuilding both descriptors (the issue loop, not the tensor pipe, bounds narrow-N contractions: ~4 cycles per
// dependent uniform instruction, profiles/r02_summary.md). acc_first: accumulate flag of the first MMA; the rest accumulate.
// (one asm block per run length: a predicated-off tcgen05.mma still costs ~30 cycles of the issuing thread, measured with
// scripts/ubench/mma_rate.cu, so the run length is a uniform branch instead of a predicate)
__device__ __forceinline__ void umma_f16_run1(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc_first, uint32_t a_step) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 al, bl;\n\t"
        ".reg .b64 a, b;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b64 a, {%1, %2};\n\t"
        "mov.b64 b, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(a_step)
        : "memory");
}
__device__ __forceinline__ void umma_f16_run2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc_first, uint32_t a_step) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 al, bl;\n\t"
        ".reg .b64 a, b;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b64 a, {%1, %2};\n\t"
        "mov.b64 b, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, p;\n\t"
        "add.u32 al, %1, %7;\n\t"
        "add.u32 bl, %3, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(a_step)
        : "memory");
}
__device__ __forceinline__ void umma_f16_run3(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc_first, uint32_t a_step) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 al, bl;\n\t"
        ".reg .b64 a, b;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b64 a, {%1, %2};\n\t"
        "mov.b64 b, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, p;\n\t"
        "add.u32 al, %1, %7;\n\t"
        "add.u32 bl, %3, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.u32 al, al, %7;\n\t"
        "add.u32 bl, bl, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(a_step)
        : "memory");
}
__device__ __forceinline__ void umma_f16_run4(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc_first, uint32_t a_step) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 al, bl;\n\t"
        ".reg .b64 a, b;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b64 a, {%1, %2};\n\t"
        "mov.b64 b, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, p;\n\t"
        "add.u32 al, %1, %7;\n\t"
        "add.u32 bl, %3, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.u32 al, al, %7;\n\t"
        "add.u32 bl, bl, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.u32 al, al, %7;\n\t"
        "add.u32 bl, bl, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(a_step)
        : "memory");
}
// descriptors are passed as (low word, high word): only the address field in the low word moves; n_k in 1..4
__device__ __forceinline__ void umma_f16_run(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t acc_first, uint32_t a_step, int n_k) {
    if (n_k == 4) umma_f16_run4(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc_first, a_step);
    else if (n_k == 3) umma_f16_run3(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc_first, a_step);
    else if (n_k == 2) umma_f16_run2(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc_first, a_step);
    else umma_f16_run1(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc_first, a_step);
}
__device__ __forceinline__ void umma_f16_run1_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc_first, uint32_t a_step) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 al, bl;\n\t"
        ".reg .b64 a, b;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b64 a, {%1, %2};\n\t"
        "mov.b64 b, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(a_step)
        : "memory");
}
__device__ __forceinline__ void umma_f16_run2_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc_first, uint32_t a_step) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 al, bl;\n\t"
        ".reg .b64 a, b;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b64 a, {%1, %2};\n\t"
        "mov.b64 b, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, p;\n\t"
        "add.u32 al, %1, %7;\n\t"
        "add.u32 bl, %3, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, t;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(a_step)
        : "memory");
}
__device__ __forceinline__ void umma_f16_run3_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc_first, uint32_t a_step) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 al, bl;\n\t"
        ".reg .b64 a, b;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b64 a, {%1, %2};\n\t"
        "mov.b64 b, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, p;\n\t"
        "add.u32 al, %1, %7;\n\t"
        "add.u32 bl, %3, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.u32 al, al, %7;\n\t"
        "add.u32 bl, bl, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, t;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(a_step)
        : "memory");
}
__device__ __forceinline__ void umma_f16_run4_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc_first, uint32_t a_step) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 al, bl;\n\t"
        ".reg .b64 a, b;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b64 a, {%1, %2};\n\t"
        "mov.b64 b, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, p;\n\t"
        "add.u32 al, %1, %7;\n\t"
        "add.u32 bl, %3, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.u32 al, al, %7;\n\t"
        "add.u32 bl, bl, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.u32 al, al, %7;\n\t"
        "add.u32 bl, bl, 2;\n\t"
        "mov.b64 a, {al, %2};\n\t"
        "mov.b64 b, {bl, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, t;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(a_step)
        : "memory");
}
// descriptors are passed as (low word, high word): only the address field in the low word moves; n_k in 1..4
__device__ __forceinline__ void umma_f16_run_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t acc_first, uint32_t a_step, int n_k) {
    if (n_k == 4) umma_f16_run4_2sm(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc_first, a_step);
    else if (n_k == 3) umma_f16_run3_2sm(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc_first, a_step);
    else if (n_k == 2) umma_f16_run2_2sm(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc_first, a_step);
    else umma_f16_run1_2sm(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc_first, a_step);
}
// Arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: this thread's lane, 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        " {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// same, tied to the destination registers of an earlier tmem_ld16 so that no use of them can move above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// Shared-memory matrix descriptor, K-major operand tile whose rows are `row_bytes` (32/64/128) wide and stored with the
// matching TMA swizzle; 8-row groups are `8*row_bytes` apart (SBO). Field layout: cute/arch/mma_sm100_desc.hpp.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t row_bytes) {
    const uint32_t sbo = (8u * row_bytes) >> 4;
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;           // LBO (ignored for swizzled K-major; canonical value 1)
    d |= static_cast<uint64_t>(sbo & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;           // descriptor version 1 (sm_100)
    d |= layout << 61;
    return d;
}

// Same with an explicit stride between 8-row groups (SBO): the patch schedule reads 8-pixel row segments of a halo patch
// whose rows are (w_box + 2) pixels apart. The start address may sit anywhere on a 128-byte row boundary: the 128-byte
// swizzle is a function of the shared-memory address bits, which is also what makes the usual +32-byte K advance work.
__device__ __forceinline__ uint64_t make_kmajor_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= 2ull << 61;
    return d;
}

// MN-major operand tile with the 128-byte swizzle (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>): in 16-byte
// units the canonical layout is Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) - 64 elements of M (or N) are contiguous
// in a 128-byte row, 8 consecutive K rows form a 1024-byte atom, atoms along K are SBO apart and blocks of 64 along M
// are LBO apart. A TMA box [64 m][64 k] with SWIZZLE_128B is 8 such atoms back to back (SBO = 1024).
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // descriptor version 1 (sm_100)
    d |= 2ull << 61;                      // SWIZZLE_128B
    return d;
}

// Instruction descriptor for kind::f16: fp32 accumulate, both operands K-major, dense.

// is_bf16: 1 = bf16 operands, 0 = fp16 operands.
__device__ __host__ __forceinline__ uint32_t make_idesc_f16(uint32_t m, uint32_t n, uint32_t is_bf16) {
    uint32_t d = 0;
    d |= 1u << 4;              // c_format = F32
    d |= (is_bf16 & 1u) << 7;  // a_format
    d |= (is_bf16 & 1u) << 10; // b_format
    d |= (n >> 3) << 17;       // n_dim
    d |= (m >> 4) << 24;       // m_dim
    return d;
}
constexpr uint32_t IDESC_A_MN_MAJOR = 1u << 15;  // A operand is M-major (see make_mnmajor_desc)

}  // namespace wfd
