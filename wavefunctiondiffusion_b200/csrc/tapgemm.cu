// tapgemm kernel (tcgen05 + TMA + TMEM), its launcher and the CUDA-core checking kernel. See tapgemm.cuh.
#include "tapgemm.cuh"
#include "ptx.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <type_traits>
#include <vector>
#include <stdlib.h>

namespace wfd {

// ------------------------------------------------------------------------------------------------ error plumbing
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }
long long g_launch_count = 0;

// WFD_DEBUG_SYNC=1: synchronise after every launch and report the failing kernel by name (debugging aid only).
int debug_sync_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("WFD_DEBUG_SYNC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}
int debug_sync_check(const char* what, cudaStream_t stream) {
    if (!debug_sync_enabled()) return 0;
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_error("%s: device fault: %s", what, cudaGetErrorString(e));
        fprintf(stderr, "wfd debug: %s: device fault: %s\n", what, cudaGetErrorString(e));
        return -20;
    }
    return 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------ launch profiling
// wfd_profile_enable(1): every launch wrapper brackets its kernel with CUDA events on the launching stream;
// profile_report() synchronises and returns "name count total_ms" lines. Used by bench.py for the roofline numbers.
struct ProfRec {
    std::string name;
    cudaEvent_t a, b;
};
static int g_prof_on = 0;
static std::vector<ProfRec> g_prof;
void profile_enable(int on) {
    g_prof_on = on;
    if (on) {
        for (auto& r : g_prof) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        g_prof.clear();
    }
}
ProfScope::ProfScope(const char* name, cudaStream_t s) : stream(s), idx(-1) {
    if (!g_prof_on) return;
    ProfRec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    g_prof.push_back(r);
    idx = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof[idx].b, stream);
}
int profile_report(char* buf, size_t cap) {
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : g_prof) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& e = agg[r.name];
        e.first += 1;
        e.second += ms;
    }
    std::string out;
    char line[384];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && cap > 0) {
        strncpy(buf, out.c_str(), cap - 1);
        buf[cap - 1] = 0;
    }
    return static_cast<int>(out.size());
}

// ------------------------------------------------------------------------------------------------ device helpers
template <bool BF16>
__device__ __forceinline__ float lo16_to_f(uint32_t v) {
    if (BF16) return __uint_as_float(v << 16);
    return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xFFFF)));
}
template <bool BF16>
__device__ __forceinline__ float hi16_to_f(uint32_t v) {
    if (BF16) return __uint_as_float(v & 0xFFFF0000u);
    return __half2float(__ushort_as_half(static_cast<unsigned short>(v >> 16)));
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float fast_sigmoid(float u) {  // 0.5 * tanh(0.5 u) + 0.5 on the hardware tanh
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
    return fmaf(0.5f, t, 0.5f);
}

struct TileCoord {
    int z, n_tile, bb, ty, tx;
};
__device__ __forceinline__ TileCoord decode_tile(const TapGemmParams& p, int t, int cg, int rank) {
    // with cta_group::2 a tile id names a PAIR of adjacent M tiles; `rank` picks this CTA's half
    TileCoord c;
    const int m_count = p.m_tiles / cg;
    int m_tile, rest;
    if (p.n_group > 1) {
        // L2-friendly rasterisation: consecutive tile ids (= the CTAs of one wave) cover a patch of
        // (#SMs / n_group) M tiles x n_group N tiles, so each A tile is fetched from HBM once per band of N tiles
        const int per_z = m_count * p.n_tiles;
        c.z = t / per_z;
        const int r0 = t - c.z * per_z;
        const int per_band = m_count * p.n_group;
        const int band = r0 / per_band;
        const int r = r0 - band * per_band;
        const int nb = min(p.n_group, p.n_tiles - band * p.n_group);
        m_tile = r / nb;
        c.n_tile = band * p.n_group + (r - m_tile * nb);
        rest = 0;
    } else {
        m_tile = t % m_count;
        rest = t / m_count;
        c.n_tile = rest % p.n_tiles;
        c.z = rest / p.n_tiles;
    }
    m_tile = m_tile * cg + rank;
    c.tx = m_tile % p.tiles_x;
    int r2 = m_tile / p.tiles_x;
    c.ty = r2 % p.tiles_y;
    c.bb = r2 / p.tiles_y;
    return c;
}

// position (in 64-element chunks) of the (tap, cc) chunk inside a packed Bm row
__device__ __host__ __forceinline__ int b_chunk_index(const TapGemmParams& p, int tap, int cc) {
    if (!p.reuse) return tap * p.cpt + cc;
    // patch order: (cc, dy, dx) - the nine taps of a 64-channel chunk are adjacent
    if (p.reuse == 2) return cc * 9 + (p.tap_dy[tap] + 1) * 3 + (p.tap_dx[tap] + 1);
    // row-reuse order: (dx, cc, dy) so that the three vertical taps of one horizontal shift are adjacent
    return ((p.tap_dx[tap] + 1) * p.cpt + cc) * 3 + (p.tap_dy[tap] + 1);
}

// ------------------------------------------------------------------------------------------------ the kernel
// Two K-loop schedules:
//   plain      : one pipeline stage = one (tap, 64-channel chunk): A box [128 px][64 ch] + B box [BN][64]; 4 MMAs.
//   row reuse  : (3x3 stride-1 taps only) one stage = one (horizontal shift dx, chunk): A box [(h_box+2) rows x w_box px]
//                [64 ch] loaded ONCE and read by the three vertical taps through descriptor row offsets, + the three
//                matching B boxes fetched by a single 3-D TMA; 12 MMAs. A third of the TMA instructions, two thirds of
//                the activation bytes.
//   patch      : (3x3 stride-1 taps, 16 x 8 pixel tiles) ONE activation box per 64-channel chunk carries the whole halo
//                patch and serves all nine taps through descriptor start offsets. XF = true adds eight transform warps that
//                apply GroupNorm(+SiLU) to every landed box in shared memory (TapGemmParams::ax_stats); that variant lays
//                the roles out by warpgroup (tapgemm.cuh) and rebalances registers with setmaxnreg: 640 threads would
//                otherwise cap every role at 96 registers, and the epilogue needs ~160.
// EPI = p.epi_fast of the launch (tapgemm_prepare): every epilogue form is its own instantiation and compiles only its own
// drain - 0 the generic column loop (WFC contraction, fp32 outputs, odd shapes), 1 the plain conv drains, 2 the attention
// drains whose epilogue is the softmax work, 3 the backward convs with the fused GroupNorm-backward reduction. With all
// of them in one kernel the register allocator paid for every path in every launch (168 registers, 120 - 216 bytes of
// spills); alone, the plain conv drain needs 151 registers and no spill, and the groups=64 launches - bound by that
// drain - run 22% faster (batch-8 step 11.83 -> 10.84 / 11.07 ms in one call, `WFD_NO_EPI_SPLIT=1` on the former build).
template <bool BF16, int CG, bool XF, int EPI>
__global__ void __launch_bounds__(XF ? TG_THREADS_XF : TG_THREADS, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ TapGemmParams p, const int stages) {
    // Role counters (p.dbg) and engine switches (p.dbg_skip) are diagnosis tools: compiled in only with -DWFD_TG_DBG_BUILD
    // (WFD_NVCC_EXTRA=-DWFD_TG_DBG_BUILD=1 python -m wavefunctiondiffusion_b200.build --force). Each test of a kernel
    // parameter is a constant-bank load in the per-tile path of the single-thread roles - about ten per tile in the MMA
    // issuer alone, and short tiles pay for them (groups=64 conv: 0.143 -> 0.136 ms per launch).
#ifdef WFD_TG_DBG_BUILD
    const bool dbg_on = p.dbg != nullptr;
    const int dbg_skip = p.dbg_skip;
#else
    constexpr bool dbg_on = false;
    constexpr int dbg_skip = 0;
#endif
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const bool patch = p.reuse == 2;
    const int a_rows = patch ? (p.h_box + 2) * (p.w_box + 2) : (p.reuse ? (p.h_box + 2) * p.w_box : TG_BM);
    const uint32_t a_box_bytes = static_cast<uint32_t>(a_rows) * 128;        // what one activation TMA lands
    const uint32_t a_stage_bytes = (a_box_bytes + 1023u) & ~1023u;           // ring pitch (swizzle atoms are 1024 bytes)
    const uint32_t b_box_bytes = static_cast<uint32_t>(p.BN / CG) * 128;  // with CG == 2 each CTA holds half of the B rows
    // resident weights (small weight set): all 9*cpt chunks of the current N tile stay in shared memory and only the
    // activation boxes stream through a ring
    const bool b_res = p.b_resident && (CG == 1 || patch);
    const uint32_t b_stage_bytes = b_res ? 0u : (patch ? p.b_taps * b_box_bytes : (p.reuse ? 3 * b_box_bytes : b_box_bytes));
    const uint32_t b_res_bytes = b_res ? 9u * p.cpt * b_box_bytes : 0u;
    // the patch schedule keeps two rings: activation boxes (a_stages deep) and weight boxes (`stages` deep, one vertical
    // tap row = three taps per stage); the other schedules pair one activation box with one weight stage
    const int a_stages = patch ? p.a_stages : stages;
    uint8_t* smA = smem;
    uint8_t* smB = smem + static_cast<size_t>(a_stages) * a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + static_cast<size_t>(stages) * b_stage_bytes + b_res_bytes);
    uint64_t* full_bar = bars;                 // [stages]
    uint64_t* empty_bar = bars + stages;       // [stages]
    // TMEM accumulator stages: 512 columns hold p.acc_stages (2, 4 or 8) accumulators of acc_cols columns each. Narrow N
    // tiles get more than two: with two, a tile's MMAs cannot start before the epilogue of the tile before last has
    // drained its stage, and the per-tile time becomes (MMA time + epilogue latency) / 2 instead of the larger of the two
    const int acc_stages = p.acc_stages;
    const uint32_t acc_cols = 512u / static_cast<uint32_t>(acc_stages);
    uint64_t* tfull_bar = bars + 2 * stages;       // [<= 8]
    uint64_t* tempty_bar = bars + 2 * stages + 8;  // [<= 8]
    uint64_t* bres_full = bars + 2 * stages + 16;  // resident weights landed
    uint64_t* tdone_bar = bars + 2 * stages + 17;  // all MMAs of a tile completed (resident weights may be replaced)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 18);
    uint64_t* afull_bar = bars + 2 * stages + 20;  // [a_stages] patch schedule only
    uint64_t* aempty_bar = afull_bar + a_stages;   // [a_stages]
    uint64_t* aready_bar = aempty_bar + a_stages;  // [a_stages] XF only: the box is transformed (leader CTA's copy is used)
    // (an even number of 8-byte barriers keeps the float4 accesses of the staging arrays 16-byte aligned)
    float* bias_s = reinterpret_cast<float*>(bars + 2 * stages + 20 + (patch ? 4 * a_stages : 0));  // [2][256]
    float* gam_s = bias_s + 512;                                      // [2][256] GroupNorm-backward gamma
    float* bet_s = gam_s + 512;                                       // [2][256] GroupNorm-backward beta
    // per-CTA partial sums of the fused GroupNorm statistics / backward reductions, flushed once at kernel end: thousands
    // of tiles would otherwise hammer the same few global addresses with atomics
    unsigned long long* stat_s = reinterpret_cast<unsigned long long*>(bet_s + 512);  // [2][TG_STAT_SLOTS] fixed point
    float* xf_s = reinterpret_cast<float*>(stat_s + 2 * TG_STAT_SLOTS);  // [512][mean | rstd] per (sample, group), XF only
    float* gnb_tab = xf_s + (XF ? 2 * 512 : 0);  // [256][mean | rstd]: fast GroupNorm-backward epilogue only (never with XF)

#ifndef WFD_NO_UNIFORM_WARP
    // The warp index through a shuffle: the compiler then knows it is warp-uniform, the role branches below become uniform
    // branches, and the loop state of the single-thread roles (stage, phase, MMA descriptors) lives in uniform registers:
    // R2UR moves in the kernel drop from 91 to 12 and the MMA-issue loop has none left (cuobjdump, profiles/r02_sass.txt).
    // Round 1 measured 13.3 -> 12.85 ms per batch-8 step; -DWFD_NO_UNIFORM_WARP=1 restores the plain index for A/B runs.
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
#else
    const int warp = threadIdx.x >> 5;
#endif
    const int lane = threadIdx.x & 31;
    const int total_tiles = (p.m_tiles / CG) * p.n_tiles * p.z_count;  // pair tiles when CG == 2
    const int k_iters = patch ? p.cpt : (p.reuse ? 3 * p.cpt : p.n_taps * p.cpt);
    const int rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
    // tile ids of this CTA (pair): strided over the grid, or one contiguous range (tile_chunked)
    int first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG, tile_end = total_tiles;
    if (p.tile_chunked) {
        const long long q = blockIdx.x / CG, Q = gridDim.x / CG;
        first_tile = static_cast<int>(total_tiles * q / Q);
        tile_end = static_cast<int>(total_tiles * (q + 1) / Q);
        tile_step = 1;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], (b_res || patch) ? 1 : 2);
            mbar_init(&empty_bar[s], 1);
        }
        if (patch)
            for (int s = 0; s < a_stages; ++s) {
                mbar_init(&afull_bar[s], 1);
                mbar_init(&aempty_bar[s], 1);
                mbar_init(&aready_bar[s], (TG_XF_THREADS / 32) * CG);  // one arrival per transform warp of the pair
            }
        mbar_init(bres_full, 1);
        mbar_init(tdone_bar, 1);
        for (int a = 0; a < acc_stages; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], (p.epi_split ? TG_EPI_THREADS / 64 : TG_EPI_THREADS / 32) * CG);  // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 2 * TG_STAT_SLOTS; i += blockDim.x) stat_s[i] = 0ull;
    if (warp == 1) {
        if (CG == 2) {
            tmem_alloc_2sm(tmem_slot, 512);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_slot, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
    tc_fence_after();
    // shfl from lane 0 marks the value warp-uniform for the compiler: tcgen05/TMA operands then live in uniform
    // registers instead of going through an ELECT / R2UR.BROADCAST loop per instruction
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // warp roles (XF: warpgroup aligned, see tapgemm.cuh)
    constexpr int W_TMAB = XF ? 2 : TG_THREADS / 32 - 1;  // weight TMA producer
    constexpr int W_EPI0 = XF ? 4 : 2;                     // first of the eight epilogue warps
    constexpr int W_XF0 = 12;                              // first transform warp (XF only)
    // The roles are written as lambdas so that the XF variant can nest them under one setmaxnreg per warpgroup (ptxas
    // allocates registers for the code that a setmaxnreg dominates; .aligned wants the whole warpgroup on the same one).
    auto producer_role = [&]() {
        // ===================================================== TMA producers (warp-uniform loops, one elected lane issues):
        // warp 0 streams the activation boxes, the last warp the weight boxes; both arrive on the stage's full barrier
        const bool load_a = (warp == 0);
        {
            // the ring this warp feeds: patch schedule = activations and weights have separate rings and barriers
            const int ring = (patch && load_a) ? a_stages : stages;
            uint64_t* const fbar = (patch && load_a) ? afull_bar : full_bar;
            uint64_t* const ebar = (patch && load_a) ? aempty_bar : empty_bar;
            int stage = 0;
            uint32_t phase = 0;
            long long w_empty = 0, w_issue = 0;
            int tiles_seen = 0, res_n_tile = -1;
            const long long t_begin = clock64();
            // Tile coordinates cost several integer divisions and one global load (a_c0): each lane decodes a different
            // upcoming tile, the warp then walks through them with shuffles (32 tiles per batch).
            for (int base = first_tile; base < tile_end; base += 32 * tile_step) {
                const int my_t = base + lane * tile_step;
                TileCoord mine = decode_tile(p, my_t < tile_end ? my_t : base, CG, rank);
                int my_c0 = 0;
                if (p.a_c0 && my_t < tile_end) my_c0 = __ldg(p.a_c0 + mine.n_tile);
                const int batch = min(32, (tile_end - base + tile_step - 1) / tile_step);
                for (int l = 0; l < batch; ++l) {
                    TileCoord tc;
                    tc.z = __shfl_sync(0xffffffffu, mine.z, l);
                    tc.n_tile = __shfl_sync(0xffffffffu, mine.n_tile, l);
                    tc.bb = __shfl_sync(0xffffffffu, mine.bb, l);
                    tc.ty = __shfl_sync(0xffffffffu, mine.ty, l);
                    tc.tx = __shfl_sync(0xffffffffu, mine.tx, l);
                    const int c0 = __shfl_sync(0xffffffffu, my_c0, l);
                    if (b_res && !load_a) {
                        // weight warp, resident mode: reload only when the N tile changes, after every MMA that reads the
                        // old weights has completed
                        if (tc.n_tile != res_n_tile) {
                            // the MMA thread commits on tdone when it has issued the last tile that reads the old weights
                            // (it knows the next tile's N tile too): one phase per reload, so the parity never aliases
                            if (res_n_tile >= 0) {
                                mbar_wait_spin(tdone_bar, static_cast<uint32_t>(tiles_seen & 1));
                                ++tiles_seen;  // counts reloads that waited
                            }
                            res_n_tile = tc.n_tile;
                            const int n0r = tc.n_tile * p.BN + rank * (p.BN / CG);
                            if (elect_one()) {
                                if (CG == 2) {  // each CTA loads its half of the rows; the leader's barrier counts all bytes
                                    if (rank == 0) mbar_expect_tx(bres_full, 2 * b_res_bytes);
                                    if (patch) {
                                        for (int cc = 0; cc < p.cpt; ++cc)
                                            tma_load_3d_2sm(smB + static_cast<size_t>(cc) * 9 * b_box_bytes, &tmB, bres_full, 0, n0r, cc * 9);
                                    } else {
                                        for (int it = 0; it < 3 * p.cpt; ++it)
                                            tma_load_3d_2sm(smB + static_cast<size_t>(it) * 3 * b_box_bytes, &tmB, bres_full, 0, n0r, it * 3);
                                    }
                                } else {
                                    mbar_expect_tx(bres_full, b_res_bytes);
                                    if (patch) {
                                        for (int cc = 0; cc < p.cpt; ++cc)
                                            tma_load_3d(smB + static_cast<size_t>(cc) * 9 * b_box_bytes, &tmB, bres_full, 0, n0r, cc * 9);
                                    } else {
                                        for (int it = 0; it < 3 * p.cpt; ++it)
                                            tma_load_3d(smB + static_cast<size_t>(it) * 3 * b_box_bytes, &tmB, bres_full, 0, n0r, it * 3);
                                    }
                                }
                            }
                            __syncwarp();
                        }
                        continue;
                    }
                    const int x0 = tc.tx * p.w_box * p.in_stride;
                    const int y0 = tc.ty * p.h_box * p.in_stride + p.a_y_off;
                    const int ab = tc.bb + tc.z * p.a_zmul;
                    const int bz = tc.z * p.b_zmul + tc.bb * p.b_bbmul;
                    const int n0 = tc.n_tile * p.BN + rank * (p.BN / CG);  // this CTA's rows of the B tile
                    // The producer warps sit on the critical path of short-K and block-sparse contractions (one warp issues
                    // dependent instructions every few cycles): the three walks below keep their per-chunk work minimal, and
                    // each warp only computes the coordinates of its own operand.
                    auto slot_wait = [&]() {
                        if (dbg_on) {
                            const long long w0 = clock64();
                            mbar_wait_spin(&ebar[stage], phase ^ 1);
                            w_empty += clock64() - w0;
                        } else {
                            mbar_wait_spin(&ebar[stage], phase ^ 1);
                        }
                    };
                    auto slot_next = [&]() {
                        __syncwarp();
                        if (++stage == ring) {
                            stage = 0;
                            phase ^= 1;
                        }
                    };
                    auto issue_a = [&](int ca, int xx, int yy) {
                        slot_wait();
                        const long long i0 = dbg_on ? clock64() : 0;
                        if (dbg_skip & 1) {
                            if (elect_one() && (rank == 0 || XF)) mbar_arrive(&fbar[stage]);
                        } else if (elect_one()) {
                            if (XF) {  // each CTA's transform warps wait for their own box: the load signals the LOCAL barrier
                                mbar_expect_tx(&fbar[stage], a_box_bytes);
                                tma_load_4d(smA + static_cast<size_t>(stage) * a_stage_bytes, &tmA, &fbar[stage], ca, xx, yy, ab);
                            } else if (CG == 2) {  // both CTAs load into their own shared memory; the leader's barrier counts all bytes
                                if (rank == 0) mbar_expect_tx(&fbar[stage], 2 * a_box_bytes);
                                tma_load_4d_2sm(smA + static_cast<size_t>(stage) * a_stage_bytes, &tmA, &fbar[stage], ca,
                                                xx, yy, ab);
                            } else {
                                mbar_expect_tx(&fbar[stage], a_box_bytes);
                                tma_load_4d(smA + static_cast<size_t>(stage) * a_stage_bytes, &tmA, &fbar[stage], ca, xx, yy, ab);
                            }
                        }
                        if (dbg_on) w_issue += clock64() - i0;
                        slot_next();
                    };
                    auto issue_b = [&](int k0, int k2) {  // weights box at (k0, n0, k2)
                        slot_wait();
                        const long long i0 = dbg_on ? clock64() : 0;
                        if (dbg_skip & 2) {
                            if (elect_one() && rank == 0) mbar_arrive(&full_bar[stage]);
                        } else if (elect_one()) {
                            if (CG == 2) {
                                if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * b_stage_bytes);
                                tma_load_3d_2sm(smB + static_cast<size_t>(stage) * b_stage_bytes, &tmB, &full_bar[stage], k0, n0, k2);
                            } else {
                                mbar_expect_tx(&full_bar[stage], b_stage_bytes);
                                tma_load_3d(smB + static_cast<size_t>(stage) * b_stage_bytes, &tmB, &full_bar[stage], k0, n0, k2);
                            }
                        }
                        if (dbg_on) w_issue += clock64() - i0;
                        slot_next();
                    };
                    if (patch) {
                        // one halo patch per 64-channel chunk; weights: one vertical tap row (three taps) per stage
                        if (load_a) {
                            if (p.a_prefetch > 0 && l + p.a_prefetch < batch) {
                                // the tile a few slots ahead: pull its patch into L2 now (DRAM latency off the critical path)
                                const int lp = l + p.a_prefetch;
                                const int px0 = __shfl_sync(0xffffffffu, mine.tx, lp) * p.w_box;
                                const int py0 = __shfl_sync(0xffffffffu, mine.ty, lp) * p.h_box + p.a_y_off;
                                const int pab = __shfl_sync(0xffffffffu, mine.bb, lp);
                                const int pc0 = __shfl_sync(0xffffffffu, my_c0, lp);
                                if (elect_one())
                                    for (int cc = 0; cc < p.cpt; ++cc) tma_prefetch_4d(&tmA, pc0 + cc * TG_KC, px0 - 1, py0 - 1, pab);
                                __syncwarp();
                            }
                            for (int cc = 0; cc < p.cpt; ++cc) issue_a(c0 + cc * TG_KC, x0 - 1, y0 - 1);
                        } else {
                            for (int cc = 0; cc < p.cpt; ++cc)
                                for (int d = 0; d < 9; d += p.b_taps) issue_b(0, cc * 9 + d);
                        }
                    } else if (p.kc_list) {
                        // skip list (plain schedule): packed entries dx+1 | (dy+1) << 2 | chunk-in-tap << 4 | chunk << 16,
                        // fetched 32 at a time (one per lane) and handed round by shuffle
                        const int n_list = __ldg(p.kc_cnt + tc.n_tile);
                        const int* list = p.kc_list + static_cast<long long>(tc.n_tile) * k_iters;
                        // the batch after the current one is already in flight, so the load latency never stalls the ring
                        int mine_e = 0, next_e = (lane < n_list) ? __ldg(list + lane) : 0;
                        for (int j = 0; j < n_list; ++j) {
                            if ((j & 31) == 0) {
                                mine_e = next_e;
                                next_e = (j + 32 + lane < n_list) ? __ldg(list + j + 32 + lane) : 0;
                            }
                            const int e = __shfl_sync(0xffffffffu, mine_e, j & 31);
                            if (load_a) issue_a(c0 + ((e >> 4) & 0xfff) * TG_KC, x0 + (e & 3) - 1, y0 + ((e >> 2) & 3) - 1);
                            else issue_b((e >> 16) * TG_KC, bz);
                        }
                    } else if (p.reuse) {
                        if (load_a) {
                            for (int o = 0; o < 3; ++o)
                                for (int cc = 0; cc < p.cpt; ++cc) issue_a(c0 + cc * TG_KC, x0 + o - 1, y0 - 1);
                        } else {
                            for (int it = 0; it < k_iters; ++it) issue_b(0, it * 3);
                        }
                    } else if (load_a && p.a_mn_major) {
                        // M-major A: two [64 m][64 k] boxes per stage (m contiguous), coordinates (m, k, 0, sample)
                        for (int cc = 0; cc < p.cpt; ++cc) {
                            slot_wait();
                            if (elect_one()) {
                                uint8_t* dst = smA + static_cast<size_t>(stage) * a_stage_bytes;
                                if (CG == 2) {
                                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * a_stage_bytes);
                                    tma_load_4d_2sm(dst, &tmA, &full_bar[stage], x0, cc * TG_KC, 0, ab);
                                    tma_load_4d_2sm(dst + 8192, &tmA, &full_bar[stage], x0 + 64, cc * TG_KC, 0, ab);
                                } else {
                                    mbar_expect_tx(&full_bar[stage], a_stage_bytes);
                                    tma_load_4d(dst, &tmA, &full_bar[stage], x0, cc * TG_KC, 0, ab);
                                    tma_load_4d(dst + 8192, &tmA, &full_bar[stage], x0 + 64, cc * TG_KC, 0, ab);
                                }
                            }
                            slot_next();
                        }
                    } else if (load_a) {
                        for (int o = 0; o < p.n_taps; ++o) {
                            const int xx = x0 + p.tap_dx[o], yy = y0 + p.tap_dy[o];
                            for (int cc = 0; cc < p.cpt; ++cc) issue_a(c0 + cc * TG_KC, xx, yy);
                        }
                    } else {
                        for (int it = 0; it < k_iters; ++it) issue_b(it * TG_KC, bz);
                    }
                }
            }
            if (dbg_on && blockIdx.x == 0 && lane == 0) {
                p.dbg[load_a ? 10 : 11] = w_issue;
                p.dbg[load_a ? 0 : 8] = clock64() - t_begin;
                p.dbg[load_a ? 1 : 9] = w_empty;
            }
        }
    };
    auto mma_role = [&]() {
      // ===================================================== MMA issuer (one thread; with CG == 2 only in the leader CTA)
      if (rank == 0) {
        const uint32_t idesc = make_idesc_f16(TG_BM * CG, static_cast<uint32_t>(p.BN), BF16 ? 1u : 0u) |
                               (p.a_mn_major ? IDESC_A_MN_MAJOR : 0u);
        const int n_sub = p.reuse ? 3 : 1;                        // vertical taps served by one stage
        const int kk_last = (p.kk_last > 0 && p.kk_last < TG_KC / 16) ? p.kk_last : TG_KC / 16;
        // The issuing thread is the busy role of every narrow-N launch (~4 cycles per dependent uniform instruction), so
        // the loop keeps running descriptors: everything that moves between MMAs is the 14-bit address field (16-byte
        // units) in the low word of a descriptor, advanced by plain 64-bit adds. Shared memory ends below 256 KB, so the
        // field never carries into its neighbours.
        // K-major A: 16 elements along K are 32 bytes inside the swizzle atom (+2); M-major A: 16 K rows are two
        // 1024-byte atoms (+128)
        const uint32_t a_step = p.a_mn_major ? 128 : 2;
        const uint64_t a_desc0 = p.a_mn_major ? make_mnmajor_desc(smem_u32(smA), 8192, 1024) : make_kmajor_desc(smem_u32(smA), 128);
        const uint64_t b_desc0 = make_kmajor_desc(smem_u32(smB), 128);
        const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
        const uint32_t a_lo0 = static_cast<uint32_t>(a_desc0), b_lo0 = static_cast<uint32_t>(b_desc0);
        const uint32_t a_stage16 = a_stage_bytes >> 4, b_stage16 = b_stage_bytes >> 4;
        const uint32_t a_sub16 = (static_cast<uint32_t>(p.w_box) * 128u) >> 4;  // one image row of the patch
        const uint32_t b_box16 = b_box_bytes >> 4;
        const int skip_mma = dbg_skip & 4;
        int stage = 0;
        uint32_t phase = 0;
        uint32_t a_cur = a_lo0, b_cur = b_lo0;  // low descriptor words of the current stage
        // patch schedule: 8-row groups of the A operand are one patch row (w_box + 2 pixels) apart
        const uint32_t patch_row16 = static_cast<uint32_t>(p.w_box + 2) * 8u;
        const uint32_t a_hi_patch = static_cast<uint32_t>(make_kmajor_desc_sbo(0, static_cast<uint32_t>(p.w_box + 2) * 128u) >> 32);
        int a_st = 0;
        uint32_t a_ph = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long w_tempty = 0, w_full = 0;
        const long long t_begin = clock64();
        int res_n_tile = -1;
        uint32_t res_phase = 0;
        for (int t = first_tile; t < tile_end; t += tile_step) {
            bool res_last = false;  // this tile is the last one that reads the resident weights of its N tile
            if (b_res) {
                const int nt = decode_tile(p, t, CG, rank).n_tile;
                if (nt != res_n_tile) {
                    res_n_tile = nt;
                    mbar_wait_spin(bres_full, res_phase);
                    res_phase ^= 1;
                    tc_fence_after();
                }
                res_last = (t + tile_step < tile_end) && decode_tile(p, t + tile_step, CG, rank).n_tile != nt;
            }
            long long w0 = dbg_on ? clock64() : 0;
            mbar_wait_spin(&tempty_bar[acc], acc_phase ^ 1);
            if (dbg_on) w_tempty += clock64() - w0;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * acc_cols;
            if (patch) {
                // patch schedule: per 64-channel chunk ONE activation box (halo patch, rows of w_box + 2 pixels) serves the
                // nine taps through descriptor start offsets; weights come three taps (one vertical offset) per stage, or
                // are resident. Tap (dy, dx) of output pixel (y, x) reads patch row (y + dy + 1) * (w_box + 2) + x + dx + 1.
                for (int cc2 = 0; cc2 < p.cpt; ++cc2) {
                    const int n_k = (cc2 == p.cpt - 1) ? kk_last : TG_KC / 16;
                    w0 = dbg_on ? clock64() : 0;
                    if (XF) {  // the transform warps of both CTAs have rewritten the box (generic proxy + fence.proxy.async)
                        mbar_wait_spin(&aready_bar[a_st], a_ph);
                    } else {
                        mbar_wait_spin(&afull_bar[a_st], a_ph);
                    }
                    if (dbg_on) w_full += clock64() - w0;
                    // vertical offsets per synchronisation point: all three when the weight stage carries nine taps (narrow N:
                    // the per-stage waits, fence and election then cost as much as the MMAs themselves), else one
                    const int d_per = (b_res || p.b_taps == 9) ? 3 : 1;
                    for (int d0 = 0; d0 < 3; d0 += d_per) {
                        if (!b_res) {
                            w0 = dbg_on ? clock64() : 0;
                            mbar_wait_spin(&full_bar[stage], phase);
                            if (dbg_on) w_full += clock64() - w0;
                        }
                        tc_fence_after();
                        if (elect_one()) {
                            if (!skip_mma) {
                                for (int dd = 0; dd < d_per; ++dd) {
                                    const int d = d0 + dd;
                                    const uint32_t ad = a_lo0 + static_cast<uint32_t>(a_st) * a_stage16 + static_cast<uint32_t>(d) * patch_row16;
                                    const uint32_t bd = b_res ? b_lo0 + static_cast<uint32_t>(cc2 * 9 + d * 3) * b_box16
                                                              : b_cur + static_cast<uint32_t>(dd * 3) * b_box16;
                                    const uint32_t first = (cc2 > 0 || d > 0) ? 1u : 0u;
                                    if (CG == 2) {
                                        umma_f16_run_2sm(d_tmem, ad, a_hi_patch, bd, b_hi, idesc, first, 2u, n_k);
                                        umma_f16_run_2sm(d_tmem, ad + 8u, a_hi_patch, bd + b_box16, b_hi, idesc, 1u, 2u, n_k);
                                        umma_f16_run_2sm(d_tmem, ad + 16u, a_hi_patch, bd + 2u * b_box16, b_hi, idesc, 1u, 2u, n_k);
                                    } else {
                                        umma_f16_run(d_tmem, ad, a_hi_patch, bd, b_hi, idesc, first, 2u, n_k);
                                        umma_f16_run(d_tmem, ad + 8u, a_hi_patch, bd + b_box16, b_hi, idesc, 1u, 2u, n_k);
                                        umma_f16_run(d_tmem, ad + 16u, a_hi_patch, bd + 2u * b_box16, b_hi, idesc, 1u, 2u, n_k);
                                    }
                                }
                            }
                            const bool d_last = d0 + d_per == 3;
                            const bool last = (cc2 == p.cpt - 1) && d_last;
                            if (CG == 2) {
                                if (!b_res) umma_commit_2sm(&empty_bar[stage]);
                                if (d_last) umma_commit_2sm(&aempty_bar[a_st]);
                                if (last) {
                                    umma_commit_2sm(&tfull_bar[acc]);
                                    if (res_last) umma_commit_2sm(tdone_bar);
                                }
                            } else {
                                if (!b_res) umma_commit(&empty_bar[stage]);
                                if (d_last) umma_commit(&aempty_bar[a_st]);
                                if (last) {
                                    umma_commit(&tfull_bar[acc]);
                                    if (res_last) umma_commit(tdone_bar);
                                }
                            }
                        }
                        __syncwarp();
                        if (!b_res) {
                            b_cur += b_stage16;
                            if (++stage == stages) {
                                stage = 0;
                                phase ^= 1;
                                b_cur = b_lo0;
                            }
                        }
                    }
                    if (++a_st == a_stages) {
                        a_st = 0;
                        a_ph ^= 1;
                    }
                }
                if (++acc == acc_stages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            int cc = 0;  // chunk index within the tap (the last chunk of a tap may hold fewer than 64 useful channels)
            const int k_walk = p.kc_cnt ? __ldg(p.kc_cnt + decode_tile(p, t, CG, rank).n_tile) : k_iters;
            for (int kc = 0; kc < k_walk; ++kc) {
                const int n_k = (cc == p.cpt - 1) ? kk_last : TG_KC / 16;
                if (++cc == p.cpt) cc = 0;
                w0 = dbg_on ? clock64() : 0;
                mbar_wait_spin(&full_bar[stage], phase);
                if (dbg_on) w_full += clock64() - w0;
                tc_fence_after();
                if (elect_one()) {
                    uint32_t ad = a_cur;
                    uint32_t bd = b_res ? b_lo0 + static_cast<uint32_t>(kc) * 3u * b_box16 : b_cur;
                    if (!skip_mma) {
                        // 16-channel steps that only meet zero-padded weights are skipped (n_k < 4)
                        if (CG == 2) umma_f16_run_2sm(d_tmem, ad, a_hi, bd, b_hi, idesc, kc > 0 ? 1u : 0u, a_step, n_k);
                        else umma_f16_run(d_tmem, ad, a_hi, bd, b_hi, idesc, kc > 0 ? 1u : 0u, a_step, n_k);
                        if (n_sub == 3) {  // the other two vertical taps: next image row of the patch, next weight box
                            ad += a_sub16; bd += b_box16;
                            if (CG == 2) umma_f16_run_2sm(d_tmem, ad, a_hi, bd, b_hi, idesc, 1u, a_step, n_k);
                            else umma_f16_run(d_tmem, ad, a_hi, bd, b_hi, idesc, 1u, a_step, n_k);
                            ad += a_sub16; bd += b_box16;
                            if (CG == 2) umma_f16_run_2sm(d_tmem, ad, a_hi, bd, b_hi, idesc, 1u, a_step, n_k);
                            else umma_f16_run(d_tmem, ad, a_hi, bd, b_hi, idesc, 1u, a_step, n_k);
                        }
                    }
                    if (CG == 2) {  // release the stage / publish the accumulator in both CTAs of the pair
                        umma_commit_2sm(&empty_bar[stage]);
                        if (kc == k_walk - 1) umma_commit_2sm(&tfull_bar[acc]);
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kc == k_walk - 1) {
                            umma_commit(&tfull_bar[acc]);
                            if (res_last) umma_commit(tdone_bar);
                        }
                    }
                }
                __syncwarp();
                a_cur += a_stage16;
                b_cur += b_stage16;
                if (++stage == stages) {
                    stage = 0;
                    phase ^= 1;
                    a_cur = a_lo0;
                    b_cur = b_lo0;
                }
            }
            if (++acc == acc_stages) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (dbg_on && blockIdx.x == 0 && lane == 0) {
            p.dbg[2] = clock64() - t_begin;
            p.dbg[3] = w_tempty;
            p.dbg[4] = w_full;
        }
      }
    };
    auto xform_role = [&]() {
        // ===================================================== operand transform (4 warps, patch schedule only): GroupNorm
        // (+SiLU) of the raw activation box in shared memory, between the TMA landing and the MMAs. A box is
        // (h_box + 2) x (w_box + 2) = 18 x 10 pixel rows of 128 bytes (64 channels) under the 128-byte swizzle: thread t
        // owns the 16-byte chunk of channels 8 * (t & 7) .. + 7 in rows t >> 3, + 16, ... Pixels outside the image keep the
        // TMA's zero fill: the reference pads the NORMALISED activation (resnet.py:409, 425 padding = 1).
        const int tx = threadIdx.x - W_XF0 * 32;  // 0 .. 255
        const int j8 = tx & 7;
        constexpr int PROW = 10;                  // patch row pitch in pixels (w_box + 2, w_box == 8 in the patch schedule)
        constexpr int RSTEP = TG_XF_THREADS / 8;  // rows between two chunks of a thread (32)
        constexpr int RITER = (18 * PROW + RSTEP - 1) / RSTEP;  // 6
        int a_st = 0;
        uint32_t a_ph = 0;
        // mean / rstd of every (sample, group): double-precision math on the integer sums exactly as in gn_apply_k, once per
        // CTA (the statistics are complete before the launch) into a shared-memory table; plans with more than
        // XF_TABLE (sample, group) pairs compute them per box instead
        constexpr int XF_TABLE = 512;
        const int n_groups = p.a_C / p.ax_cpg;
        const int n_pairs = p.a_B * n_groups;
        const bool use_table = n_pairs <= XF_TABLE;
        const double n_stat = static_cast<double>(p.ax_cpg) * static_cast<double>(p.ax_hw);
        auto finalize = [&](int pair, float& mean, float& rstd) {
            const acc_t* st = p.ax_stats + static_cast<long long>(pair) * 2;
            const double m = static_cast<double>(__ldg(st)) * (1.0 / WFD_FX_STATS) / n_stat;
            double var = static_cast<double>(__ldg(st + 1)) * (1.0 / WFD_FX_STATS) / n_stat - m * m;
            if (var < 0) var = 0;
            mean = static_cast<float>(m);
            rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.ax_eps)));
        };
        if (use_table) {
            for (int i = tx; i < n_pairs; i += TG_XF_THREADS) {
                float mean, rstd;
                finalize(i, mean, rstd);
                xf_s[2 * i] = mean;
                xf_s[2 * i + 1] = rstd;
            }
            asm volatile("bar.sync 4, %0;" ::"n"(TG_XF_THREADS) : "memory");
        }
        const uint32_t smA_u32 = smem_u32(smA);
        long long x_wait = 0, x_math = 0;
        const long long x_begin = clock64();
        // this thread's rows of a box never change: row (tx >> 3) + 16 i, its patch coordinates, its swizzled byte offset
        const int r0 = tx >> 3;
        uint32_t rmask = 0;  // rows that exist (the last of the twelve only for the first 4 row lanes)
#pragma unroll
        for (int i = 0; i < RITER; ++i)
            if (r0 + i * RSTEP < 18 * PROW) rmask |= 1u << i;
        // (r & 7) == (r0 & 7) for every row of the thread (32-row steps): one swizzled chunk offset serves them all
        const uint32_t chunk_off = static_cast<uint32_t>(r0 * 128 + ((j8 ^ (r0 & 7)) << 4));
        int ch_cached = -1;
        float ga[8], be[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ga[j] = be[j] = 0.f;
        for (int base = first_tile; base < tile_end; base += 32 * tile_step) {
            const int my_t = base + lane * tile_step;
            const TileCoord mine = decode_tile(p, my_t < tile_end ? my_t : base, CG, rank);
            int my_c0 = 0;
            if (p.a_c0 && my_t < tile_end) my_c0 = __ldg(p.a_c0 + mine.n_tile);
            const int batch = min(32, (tile_end - base + tile_step - 1) / tile_step);
            for (int l = 0; l < batch; ++l) {
                const int bb = __shfl_sync(0xffffffffu, mine.bb, l);
                const int ty = __shfl_sync(0xffffffffu, mine.ty, l);
                const int txx = __shfl_sync(0xffffffffu, mine.tx, l);
                const int c0 = __shfl_sync(0xffffffffu, my_c0, l);
                const int x_first = txx * p.w_box - 1;
                const int y_first = ty * p.h_box + p.a_y_off - 1;
                // rows of the box that are image pixels (bit i: row r0 + 16 i); interior tiles keep every row
                uint32_t valid = rmask;
                if (x_first < 0 || x_first + PROW > p.a_W || y_first < p.ax_y0 || y_first + 18 > p.ax_y1) {
                    valid = 0;
#pragma unroll
                    for (int i = 0; i < RITER; ++i) {
                        const int r = r0 + i * RSTEP;
                        const int py = r / PROW, px = r - py * PROW;
                        const int y = y_first + py, x = x_first + px;
                        if (x >= 0 && x < p.a_W && y >= p.ax_y0 && y < p.ax_y1) valid |= 1u << i;
                    }
                    valid &= rmask;
                }
                for (int cc = 0; cc < p.cpt; ++cc) {
                    // scale / shift of this thread's eight channels; gamma / beta are re-read only when the channels change
                    // (consecutive tiles of a CTA mostly differ in the sample, i.e. in mean / rstd only)
                    const int ch = c0 + cc * TG_KC + j8 * 8;
                    float sc[8], sh[8];
                    if (ch < p.a_C) {
                        if (ch != ch_cached) {
                            ch_cached = ch;
                            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ax_gamma + ch));
                            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ax_gamma + ch) + 1);
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ax_beta + ch));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.ax_beta + ch) + 1);
                            ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
                            be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
                        }
                        float mean, rstd;
                        const int pair = bb * n_groups + ch / p.ax_cpg;
                        if (use_table) {
                            const float2 mr = *reinterpret_cast<const float2*>(xf_s + 2 * pair);
                            mean = mr.x;
                            rstd = mr.y;
                        } else {
                            finalize(pair, mean, rstd);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            sc[j] = rstd * ga[j];
                            sh[j] = be[j] - mean * rstd * ga[j];
                        }
                    } else {  // beyond the tensor: the TMA filled zeros and the weights there are zero; keep them zero
                        ch_cached = -1;
#pragma unroll
                        for (int j = 0; j < 8; ++j) sc[j] = sh[j] = 0.f;
                    }
                    const long long xw0 = dbg_on ? clock64() : 0;
                    mbar_wait(&afull_bar[a_st], a_ph);
                    const long long xw1 = dbg_on ? clock64() : 0;
                    const uint32_t box = smA_u32 + static_cast<uint32_t>(a_st) * a_stage_bytes + chunk_off;
                    if (!(dbg_skip & 16)) {
                        // two rows in flight per thread, branch-free math (only the loads and stores are predicated); two
                        // transform warps share a scheduler, so one warp's FMAs fill the issue slots the other's tanh leaves
                        constexpr int RU = 2;
#pragma unroll
                        for (int i0 = 0; i0 < RITER; i0 += RU) {
                            uint32_t w[RU][4];
#pragma unroll
                            for (int u = 0; u < RU; ++u) {
                                w[u][0] = w[u][1] = w[u][2] = w[u][3] = 0u;
                                if (valid & (1u << (i0 + u))) lds128(box + (i0 + u) * RSTEP * 128, w[u]);
                            }
                            uint32_t o[RU][4];
#pragma unroll
                            for (int u = 0; u < RU; ++u) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float v0 = lo16_to_f<BF16>(w[u][j]) * sc[2 * j] + sh[2 * j];
                                    float v1 = hi16_to_f<BF16>(w[u][j]) * sc[2 * j + 1] + sh[2 * j + 1];
                                    if (p.ax_silu) {
                                        v0 = v0 * fast_sigmoid(v0);
                                        v1 = v1 * fast_sigmoid(v1);
                                    }
                                    o[u][j] = pack2<BF16>(v0, v1);
                                }
                            }
#pragma unroll
                            for (int u = 0; u < RU; ++u)
                                if (valid & (1u << (i0 + u))) sts128(box + (i0 + u) * RSTEP * 128, o[u]);
                        }
                    }
                    // the rewritten box must be visible to the tensor core (async proxy) before the MMA thread is told
                    // (a cluster-scope release on the remote arrive compiles to MEMBAR.ALL.GPU - microseconds per box; the
                    // proxy fence already orders this thread's shared-memory writes before the arrive)
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_leader(&aready_bar[a_st]);
                        else mbar_arrive(&aready_bar[a_st]);
                    }
                    if (dbg_on) {
                        x_wait += xw1 - xw0;
                        x_math += clock64() - xw1;
                    }
                    if (++a_st == a_stages) {
                        a_st = 0;
                        a_ph ^= 1;
                    }
                }
            }
        }
        if (dbg_on && blockIdx.x == 0 && tx == 0) {
            p.dbg[12] = clock64() - x_begin;
            p.dbg[13] = x_wait;
            p.dbg[14] = x_math;
        }
    };
    auto epilogue_role = [&]() {
        // ===================================================== epilogue (8 warps: a TMEM lane quarter is shared by two
        // warps that take the even / odd 16-column chunks of the tile)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int half = (warp - W_EPI0) >> 2;    // 0: even chunks, 1: odd chunks (tile-split mode: even / odd tiles)
        // Narrow tiles (a few 16-column chunks) are bound by the per-tile bookkeeping of the epilogue, not by its columns:
        // in tile-split mode each half-set of four warps owns one accumulator stage and drains every other tile on its own
        const bool split = p.epi_split != 0;
        const int etid = split ? (threadIdx.x - W_EPI0 * 32) & 127 : threadIdx.x - W_EPI0 * 32;  // index among the threads that share a tile
        const int eset = split ? TG_EPI_THREADS / 2 : TG_EPI_THREADS;
        const int c_first = split ? 0 : half, c_step = split ? 1 : 2;
        int acc = 0, my_tiles = 0;
        int staged_nt[2] = {-1, -1};  // N tile whose bias / gamma / beta sit in each of this set's two staging buffers
        uint32_t acc_phase = 0;
        long long w_tfull = 0, e_busy = 0;
        const long long t_begin = clock64();
        const int yy = row / p.w_box, xx = row - yy * p.w_box;
        if (EPI == 3) {
            // mean / rstd of every (sample, group) of the saved activation, once per CTA (same arithmetic as the generic loop)
            const int per_sample = p.tiles_x * p.tiles_y;
            const int n_pairs = (p.m_tiles / per_sample) * p.gn_G;
            for (int i = threadIdx.x - W_EPI0 * 32; i < n_pairs; i += TG_EPI_THREADS) {
                const acc_t* st = p.gnb_stats + static_cast<long long>(i) * 2;
                const double dm = static_cast<double>(__ldg(st)) * (1.0 / WFD_FX_STATS) * p.gnb_inv_n;
                const double dv = static_cast<double>(__ldg(st + 1)) * (1.0 / WFD_FX_STATS) * p.gnb_inv_n - dm * dm;
                gnb_tab[2 * i] = static_cast<float>(dm);
                gnb_tab[2 * i + 1] = rsqrtf(fmaxf(static_cast<float>(dv), 0.f) + p.gnb_eps);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TG_EPI_THREADS) : "memory");
        }
        for (int base = first_tile; base < tile_end; base += 32 * tile_step) {
          const int my_t = base + lane * tile_step;
          const TileCoord mine = decode_tile(p, my_t < tile_end ? my_t : base, CG, rank);
          const int batch = min(32, (tile_end - base + tile_step - 1) / tile_step);
          for (int l = 0; l < batch; ++l) {
            TileCoord tc;
            tc.z = __shfl_sync(0xffffffffu, mine.z, l);
            tc.n_tile = __shfl_sync(0xffffffffu, mine.n_tile, l);
            tc.bb = __shfl_sync(0xffffffffu, mine.bb, l);
            tc.ty = __shfl_sync(0xffffffffu, mine.ty, l);
            tc.tx = __shfl_sync(0xffffffffu, mine.tx, l);
            if (split && (acc & 1) != half) {  // the other half-set's tile (stages alternate between the half-sets)
                if (++acc == acc_stages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            const long long pixel =
                (static_cast<long long>(tc.bb) * p.H + tc.ty * p.h_box + yy) * p.W + tc.tx * p.w_box + xx;
            const int n0 = tc.n_tile * p.BN;
            const long long out_off = tc.z * p.c_zstride + tc.bb * p.out_sb +
                                      static_cast<long long>(tc.ty * p.h_box + yy) * p.out_sy +
                                      static_cast<long long>(tc.tx * p.w_box + xx) * p.out_sx;
            // stage this tile's bias in shared memory while the MMAs are still running
            // tile-split mode: a half-set always drains the same accumulator, so its staging buffers alternate with its own
            // tile count instead (four 128-float slices of each 2 x 256 array; BN <= 128 there)
            const int stage_buf = split ? (my_tiles & 1) : (acc & 1);
            const int stage_off = split ? (half * 2 + stage_buf) * 128 : (acc & 1) * 256;
            ++my_tiles;
            float* bs = bias_s + stage_off;
            float* gms = gam_s + stage_off;
            float* bts = bet_s + stage_off;
            // consecutive tiles of a CTA mostly share the N tile (M runs fastest): the per-column vectors are staged again
            // only when it changes. The threads of a set then run unsynchronised between restages, so a restage is
            // bracketed by two barriers (nobody still reads the buffer / everybody sees the new contents).
            if ((p.bias || p.gnb_red) && staged_nt[stage_buf] != tc.n_tile) {
                staged_nt[stage_buf] = tc.n_tile;
                auto set_sync = [&]() {
                    if (split) {
                        if (half == 0) asm volatile("bar.sync 2, %0;" ::"n"(TG_EPI_THREADS / 2) : "memory");
                        else asm volatile("bar.sync 3, %0;" ::"n"(TG_EPI_THREADS / 2) : "memory");
                    } else {
                        asm volatile("bar.sync 1, %0;" ::"n"(TG_EPI_THREADS) : "memory");
                    }
                };
                set_sync();
                if (p.bias) {
                    for (int j = etid; j < p.BN; j += eset) bs[j] = (n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
                }
                if (p.gnb_red) {
                    for (int j = etid; j < p.BN; j += eset) {
                        gms[j] = (n0 + j < p.N) ? __ldg(p.gnb_gamma + n0 + j) : 0.f;
                        bts[j] = (n0 + j < p.N) ? __ldg(p.gnb_beta + n0 + j) : 0.f;
                    }
                }
                set_sync();
            }
            if (EPI == 3) {
                // Specialised drain of the backward convs with the fused GroupNorm(+SiLU) backward reduction (16-bit store of
                // dy, then sum(du*gamma), sum(du*gamma*xhat) against the saved activation x; <= six chunks per warp): same
                // arithmetic as the generic loop without its option tests, gamma / beta as 16-byte shared-memory loads,
                // mean / rstd from the per-CTA table.
                const int n_chunks_f = (min(p.BN, p.N - n0) + 15) / 16;
                const int n_mine = split ? n_chunks_f : (n_chunks_f + 1 - half) / 2;
                uint16_t* outp = reinterpret_cast<uint16_t*>(p.out) + out_off + n0;
                const uint16_t* xp = reinterpret_cast<const uint16_t*>(p.gnb_x) + out_off + n0;
                const bool silu = p.gnb_silu != 0;
                uint4 rs[3][2];
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k < n_mine) ldg256(xp + (c_first + k * c_step) * 16, rs[k][0], rs[k][1]);
                mbar_wait_epi(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc) * acc_cols + (static_cast<uint32_t>(q * 32) << 16);
                const int g_cpg = p.gnb_cpg;
                int g_run = n0 / g_cpg, g_next = (g_run + 1) * g_cpg, bgroup = -1;
                float b1 = 0.f, b2 = 0.f, bmean = 0.f, brstd = 0.f;
                auto flush_red = [&]() {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        b1 += __shfl_xor_sync(0xffffffffu, b1, o);
                        b2 += __shfl_xor_sync(0xffffffffu, b2, o);
                    }
                    if (lane == 0) {
                        const int slot = (tc.bb * p.gn_G + bgroup) * 2;
                        const unsigned long long v0 = static_cast<unsigned long long>(__float2ll_rn(b1 * WFD_FX_GRAD));
                        const unsigned long long v1 = static_cast<unsigned long long>(__float2ll_rn(b2 * WFD_FX_GRAD));
                        if (p.stat_private) {  // this warp's own copy (<= 64 (sample, group) pairs)
                            unsigned long long* sp = stat_s + TG_STAT_SLOTS + (warp - W_EPI0) * 128 + slot;
                            sp[0] += v0;
                            sp[1] += v1;
                        } else {  // <= 256 pairs: the CTA's shared slots
                            atomicAdd(stat_s + TG_STAT_SLOTS + slot, v0);
                            atomicAdd(stat_s + TG_STAT_SLOTS + slot + 1, v1);
                        }
                    }
                };
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    if (k < n_mine) {
                        const int j0 = (c_first + k * c_step) * 16;
                        uint32_t r[16];
                        tmem_ld16(taddr + j0, r);
                        tmem_ld_wait(r);
                        uint32_t w[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = pack2<BF16>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                        stg256(outp + j0, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
                        const int n = n0 + j0;
                        while (n >= g_next) {
                            ++g_run;
                            g_next += g_cpg;
                        }
                        if (g_run != bgroup) {
                            if (bgroup >= 0) flush_red();
                            bgroup = g_run;
                            b1 = 0.f;
                            b2 = 0.f;
                            const float2 mr = *reinterpret_cast<const float2*>(gnb_tab + 2 * (tc.bb * p.gn_G + g_run));
                            bmean = mr.x;
                            brstd = mr.y;
                        }
                        const uint32_t xw[8] = {rs[k % 3][0].x, rs[k % 3][0].y, rs[k % 3][0].z, rs[k % 3][0].w,
                                                rs[k % 3][1].x, rs[k % 3][1].y, rs[k % 3][1].z, rs[k % 3][1].w};
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const float4 g4 = *reinterpret_cast<const float4*>(gms + j0 + 4 * q4);
                            const float4 b4 = *reinterpret_cast<const float4*>(bts + j0 + 4 * q4);
                            const float ga4[4] = {g4.x, g4.y, g4.z, g4.w}, be4[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                            for (int t4 = 0; t4 < 4; ++t4) {
                                const int col = 4 * q4 + t4;
                                const float dyv = (col & 1) ? hi16_to_f<BF16>(w[col >> 1]) : lo16_to_f<BF16>(w[col >> 1]);
                                const float xv = (col & 1) ? hi16_to_f<BF16>(xw[col >> 1]) : lo16_to_f<BF16>(xw[col >> 1]);
                                const float xh = (xv - bmean) * brstd;
                                float du = dyv;
                                if (silu) {
                                    const float u = fmaf(xh, ga4[t4], be4[t4]);
                                    const float sg = fast_sigmoid(u);
                                    du *= sg * fmaf(u, 1.f - sg, 1.f);
                                }
                                const float tt = du * ga4[t4];
                                b1 += tt;
                                b2 = fmaf(tt, xh, b2);
                            }
                        }
                        if (k + 3 < 6 && k + 3 < n_mine) ldg256(xp + (c_first + (k + 3) * c_step) * 16, rs[k % 3][0], rs[k % 3][1]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);
                    else mbar_arrive(&tempty_bar[acc]);
                }
                if (bgroup >= 0) flush_red();
                if (++acc == acc_stages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            if (EPI == 2) {
                // Specialised drain of the attention products whose epilogue IS the softmax work (row statistics /
                // probabilities / softmax backward; 16-bit or no store, up to eight chunks per warp): with K = 384 a
                // 256 x 256 tile is 24 MMAs, and the generic loop's ~260 instructions per chunk made the epilogue the
                // busy role (the MMA thread waited 55-74% of its time for a free accumulator).
                const int n_chunks_f = (min(p.BN, p.N - n0) + 15) / 16;
                const int n_mine = split ? n_chunks_f : (n_chunks_f + 1 - half) / 2;
                const long long grow = pixel + tc.z * static_cast<long long>(p.m_tiles) * TG_BM;
                auto drain_attn = [&](auto MODE_T) {
                    constexpr int MODE = decltype(MODE_T)::value;  // 0: row statistics, 1: probabilities, 2: softmax backward
                    uint16_t* outp = MODE == 0 ? nullptr : reinterpret_cast<uint16_t*>(p.out) + out_off + n0;
                    const uint16_t* pp = MODE == 2 ? reinterpret_cast<const uint16_t*>(p.residual) + out_off + n0 : nullptr;
                    const float alpha = p.alpha;
                    float2 st = make_float2(0.f, 0.f);
                    float drow = 0.f, sscale = 0.f;
                    if (MODE == 1) st = __ldg(p.rs_in + grow);
                    if (MODE == 2) {
                        drow = __ldg(p.sb_d + grow);
                        sscale = p.sb_scale;
                    }
                    uint4 rs[3][2];
                    if (MODE == 2) {
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (k < n_mine) ldg256(pp + (c_first + k * c_step) * 16, rs[k][0], rs[k][1]);
                    }
                    mbar_wait_epi(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc) * acc_cols + (static_cast<uint32_t>(q * 32) << 16);
                    float rs_m = -INFINITY, rs_l = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (k < n_mine) {
                            const int j0 = (c_first + k * c_step) * 16;
                            uint32_t r[16];
                            tmem_ld16(taddr + j0, r);
                            tmem_ld_wait(r);
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * alpha;
                            if (MODE == 0) {
                                float cm = v[0];
#pragma unroll
                                for (int j = 1; j < 16; ++j) cm = fmaxf(cm, v[j]);
                                if (cm > rs_m) {
                                    rs_l *= __expf(rs_m - cm);
                                    rs_m = cm;
                                }
#pragma unroll
                                for (int j = 0; j < 16; ++j) rs_l += __expf(v[j] - rs_m);
                            } else {
                                if (MODE == 1) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[j] = __expf(v[j] - st.x) * st.y;
                                } else {
#pragma unroll
                                    for (int h2 = 0; h2 < 2; ++h2) {
                                        const uint32_t w4[4] = {rs[k % 3][h2].x, rs[k % 3][h2].y, rs[k % 3][h2].z, rs[k % 3][h2].w};
#pragma unroll
                                        for (int j = 0; j < 4; ++j) {
                                            v[h2 * 8 + 2 * j] = lo16_to_f<BF16>(w4[j]) * (v[h2 * 8 + 2 * j] - drow) * sscale;
                                            v[h2 * 8 + 2 * j + 1] = hi16_to_f<BF16>(w4[j]) * (v[h2 * 8 + 2 * j + 1] - drow) * sscale;
                                        }
                                    }
                                    if (k + 3 < 8 && k + 3 < n_mine)
                                        ldg256(pp + (c_first + (k + 3) * c_step) * 16, rs[k % 3][0], rs[k % 3][1]);
                                }
                                uint32_t w[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) w[j] = pack2<BF16>(v[2 * j], v[2 * j + 1]);
                                stg256(outp + j0, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
                            }
                        }
                    }
                    if (MODE == 0) {
                        const long long slot = (grow * p.n_tiles + tc.n_tile) * 2;
                        if (split) {
                            p.rs_out[slot] = make_float2(rs_m, rs_l);
                            p.rs_out[slot + 1] = make_float2(-INFINITY, 0.f);
                        } else {
                            p.rs_out[slot + half] = make_float2(rs_m, rs_l);
                        }
                    }
                };
                if (p.rs_out) drain_attn(std::integral_constant<int, 0>{});
                else if (p.rs_in) drain_attn(std::integral_constant<int, 1>{});
                else drain_attn(std::integral_constant<int, 2>{});
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);
                    else mbar_arrive(&tempty_bar[acc]);
                }
                if (++acc == acc_stages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            if (EPI == 1) {
                // Compile-time specialised drain of the plain conv cases (16-bit store, 256-bit accesses, alpha = 1; optional
                // bias, residual, fused GroupNorm statistics; at most six chunks per warp). The generic loop below spends
                // ~260 instructions per 16-column chunk and ~360 per tile on option tests, address arithmetic and the
                // rotation of its side-input registers; the forward groups=64 convs (residual + statistics) were bound by
                // exactly that (the MMA thread waited 56% of its time for a free accumulator). Same arithmetic, same order.
                const int n_chunks_f = (min(p.BN, p.N - n0) + 15) / 16;
                const int n_mine = split ? n_chunks_f : (n_chunks_f + 1 - half) / 2;
                uint16_t* outp = reinterpret_cast<uint16_t*>(p.out) + out_off + n0;
                float gs = 0.f, gss = 0.f;
                int ggroup = -1;
                auto flush_stats = [&]() {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        gs += __shfl_xor_sync(0xffffffffu, gs, o);
                        gss += __shfl_xor_sync(0xffffffffu, gss, o);
                    }
                    if (lane == 0) {
                        const int slot = (tc.bb * p.gn_G + ggroup) * 2;
                        const unsigned long long v0 = static_cast<unsigned long long>(__float2ll_rn(gs * WFD_FX_STATS));
                        const unsigned long long v1 = static_cast<unsigned long long>(__float2ll_rn(gss * WFD_FX_STATS));
                        if (p.stat_private) {
                            // this warp's own copy of the slots (only its lane 0 touches it): plain load + store instead of
                            // a 64-bit shared-memory atomic, which is a compare-and-swap loop that the eight epilogue
                            // warps of a CTA fight over (they mostly hit the same sample and group)
                            unsigned long long* sp = stat_s + (warp - W_EPI0) * 128 + slot;
                            sp[0] += v0;
                            sp[1] += v1;
                        } else if (slot + 1 < TG_STAT_SLOTS) {
                            atomicAdd(stat_s + slot, v0);
                            atomicAdd(stat_s + slot + 1, v1);
                        } else {
                            atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) + slot, v0);
                            atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) + slot + 1, v1);
                        }
                    }
                };
                auto drain = [&](auto HAS_BIAS, auto HAS_RES, auto HAS_GNS) {
                    constexpr bool BIAS = decltype(HAS_BIAS)::value, RES = decltype(HAS_RES)::value, GNS = decltype(HAS_GNS)::value;
                    const uint16_t* resp = RES ? reinterpret_cast<const uint16_t*>(p.residual) + out_off + n0 : nullptr;
                    uint4 rs[3][2];
                    if (RES) {  // the first three residual pieces of this warp are in flight before the accumulator wait
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (k < n_mine) ldg256(resp + (c_first + k * c_step) * 16, rs[k][0], rs[k][1]);
                    }
                    mbar_wait_epi(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc) * acc_cols + (static_cast<uint32_t>(q * 32) << 16);
                    const int g_cpg = GNS ? p.gn_cpg : 1;
                    int g_run = GNS ? n0 / g_cpg : 0;
                    int g_next = (g_run + 1) * g_cpg;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        if (k < n_mine) {
                            const int j0 = (c_first + k * c_step) * 16;
                            uint32_t r[16];
                            tmem_ld16(taddr + j0, r);
                            tmem_ld_wait(r);
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                            if (BIAS) {
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    const float4 b4 = *reinterpret_cast<const float4*>(bs + j0 + j);
                                    v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                                }
                            }
                            if (RES) {
#pragma unroll
                                for (int h2 = 0; h2 < 2; ++h2) {
                                    const uint32_t w4[4] = {rs[k % 3][h2].x, rs[k % 3][h2].y, rs[k % 3][h2].z, rs[k % 3][h2].w};
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        v[h2 * 8 + 2 * j] += lo16_to_f<BF16>(w4[j]);
                                        v[h2 * 8 + 2 * j + 1] += hi16_to_f<BF16>(w4[j]);
                                    }
                                }
                                // the slot is free: the piece three chunks ahead takes it
                                if (k + 3 < 6 && k + 3 < n_mine)
                                    ldg256(resp + (c_first + (k + 3) * c_step) * 16, rs[k % 3][0], rs[k % 3][1]);
                            }
                            uint32_t w[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) w[j] = pack2<BF16>(v[2 * j], v[2 * j + 1]);
                            stg256(outp + j0, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
                            if (GNS) {
                                const int n = n0 + j0;
                                while (n >= g_next) {
                                    ++g_run;
                                    g_next += g_cpg;
                                }
                                if (g_run != ggroup) {
                                    if (ggroup >= 0) flush_stats();
                                    ggroup = g_run;
                                    gs = 0.f;
                                    gss = 0.f;
                                }
                                // statistics of the values before their 16-bit rounding (what the reference's fp32
                                // GroupNorm sees; measured equally fast as statistics of the re-unpacked stores)
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    gs += v[j];
                                    gss = fmaf(v[j], v[j], gss);
                                }
                            }
                        }
                    }
                };
                using T = std::true_type;
                using F = std::false_type;
                const int mode = (p.bias ? 4 : 0) | (p.residual ? 2 : 0) | (p.gn_sums ? 1 : 0);
                switch (mode) {  // (the combinations the decoder launches; tapgemm_prepare admits no others)
                    case 0: drain(F{}, F{}, F{}); break;
                    case 4: drain(T{}, F{}, F{}); break;
                    case 5: drain(T{}, F{}, T{}); break;
                    default: drain(T{}, T{}, T{}); break;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);
                    else mbar_arrive(&tempty_bar[acc]);
                }
                if (ggroup >= 0) flush_stats();
                if (++acc == acc_stages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            // one row-wise side input per launch: residual (forward), dot operand (WFC) or saved activation (GroupNorm
            // backward). Its 32-byte pieces are prefetched three chunks deep, the first three before the accumulator wait,
            // so that DRAM latency overlaps the MMAs instead of serialising the column loop.
            const int side_kind = p.residual ? 1 : (p.rowdot ? 2 : (p.gnb_red ? 3 : 0));
            const uint4* side_ptr = nullptr;
            if (side_kind == 1)
                side_ptr = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + out_off + n0);
            else if (side_kind == 2)
                side_ptr = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.dotsrc) + pixel * p.ld_dot + n0);
            else if (side_kind == 3)
                side_ptr = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.gnb_x) + out_off + n0);
            const int n_chunks = (min(p.BN, p.N - n0) + 15) / 16;
            const float sb_row = p.sb_d ? __ldg(p.sb_d + pixel + tc.z * static_cast<long long>(p.m_tiles) * TG_BM) : 0.f;
            uint4 sb0[2], sb1[2], sb2[2];
            auto load_side = [&](int chunk, uint4(&dst)[2]) {
                if (p.wide_io) {
                    ldg256(side_ptr + 2 * chunk, dst[0], dst[1]);
                } else {
                    dst[0] = __ldg(side_ptr + 2 * chunk);
                    dst[1] = __ldg(side_ptr + 2 * chunk + 1);
                }
            };
            if (side_ptr && !(dbg_skip & 8)) {
                if (c_first < n_chunks) load_side(c_first, sb0);
                if (c_first + c_step < n_chunks) load_side(c_first + c_step, sb1);
                if (c_first + 2 * c_step < n_chunks) load_side(c_first + 2 * c_step, sb2);
            }
            {
                const long long w0 = dbg_on ? clock64() : 0;
                mbar_wait_epi(&tfull_bar[acc], acc_phase);
                if (dbg_on) w_tfull += clock64() - w0;
            }
            tc_fence_after();
            const long long e0 = dbg_on ? clock64() : 0;
            const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc) * acc_cols + (static_cast<uint32_t>(q * 32) << 16);
            float dot = 0.f;
            float rs_m = -INFINITY, rs_l = 0.f;                       // pass 1 of the row softmax
            const float2 rs_row = p.rs_in ? __ldg(p.rs_in + pixel + tc.z * static_cast<long long>(p.m_tiles) * TG_BM) : make_float2(0.f, 0.f);
            float gs = 0.f, gss = 0.f;
            int ggroup = -1;
            // GroupNorm group of a 16-column chunk without a division per chunk: the group of the tile's first column
            // (one division per tile) and the first column of the next group; chunks are visited in increasing order
            const int g_cpg = p.gn_sums ? p.gn_cpg : (p.gnb_red ? p.gnb_cpg : 0);
            int g_run = g_cpg ? n0 / g_cpg : 0;
            int g_next = g_cpg ? (g_run + 1) * g_cpg : 0x7fffffff;
            float b1 = 0.f, b2 = 0.f, bmean = 0.f, brstd = 0.f;
            int bgroup = -1;
            for (int c = c_first; c < ((dbg_skip & 8) ? 0 : n_chunks); c += c_step) {
                const int j0 = c * 16;
                const int n = n0 + j0;
                uint4 side_cur[2];
                if (side_ptr) {
                    side_cur[0] = sb0[0]; side_cur[1] = sb0[1];
                    sb0[0] = sb1[0]; sb0[1] = sb1[1];
                    sb1[0] = sb2[0]; sb1[1] = sb2[1];
                    if (c + 3 * c_step < n_chunks) load_side(c + 3 * c_step, sb2);
                }
                uint32_t r[16];
                tmem_ld16(taddr + j0, r);
                tmem_ld_wait(r);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bs + j0 + j);
                        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                    }
                }
                if (side_kind == 1 && p.sb_d) {  // softmax backward: dS = P * (dP - rowdot) * scale
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w[4] = {side_cur[h].x, side_cur[h].y, side_cur[h].z, side_cur[h].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[h * 8 + 2 * j] = lo16_to_f<BF16>(w[j]) * (v[h * 8 + 2 * j] - sb_row) * p.sb_scale;
                            v[h * 8 + 2 * j + 1] = hi16_to_f<BF16>(w[j]) * (v[h * 8 + 2 * j + 1] - sb_row) * p.sb_scale;
                        }
                    }
                } else if (side_kind == 1) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w[4] = {side_cur[h].x, side_cur[h].y, side_cur[h].z, side_cur[h].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[h * 8 + 2 * j] += lo16_to_f<BF16>(w[j]);
                            v[h * 8 + 2 * j + 1] += hi16_to_f<BF16>(w[j]);
                        }
                    }
                }
                if (side_kind == 2) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w[4] = {side_cur[h].x, side_cur[h].y, side_cur[h].z, side_cur[h].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dot += v[h * 8 + 2 * j] * lo16_to_f<BF16>(w[j]);
                            dot += v[h * 8 + 2 * j + 1] * hi16_to_f<BF16>(w[j]);
                        }
                    }
                }
                if (p.rs_out) {  // row softmax, pass 1: statistics only
                    float cm = v[0];
#pragma unroll
                    for (int j = 1; j < 16; ++j) cm = fmaxf(cm, v[j]);
                    if (cm > rs_m) {
                        rs_l *= __expf(rs_m - cm);
                        rs_m = cm;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) rs_l += __expf(v[j] - rs_m);
                    continue;
                }
                if (p.rs_in) {   // row softmax, pass 2: probabilities
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __expf(v[j] - rs_row.x) * rs_row.y;
                }
                if (p.out_fp32) {
                    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + n);
                    if (p.wide_io) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            stg256(op + 2 * j,
                                   make_uint4(__float_as_uint(v[8 * j]), __float_as_uint(v[8 * j + 1]), __float_as_uint(v[8 * j + 2]), __float_as_uint(v[8 * j + 3])),
                                   make_uint4(__float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]), __float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7])));
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                } else {
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = pack2<BF16>(v[2 * j], v[2 * j + 1]);
                    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + out_off + n);
                    if (p.wide_io) {
                        stg256(op, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
                    } else {
                        op[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        op[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                    if (g_cpg) {
                        while (n >= g_next) {
                            ++g_run;
                            g_next += g_cpg;
                        }
                    }
                    if (p.gn_sums) {
                        const int g = g_run;
                        if (g != ggroup) {
                            if (ggroup >= 0) {
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    gs += __shfl_xor_sync(0xffffffffu, gs, o);
                                    gss += __shfl_xor_sync(0xffffffffu, gss, o);
                                }
                                if (lane == 0) {
                                    const int slot = (tc.bb * p.gn_G + ggroup) * 2;
                                    if (slot + 1 < TG_STAT_SLOTS) {
                                        atomicAdd(stat_s + slot, static_cast<unsigned long long>(__float2ll_rn((gs) * WFD_FX_STATS)));
                                        atomicAdd(stat_s + slot + 1, static_cast<unsigned long long>(__float2ll_rn((gss) * WFD_FX_STATS)));
                                    } else {
                                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) + slot, static_cast<unsigned long long>(__float2ll_rn((gs) * WFD_FX_STATS)));
                                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) + slot + 1, static_cast<unsigned long long>(__float2ll_rn((gss) * WFD_FX_STATS)));
                                    }
                                }
                            }
                            ggroup = g;
                            gs = 0.f;
                            gss = 0.f;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float a = lo16_to_f<BF16>(w[j]), b = hi16_to_f<BF16>(w[j]);
                            gs += a + b;
                            gss += a * a + b * b;
                        }
                    }
                    if (side_kind == 3) {
                        // fused GroupNorm(+SiLU) backward reduction over the just-stored gradient dy and the saved
                        // forward activation x: sum(du*gamma), sum(du*gamma*xhat), du = dy * act'(xhat*gamma+beta)
                        const int g = g_run;
                        if (g != bgroup) {
                            if (bgroup >= 0) {
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    b1 += __shfl_xor_sync(0xffffffffu, b1, o);
                                    b2 += __shfl_xor_sync(0xffffffffu, b2, o);
                                }
                                if (lane == 0) {
                                    const int slot = (tc.bb * p.gn_G + bgroup) * 2;
                                    if (slot + 1 < TG_STAT_SLOTS) {
                                        atomicAdd(stat_s + TG_STAT_SLOTS + slot, static_cast<unsigned long long>(__float2ll_rn((b1) * WFD_FX_GRAD)));
                                        atomicAdd(stat_s + TG_STAT_SLOTS + slot + 1, static_cast<unsigned long long>(__float2ll_rn((b2) * WFD_FX_GRAD)));
                                    } else {
                                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gnb_red) + slot, static_cast<unsigned long long>(__float2ll_rn((b1) * WFD_FX_GRAD)));
                                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gnb_red) + slot + 1, static_cast<unsigned long long>(__float2ll_rn((b2) * WFD_FX_GRAD)));
                                    }
                                }
                            }
                            bgroup = g;
                            b1 = 0.f;
                            b2 = 0.f;
                            const acc_t* st = p.gnb_stats + (static_cast<long long>(tc.bb) * p.gn_G + g) * 2;
                            const double dm = static_cast<double>(__ldg(st)) * (1.0 / WFD_FX_STATS) * p.gnb_inv_n;
                            const double dv = static_cast<double>(__ldg(st + 1)) * (1.0 / WFD_FX_STATS) * p.gnb_inv_n - dm * dm;
                            const float m = static_cast<float>(dm);
                            const float var = fmaxf(static_cast<float>(dv), 0.f);
                            bmean = m;
                            brstd = rsqrtf(var + p.gnb_eps);
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t xw[4] = {side_cur[h].x, side_cur[h].y, side_cur[h].z, side_cur[h].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int col = h * 8 + 2 * j + e;
                                    const float dyv = e ? hi16_to_f<BF16>(w[h * 4 + j]) : lo16_to_f<BF16>(w[h * 4 + j]);
                                    const float xv = e ? hi16_to_f<BF16>(xw[j]) : lo16_to_f<BF16>(xw[j]);
                                    const float xh = (xv - bmean) * brstd;
                                    const float ga = gms[j0 + col];
                                    float du = dyv;
                                    if (p.gnb_silu) {
                                        const float u = fmaf(xh, ga, bts[j0 + col]);
                                        const float sg = fast_sigmoid(u);
                                        du *= sg * fmaf(u, 1.f - sg, 1.f);
                                    }
                                    const float tt = du * ga;
                                    b1 += tt;
                                    b2 = fmaf(tt, xh, b2);
                                }
                            }
                        }
                    }
                }
            }
            // the accumulator stage is drained: hand it back to the MMA warp before the remaining bookkeeping
            tc_fence_before();
            __syncwarp();  // every lane's TMEM reads are complete: one lane hands the warp's share back
            if (lane == 0) {
                if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);  // the MMA thread lives in the leader CTA
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (p.rs_out) {
                // one partial per (row, N tile, half): in chunk-split mode the two warps of a lane quarter saw the even /
                // odd chunks, in tile-split mode the tile's only warp writes slot 0 and an empty slot 1
                const long long slot = ((pixel + tc.z * static_cast<long long>(p.m_tiles) * TG_BM) * p.n_tiles + tc.n_tile) * 2;
                if (split) {
                    p.rs_out[slot] = make_float2(rs_m, rs_l);
                    p.rs_out[slot + 1] = make_float2(-INFINITY, 0.f);
                } else {
                    p.rs_out[slot + half] = make_float2(rs_m, rs_l);
                }
            }
            if (p.gn_sums && ggroup >= 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    gs += __shfl_xor_sync(0xffffffffu, gs, o);
                    gss += __shfl_xor_sync(0xffffffffu, gss, o);
                }
                if (lane == 0) {
                    const int slot = (tc.bb * p.gn_G + ggroup) * 2;
                    if (slot + 1 < TG_STAT_SLOTS) {
                        atomicAdd(stat_s + slot, static_cast<unsigned long long>(__float2ll_rn((gs) * WFD_FX_STATS)));
                        atomicAdd(stat_s + slot + 1, static_cast<unsigned long long>(__float2ll_rn((gss) * WFD_FX_STATS)));
                    } else {
                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) + slot, static_cast<unsigned long long>(__float2ll_rn((gs) * WFD_FX_STATS)));
                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) + slot + 1, static_cast<unsigned long long>(__float2ll_rn((gss) * WFD_FX_STATS)));
                    }
                }
            }
            if (p.gnb_red && bgroup >= 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    b1 += __shfl_xor_sync(0xffffffffu, b1, o);
                    b2 += __shfl_xor_sync(0xffffffffu, b2, o);
                }
                if (lane == 0) {
                    const int slot = (tc.bb * p.gn_G + bgroup) * 2;
                    if (slot + 1 < TG_STAT_SLOTS) {
                        atomicAdd(stat_s + TG_STAT_SLOTS + slot, static_cast<unsigned long long>(__float2ll_rn((b1) * WFD_FX_GRAD)));
                        atomicAdd(stat_s + TG_STAT_SLOTS + slot + 1, static_cast<unsigned long long>(__float2ll_rn((b2) * WFD_FX_GRAD)));
                    } else {
                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gnb_red) + slot, static_cast<unsigned long long>(__float2ll_rn((b1) * WFD_FX_GRAD)));
                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gnb_red) + slot + 1, static_cast<unsigned long long>(__float2ll_rn((b2) * WFD_FX_GRAD)));
                    }
                }
            }
            if (p.rowdot)
                atomicAdd(reinterpret_cast<unsigned long long*>(p.rowdot) + pixel, static_cast<unsigned long long>(__float2ll_rn((dot) * WFD_FX_GRAD)));
            if (dbg_on) e_busy += clock64() - e0;
            if (++acc == acc_stages) {
                acc = 0;
                acc_phase ^= 1;
            }
          }
        }
        if (dbg_on && blockIdx.x == 0 && threadIdx.x == W_EPI0 * 32) {
            p.dbg[5] = clock64() - t_begin;
            p.dbg[6] = w_tfull;
            p.dbg[7] = e_busy;
        }
    };
    if (XF) {
        // registers per thread: 96 at launch (640 threads = 61440, the pool setmaxnreg moves registers within); the producer
        // / MMA warpgroup and the transform warps give part of their share up, the epilogue warps grow to 136
        // (4 x 40 + 8 x 136 + 8 x 80 warps x 32 lanes = 60416 <= 61440 - a request beyond the pool would wait forever)
        if (warp < 4) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
            if (warp == 0 || warp == W_TMAB) producer_role();
            else if (warp == 1) mma_role();
        } else if (warp >= W_XF0) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
            xform_role();
        } else {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
            epilogue_role();
        }
    } else {
        if (warp == 0 || warp == W_TMAB) producer_role();
        else if (warp == 1) mma_role();
        else epilogue_role();
    }

    tc_fence_before();
    __syncthreads();
    if (EPI == 3 && p.gnb_red && p.stat_private) {  // the fast GroupNorm-backward epilogue's per-warp copies
        for (int i = threadIdx.x; i < 128; i += blockDim.x) {
            unsigned long long v = 0ull;
            for (int wv = 0; wv < TG_EPI_THREADS / 32; ++wv) v += stat_s[TG_STAT_SLOTS + wv * 128 + i];
            if (v != 0ull) atomicAdd(reinterpret_cast<unsigned long long*>(p.gnb_red) + i, v);
        }
    } else if (p.gn_sums && p.stat_private) {  // sum the eight per-warp copies (integers: the order does not matter)
        for (int i = threadIdx.x; i < 128; i += blockDim.x) {
            unsigned long long v = 0ull;
            for (int wv = 0; wv < TG_EPI_THREADS / 32; ++wv) v += stat_s[wv * 128 + i];
            if (v != 0ull) atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) + i, v);
        }
    } else if (p.gn_sums || p.gnb_red) {
        for (int i = threadIdx.x; i < TG_STAT_SLOTS; i += blockDim.x) {
            if (p.gn_sums) {
                const unsigned long long v = stat_s[i];
                if (v != 0ull) atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) + i, v);
            }
            if (p.gnb_red) {
                const unsigned long long v = stat_s[TG_STAT_SLOTS + i];
                if (v != 0ull) atomicAdd(reinterpret_cast<unsigned long long*>(p.gnb_red) + i, v);
            }
        }
    }
    if (CG == 2) cluster_sync_all();  // neither CTA may free TMEM / exit while the pair still works on its memory
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ checking kernel
// One thread per output element; reads the same A / Bm buffers through raw pointers. Test infrastructure only.
template <bool BF16>
__global__ void tapgemm_ref_kernel(const TapGemmParams p) {
    const long long rows_per_sample = static_cast<long long>(p.H) * p.W;
    const long long total_rows = static_cast<long long>(p.m_tiles) * TG_BM;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long per_z = total_rows * p.N;
    if (idx >= per_z * p.z_count) return;
    const int z = static_cast<int>(idx / per_z);
    const long long rem = idx - z * per_z;
    const long long pixel = rem / p.N;
    const int n = static_cast<int>(rem - pixel * p.N);
    const int bb = static_cast<int>(pixel / rows_per_sample);
    const int pr = static_cast<int>(pixel - bb * rows_per_sample);
    const int y = pr / p.W, x = pr - y * p.W;
    const int n_tile = n / p.BN;
    const int c0 = p.a_c0 ? p.a_c0[n_tile] : 0;
    const int ab = bb + z * p.a_zmul;
    const int bz = z * p.b_zmul + bb * p.b_bbmul;
    const uint16_t* A = reinterpret_cast<const uint16_t*>(p.a_ptr);
    const uint16_t* Bm = reinterpret_cast<const uint16_t*>(p.b_ptr) + bz * p.b_zstride + static_cast<long long>(n) * p.ldb;
    float acc = 0.f;
    for (int tap = 0; tap < p.n_taps; ++tap) {
        const int xi = x * p.in_stride + p.tap_dx[tap];
        const int yi = y * p.in_stride + p.tap_dy[tap] + p.a_y_off;
        if (xi < 0 || xi >= p.a_W || yi < 0 || yi >= p.a_H || ab >= p.a_B) continue;
        const uint16_t* arow = A + ((static_cast<long long>(ab) * p.a_H + yi) * p.a_W + xi) * p.a_ld;
        if (p.a_mn_major) arow = A + static_cast<long long>(ab) * p.a_C * p.a_ld + xi;  // column xi of the [K][M] sample
        const long long a_kstride = p.a_mn_major ? p.a_ld : 1;
        const bool ax_pad = p.ax_stats && (yi < p.ax_y0 || yi >= p.ax_y1);  // a zeroed halo row at the map border
        for (int cc = 0; cc < p.cpt; ++cc) {
            const uint16_t* brow = Bm + static_cast<long long>(b_chunk_index(p, tap, cc)) * TG_KC;
            for (int k = 0; k < TG_KC; ++k) {
                const int c = c0 + cc * TG_KC + k;
                if (c >= p.a_C) break;
                float av = lo16_to_f<BF16>(arow[c * a_kstride]);
                if (p.ax_stats) {
                    // operand transform: GroupNorm(+SiLU) of the raw activation, rounded to the operand type as the
                    // transform warps of the tensor-core kernel store it
                    const double n_stat = static_cast<double>(p.ax_cpg) * static_cast<double>(p.ax_hw);
                    const acc_t* st = p.ax_stats + (static_cast<long long>(ab) * (p.a_C / p.ax_cpg) + c / p.ax_cpg) * 2;
                    const double m = static_cast<double>(st[0]) * (1.0 / WFD_FX_STATS) / n_stat;
                    double var = static_cast<double>(st[1]) * (1.0 / WFD_FX_STATS) / n_stat - m * m;
                    if (var < 0) var = 0;
                    const float mean = static_cast<float>(m);
                    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.ax_eps)));
                    const float ga = p.ax_gamma[c];
                    const float sc = rstd * ga, sh = p.ax_beta[c] - mean * rstd * ga;
                    float v = av * sc + sh;
                    if (p.ax_silu) v = v * fast_sigmoid(v);
                    av = ax_pad ? 0.f : lo16_to_f<BF16>(pack2<BF16>(v, 0.f));
                }
                acc = fmaf(av, lo16_to_f<BF16>(brow[k]), acc);
            }
        }
    }
    float v = acc * p.alpha;
    if (p.bias) v += p.bias[n];
    const long long off = z * p.c_zstride + bb * p.out_sb + static_cast<long long>(y) * p.out_sy +
                          static_cast<long long>(x) * p.out_sx + n;
    if (p.rs_in) {
        const float2 st = p.rs_in[pixel + z * total_rows];
        v = __expf(v - st.x) * st.y;
    }
    if (p.residual && p.sb_d)
        v = lo16_to_f<BF16>(reinterpret_cast<const uint16_t*>(p.residual)[off]) * (v - p.sb_d[pixel + z * total_rows]) * p.sb_scale;
    else if (p.residual) v += lo16_to_f<BF16>(reinterpret_cast<const uint16_t*>(p.residual)[off]);
    if (p.rowdot)
        atomicAdd(reinterpret_cast<unsigned long long*>(p.rowdot) + pixel,
                  static_cast<unsigned long long>(__float2ll_rn((v * lo16_to_f<BF16>(reinterpret_cast<const uint16_t*>(p.dotsrc)[pixel * p.ld_dot + n])) * WFD_FX_GRAD)));
    if (p.out_fp32) {
        reinterpret_cast<float*>(p.out)[off] = v;
    } else {
        const uint32_t w = pack2<BF16>(v, 0.f);
        reinterpret_cast<uint16_t*>(p.out)[off] = static_cast<uint16_t>(w & 0xFFFF);
        if (p.gn_sums) {
            const float r = lo16_to_f<BF16>(w);
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.gn_sums) + (static_cast<long long>(bb) * p.gn_G + n / p.gn_cpg) * 2;
            atomicAdd(dst, static_cast<unsigned long long>(__float2ll_rn((r) * WFD_FX_STATS)));
            atomicAdd(dst + 1, static_cast<unsigned long long>(__float2ll_rn((r * r) * WFD_FX_STATS)));
        }
        if (p.gnb_red) {
            const int g = n / p.gnb_cpg;
            const acc_t* st = p.gnb_stats + (static_cast<long long>(bb) * p.gn_G + g) * 2;
            const double dm = static_cast<double>(st[0]) * (1.0 / WFD_FX_STATS) * p.gnb_inv_n;
            const float m = static_cast<float>(dm);
            const float var = fmaxf(static_cast<float>(static_cast<double>(st[1]) * (1.0 / WFD_FX_STATS) * p.gnb_inv_n - dm * dm), 0.f);
            const float rstd = rsqrtf(var + p.gnb_eps);
            const float xh = (lo16_to_f<BF16>(reinterpret_cast<const uint16_t*>(p.gnb_x)[off]) - m) * rstd;
            const float ga = p.gnb_gamma[n];
            float du = lo16_to_f<BF16>(w);
            if (p.gnb_silu) {
                const float u = fmaf(xh, ga, p.gnb_beta[n]);
                const float sg = fast_sigmoid(u);
                du *= sg * fmaf(u, 1.f - sg, 1.f);
            }
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.gnb_red) + (static_cast<long long>(bb) * p.gn_G + g) * 2;
            atomicAdd(dst, static_cast<unsigned long long>(__float2ll_rn((du * ga) * WFD_FX_GRAD)));
            atomicAdd(dst + 1, static_cast<unsigned long long>(__float2ll_rn((du * ga * xh) * WFD_FX_GRAD)));
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// Row reuse applies to full 3x3 stride-1 tap sets; the caller asks for it with p.reuse = 1 and must have packed Bm in
// the (dx, cc, dy) chunk order (b_chunk_index). It is dropped here when a stage would not leave room for 3 stages.
bool tapgemm_reuse_ok(const TapGemmParams& p) {
    if (p.n_taps != 9 || p.in_stride != 1 || p.w_box < 8 || p.z_count != 1 || p.b_zmul || p.b_bbmul) return false;
    bool seen[9] = {false};
    for (int t = 0; t < 9; ++t) {
        const int dx = p.tap_dx[t], dy = p.tap_dy[t];
        if (dx < -1 || dx > 1 || dy < -1 || dy > 1) return false;
        seen[(dy + 1) * 3 + dx + 1] = true;
    }
    for (bool b : seen)
        if (!b) return false;
    const size_t per_stage = static_cast<size_t>(p.h_box + 2) * p.w_box * 128 + 3 * static_cast<size_t>(p.BN) * 128;
    return per_stage * 3 <= 216 * 1024;
}

int tapgemm_prepare(TapGemmOp* op) {
    TapGemmParams& p = op->p;
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return -2;
    }
    if (p.BN % 16 != 0 || p.BN < 16 || p.BN > 256 || p.N % 16 != 0 || p.w_box * p.h_box != TG_BM ||
        p.n_taps < 1 || p.n_taps > TG_MAX_TAPS || p.cpt < 1) {
        set_error("tapgemm_prepare: bad tiling (BN=%d N=%d box=%dx%d taps=%d cpt=%d)", p.BN, p.N, p.w_box, p.h_box,
                  p.n_taps, p.cpt);
        return -3;
    }
    if (p.rs_out && (p.residual || p.rowdot || p.gnb_red || p.gn_sums || p.bias || p.rs_in)) {
        set_error("tapgemm_prepare: the row-statistics epilogue stores nothing and takes no other epilogue option");
        return -3;
    }
    if ((p.residual ? 1 : 0) + (p.rowdot ? 1 : 0) + (p.gnb_red ? 1 : 0) > 1) {
        set_error("tapgemm_prepare: residual, row-dot and GroupNorm-backward epilogues are mutually exclusive");
        return -3;
    }
    {
        // 256-bit epilogue accesses need every row piece 32-byte aligned
        const long long es = p.out_fp32 ? 4 : 2;
        auto al = [](long long v) { return (v & 31) == 0; };
        bool ok = al(reinterpret_cast<long long>(p.out)) && al(p.out_sx * es) && al(p.out_sy * es) && al(p.out_sb * es) &&
                  al(p.c_zstride * es);
        if (p.residual) ok = ok && !p.out_fp32 && al(reinterpret_cast<long long>(p.residual));
        if (p.gnb_red) ok = ok && !p.out_fp32 && al(reinterpret_cast<long long>(p.gnb_x));
        if (p.rowdot) ok = ok && al(reinterpret_cast<long long>(p.dotsrc)) && al(p.ld_dot * 2);
        static int no_wide = -1;
        if (no_wide < 0) {
            const char* e = getenv("WFD_NO_WIDE_IO");
            no_wide = (e && e[0] == '1') ? 1 : 0;
        }
        p.wide_io = (ok && !no_wide) ? 1 : 0;
    }
    {
        static int split_max = -1;  // widest N tile drained in tile-split mode (WFD_EPI_SPLIT_MAX, 0 = never; <= 128)
        if (split_max < 0) {
            const char* e = getenv("WFD_EPI_SPLIT_MAX");
            split_max = e ? atoi(e) : 96;  // measured on B200: 13.76 / 13.49 / 13.31 / 13.39 ms per step for 0 / 64 / 96 / 128
            if (split_max > 128) split_max = 128;
        }
        p.epi_split = (p.BN <= split_max) ? 1 : 0;
    }
    if (p.kc_list && (p.reuse || p.kk_last || !p.kc_cnt || p.b_resident)) {
        set_error("tapgemm_prepare: a K-chunk skip list needs the plain schedule with whole 64-channel chunks");
        return -3;
    }
    if (p.ax_stats && (p.reuse != 2 || !p.ax_gamma || !p.ax_beta || p.ax_cpg < 8 || p.ax_cpg % 8 || p.a_C % p.ax_cpg || p.ax_hw < 1)) {
        set_error("tapgemm_prepare: the operand transform needs the patch schedule and whole 8-channel vectors per GroupNorm group");
        return -3;
    }
    if (p.reuse == 2) {
        bool ok = p.n_taps == 9 && p.in_stride == 1 && p.z_count == 1 && !p.b_zmul && !p.b_bbmul && p.w_box == 8 &&
                  p.h_box == 16 && !p.kc_list && !p.a_mn_major && !p.a_zmul;
        bool seen[9] = {false};
        for (int t = 0; ok && t < 9; ++t) {
            const int dx = p.tap_dx[t], dy = p.tap_dy[t];
            if (dx < -1 || dx > 1 || dy < -1 || dy > 1) ok = false;
            else seen[(dy + 1) * 3 + dx + 1] = true;
        }
        for (bool b : seen) ok = ok && b;
        if (!ok) {
            set_error("tapgemm_prepare: the patch schedule needs a full 3x3 stride-1 tap set and 16 x 8 pixel patches");
            return -3;
        }
    } else if (p.reuse && !tapgemm_reuse_ok(p)) {
        set_error("tapgemm_prepare: row-reuse schedule requested for an unsupported tap set / tile");
        return -3;
    }
    // CTA pairs (cta_group::2, one 256 x BN tile per pair, each CTA holding half of the B rows) halve the shared-memory
    // traffic per MMA; used for the wide-N plain-schedule contractions whenever M tiles come in pairs
    {
        static int no2 = -1;
        if (no2 < 0) {
            const char* e = getenv("WFD_NO_2CTA");
            no2 = (e && e[0] == '1') ? 1 : 0;
        }
        const bool pairs_ok = ((p.tiles_x * p.tiles_y) % 2 == 0) && !p.a_zmul;
        // row-reuse (grouped 3x3) contractions have narrow N tiles: their MMA-issuing thread is the bottleneck, and a pair
        // halves the instructions per MAC (one 256 x BN MMA instead of two 128 x BN)
        static int reuse2 = -1;
        if (reuse2 < 0) {
            const char* e = getenv("WFD_REUSE_2CTA");
            reuse2 = (e && e[0] == '0') ? 0 : 1;
        }
        const bool wide = !p.reuse && p.BN >= 128 && p.BN % 32 == 0;
        const bool narrow_reuse = p.reuse && reuse2 && p.BN % 16 == 0;
        p.cg = (!no2 && p.cg != 1 && (wide || narrow_reuse) && pairs_ok) ? 2 : 1;
    }
    {
        // accumulator stages in TMEM (512 columns): two 256-column ones for wide tiles, four for N tiles of <= 128 columns,
        // eight for <= 64 (WFD_ACC_STAGES caps it, for A/B runs)
        static int cap = -1;
        if (cap < 0) {
            const char* e = getenv("WFD_ACC_STAGES");
            cap = e ? atoi(e) : 8;
            if (cap != 2 && cap != 4 && cap != 8) cap = 8;
        }
        int s = p.BN <= 64 ? 8 : (p.BN <= 128 ? 4 : 2);
        if (s > cap) s = cap;
        p.acc_stages = s;
    }
    const bool patch = p.reuse == 2;
    if (patch) {
        // weights of one N tile stay resident when they leave room for >= 3 activation boxes (grouped layers); the CTAs
        // then walk contiguous tile ranges so that the N tile changes a few times per launch, not every other tile
        // Measured on B200 (round 2, profiles/r02_summary.md): resident weights + contiguous tile ranges are SLOWER than
        // streaming the weight boxes through their ring (batch-8 step 12.44 vs 12.12 ms; groups=64 convs 2.71 vs 2.35 ms),
        // so the mode is opt-in (WFD_PATCH_RESIDENT=1).
        static int want_res = -1;
        if (want_res < 0) {
            const char* e = getenv("WFD_PATCH_RESIDENT");
            want_res = (e && e[0] == '1') ? 1 : 0;
        }
        const size_t res_bytes = 9 * static_cast<size_t>(p.cpt) * (p.BN / p.cg) * 128;
        p.b_resident = (want_res && res_bytes <= 120 * 1024) ? 1 : 0;
        static int chunked = -1, prefetch = -1;
        if (chunked < 0) {
            const char* e = getenv("WFD_PATCH_CHUNKED");
            chunked = (e && e[0] == '0') ? 0 : 1;
            const char* f = getenv("WFD_PATCH_PREFETCH");
            prefetch = f ? atoi(f) : 0;
        }
        p.tile_chunked = (p.b_resident && p.n_group <= 1 && chunked) ? 1 : 0;
        p.a_prefetch = prefetch;
        // weight stage = all nine taps of a 64-channel chunk when two such stages fit beside three activation boxes
        // (narrow N tiles: one synchronisation point per chunk instead of three), else one vertical offset (three taps)
        static int taps9 = -1;
        if (taps9 < 0) {
            const char* e = getenv("WFD_PATCH_TAPS9");
            taps9 = (e && e[0] == '0') ? 0 : 1;
        }
        const size_t box_b = static_cast<size_t>(p.BN / p.cg) * 128;
        p.b_taps = (taps9 && 2 * 9 * box_b + 3 * 24 * 1024 <= 196 * 1024) ? 9 : 3;
    }
    const CUtensorMapDataType dt = p.is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (p.a_mn_major) {
        if (p.n_taps != 1 || p.reuse || p.a_H != 1 || p.w_box != TG_BM || p.a_c0 || p.kk_last || p.in_stride != 1 || p.a_zmul) {
            set_error("tapgemm_prepare: an M-major A operand needs a plain one-tap GEMM over whole 64-row K chunks");
            return -3;
        }
        // A as stored: [a_B][a_C = K rows][a_W = M columns], M contiguous; box = 64 m x 64 k
        cuuint64_t dims[4] = {(cuuint64_t)p.a_W, (cuuint64_t)p.a_C, 1, (cuuint64_t)p.a_B};
        cuuint64_t strides[3] = {(cuuint64_t)p.a_ld * 2, (cuuint64_t)p.a_ld * 2 * p.a_C, (cuuint64_t)p.a_ld * 2 * p.a_C};
        cuuint32_t box[4] = {64, TG_KC, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&op->tmA, dt, 4, const_cast<void*>(p.a_ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(A, M-major) failed: %d (M=%d K=%d B=%d ld=%lld)", (int)r, p.a_W, p.a_C, p.a_B, p.a_ld);
            return -4;
        }
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)p.a_C, (cuuint64_t)p.a_W, (cuuint64_t)p.a_H, (cuuint64_t)p.a_B};
        cuuint64_t strides[3] = {(cuuint64_t)p.a_ld * 2, (cuuint64_t)p.a_ld * 2 * p.a_W,
                                 (cuuint64_t)p.a_ld * 2 * p.a_W * p.a_H};
        // strided traversal (stride-2 downsample conv): the box spans in_stride * (n - 1) + 1 input elements
        cuuint32_t box[4] = {TG_KC, (cuuint32_t)((p.w_box - 1) * p.in_stride + 1),
                             (cuuint32_t)((p.h_box - 1) * p.in_stride + 1), 1};
        if (p.reuse) box[2] = (cuuint32_t)(p.h_box + 2);
        if (patch) box[1] = (cuuint32_t)(p.w_box + 2);
        cuuint32_t estr[4] = {1, (cuuint32_t)p.in_stride, (cuuint32_t)p.in_stride, 1};
        CUresult r = enc(&op->tmA, dt, 4, const_cast<void*>(p.a_ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d B=%d ld=%lld box=%ux%u)", (int)r, p.a_C,
                      p.a_W, p.a_H, p.a_B, p.a_ld, box[1], box[2]);
            return -4;
        }
    }
    if (p.reuse) {
        // weights as (64 elements, rows, chunk): one box fetches the three vertical-tap chunks of a (dx, cc) pair
        cuuint64_t dims[3] = {(cuuint64_t)TG_KC, (cuuint64_t)p.b_rows, (cuuint64_t)(9 * p.cpt)};
        cuuint64_t strides[2] = {(cuuint64_t)p.ldb * 2, (cuuint64_t)TG_KC * 2};
        cuuint32_t box[3] = {TG_KC, (cuuint32_t)(p.BN / p.cg), (cuuint32_t)((patch && p.b_resident) ? 9 : (patch ? p.b_taps : 3))};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&op->tmB, dt, 3, const_cast<void*>(p.b_ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(B, reuse) failed: %d (rows=%d ldb=%lld BN=%d cpt=%d)", (int)r, p.b_rows,
                      p.ldb, p.BN, p.cpt);
            return -5;
        }
    } else {
        const long long ktot = static_cast<long long>(p.n_taps) * p.cpt * TG_KC;
        const int zdim = p.b_Z > 0 ? p.b_Z : 1;
        cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)p.b_rows, (cuuint64_t)zdim};
        cuuint64_t strides[2] = {(cuuint64_t)p.ldb * 2, (cuuint64_t)(p.b_zstride > 0 ? p.b_zstride : p.ldb * p.b_rows) * 2};
        cuuint32_t box[3] = {TG_KC, (cuuint32_t)(p.BN / p.cg), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&op->tmB, dt, 3, const_cast<void*>(p.b_ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(B) failed: %d (ktot=%lld rows=%d ldb=%lld zstride=%lld BN=%d)", (int)r,
                      ktot, p.b_rows, p.ldb, p.b_zstride, p.BN);
            return -5;
        }
    }
    const size_t a_stage = patch ? ((static_cast<size_t>(p.h_box + 2) * (p.w_box + 2) * 128 + 1023) & ~size_t(1023))
                                 : (p.reuse ? static_cast<size_t>(p.h_box + 2) * p.w_box * 128 : TG_BM * 128);
    // weights of one N tile resident in shared memory when they are small (grouped layers with few channels per group)
    const size_t b_res_bytes = 9 * static_cast<size_t>(p.cpt) * (p.BN / p.cg) * 128;
    if (!patch) {
        // row-reuse schedule, measured on B200 (round 1): 0.27 ms vs 0.21 ms per 3072-channel groups=64 conv with / without
        // resident weights, so the mode is opt-in there (WFD_RESIDENT_B=1)
        static int res = -1;
        if (res < 0) {
            const char* e = getenv("WFD_RESIDENT_B");
            res = (e && e[0] == '1') ? 1 : 0;
        }
        p.b_resident = (res && p.reuse && p.cg == 1 && b_res_bytes <= 64 * 1024) ? 1 : 0;
    }
    const size_t b_stage = p.b_resident ? 0 : (patch ? p.b_taps : (p.reuse ? 3 : 1)) * static_cast<size_t>(p.BN / p.cg) * 128;
    const size_t budget = 196 * 1024;
    int stages, a_stages = 0;
    if (patch) {
        if (p.b_resident) {
            a_stages = static_cast<int>((budget - b_res_bytes) / a_stage);
            if (a_stages > 6) a_stages = 6;
            stages = 1;  // the weight ring is not used
        } else {
            a_stages = 3;
            stages = static_cast<int>((budget - a_stages * a_stage) / b_stage);
            if (stages > 8) stages = 8;
            if (stages >= 6 && budget - a_stages * a_stage - stages * b_stage >= a_stage) a_stages = 4;
            static int a_over = -1, b_over = -1;  // WFD_PATCH_ASTAGES / WFD_PATCH_BSTAGES: experiment knobs
            if (a_over < 0) {
                const char* e = getenv("WFD_PATCH_ASTAGES");
                a_over = e ? atoi(e) : 0;
                const char* f = getenv("WFD_PATCH_BSTAGES");
                b_over = f ? atoi(f) : 0;
            }
            if (a_over > 0 || b_over > 0) {
                int as = a_over > 0 ? a_over : a_stages, bs = b_over > 0 ? b_over : stages;
                while (bs > 2 && as * a_stage + bs * b_stage > budget) --bs;
                while (as > 2 && as * a_stage + bs * b_stage > budget) --as;
                a_stages = as;
                stages = bs;
            }
        }
        if (a_stages < 2 || stages < 1 || (!p.b_resident && stages < 2)) {
            set_error("tapgemm_prepare: the patch schedule does not fit shared memory (BN=%d cpt=%d)", p.BN, p.cpt);
            return -3;
        }
    } else {
        const size_t per_stage = a_stage + b_stage;
        stages = static_cast<int>((budget - (p.b_resident ? b_res_bytes : 0)) / per_stage);
        if (stages > 8) stages = 8;
        {
            static int cap = -1;  // WFD_MAX_STAGES: experiment knob
            if (cap < 0) {
                const char* e = getenv("WFD_MAX_STAGES");
                cap = e ? atoi(e) : 0;
            }
            if (cap > 0 && stages > cap) stages = cap;
        }
        if (stages < 2) stages = 2;
    }
    p.a_stages = a_stages;
    op->stages = stages;
    {
        // specialised epilogue for the plain conv cases (see tapgemm_kernel); WFD_NO_EPI_FAST=1 keeps the generic loop
        const char* nf = getenv("WFD_NO_EPI_FAST");   // 1: generic loop everywhere, 2: generic loop for the attention modes
        const bool no_fast = nf && nf[0] == '1', no_attn = nf && (nf[0] == '1' || nf[0] == '2');
        const int chunks = p.BN / 16;
        const int per_warp = p.epi_split ? chunks : (chunks + 1) / 2;
        p.epi_fast = (!p.out_fp32 && p.wide_io && !p.rowdot && !p.gnb_red && !p.rs_out && !p.rs_in && !p.sb_d && p.alpha == 1.f &&
                      per_warp <= 6 && !no_fast) ? 1 : 0;
        const int mode = (p.bias ? 4 : 0) | (p.residual ? 2 : 0) | (p.gn_sums ? 1 : 0);
        if (mode != 0 && mode != 4 && mode != 5 && mode != 7) p.epi_fast = 0;
        // ... for the backward convs with the fused GroupNorm-backward reduction (per-CTA mean / rstd table of <= 256 pairs)
        {
            const int per_sample = p.tiles_x * p.tiles_y;
            const int samples = per_sample > 0 ? (p.m_tiles + per_sample - 1) / per_sample : 1;
            if (!p.epi_fast && !p.out_fp32 && p.wide_io && p.gnb_red && !p.residual && !p.rowdot && !p.bias && !p.gn_sums &&
                !p.rs_out && !p.rs_in && !p.sb_d && !p.ax_stats && p.alpha == 1.f && per_warp <= 6 && p.z_count == 1 &&
                samples * p.gn_G <= 256 && p.gnb_cpg % 16 == 0 && !no_fast && !(nf && nf[0] == '3'))
                p.epi_fast = 3;
        }
        // ... and for the attention products whose epilogue is the softmax work (up to eight chunks per warp)
        if (!p.epi_fast && !p.out_fp32 && p.wide_io && !p.rowdot && !p.gnb_red && !p.bias && !p.gn_sums && per_warp <= 8 &&
            ((p.rs_out ? 1 : 0) + (p.rs_in ? 1 : 0) + (p.sb_d ? 1 : 0)) == 1 && (p.sb_d ? p.residual != nullptr : p.residual == nullptr) &&
            !no_attn)
            p.epi_fast = 2;
        // per-warp private statistic slots when (samples x GroupNorm groups x 2) of the launch fit 128
        const int per_sample = p.tiles_x * p.tiles_y;
        const int samples = per_sample > 0 ? (p.m_tiles + per_sample - 1) / per_sample : 1;
        p.stat_private = (p.epi_fast && ((p.gn_sums && !p.gnb_red) || p.epi_fast == 3) && p.z_count == 1 && samples * p.gn_G * 2 <= 128 &&
                          getenv("WFD_NO_STAT_PRIVATE") == nullptr) ? 1 : 0;
    }
    op->threads = p.ax_stats ? TG_THREADS_XF : TG_THREADS;
    op->smem_bytes = 1024 + (patch ? a_stages * a_stage + stages * b_stage : stages * (a_stage + b_stage)) +
                     (p.b_resident ? b_res_bytes : 0) + (2 * stages + 20 + 4 * a_stages) * sizeof(uint64_t) +
                     6 * 256 * sizeof(float) + 2 * TG_STAT_SLOTS * sizeof(unsigned long long) + (p.ax_stats ? 2 * 512 * sizeof(float) : 0) +
                     (p.epi_fast == 3 ? 2 * 256 * sizeof(float) : 0) + 16;
    return 0;
}

template <bool BF16, int CG, bool XF, int EPI>
static cudaError_t launch_variant(const TapGemmOp* op, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_kernel<BF16, CG, XF, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(XF ? TG_THREADS_XF : TG_THREADS);
    cfg.dynamicSmemBytes = op->smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tapgemm_kernel<BF16, CG, XF, EPI>, op->tmA, op->tmB, op->p, op->stages);
}

int tapgemm_launch(const TapGemmOp* op, cudaStream_t stream) {
    const TapGemmParams& p = op->p;
    const int cg = p.cg == 2 ? 2 : 1;
    if (cg == 2 && (p.m_tiles % 2)) {
        set_error("tapgemm launch: CTA-pair schedule needs an even number of M tiles (got %d)", p.m_tiles);
        return -3;
    }
    const int total = (p.m_tiles / cg) * p.n_tiles * p.z_count;  // (pair) tiles
    if (total <= 0) return 0;
    int sms = num_sms();
    {
        static int limit = -1;  // WFD_GRID_LIMIT: experiment knob (fewer resident CTAs), not used in production
        if (limit < 0) {
            const char* e = getenv("WFD_GRID_LIMIT");
            limit = e ? atoi(e) : 0;
        }
        if (limit > 0 && sms > limit) sms = limit;
    }
    int grid = total * cg < sms ? total * cg : (sms / cg) * cg;
    char pname[160];
    if (g_prof_on)
        snprintf(pname, sizeof(pname), "tapgemm[%.11s]:N=%d,BN=%d,taps=%d,cpt=%d,aC=%d,f32=%d,reuse=%d,side=%d,gns=%d,cg=%d,xf=%d", op->tag, p.N, p.BN,
                 p.n_taps, p.cpt, p.a_C, p.out_fp32, p.reuse, p.residual ? 1 : (p.rowdot ? 2 : (p.gnb_red ? 3 : 0)),
                 p.gn_sums ? 1 : 0, cg, p.ax_stats ? 1 : 0);
    ProfScope prof_scope(g_prof_on ? pname : "", stream);
    cudaError_t e;
    // WFD_TG_DBG=1 (diagnosis only): every launch runs with the per-role wait counters of CTA 0 and prints them
    static int tg_dbg = -1;
    static long long* tg_dbg_buf = nullptr;
    if (tg_dbg < 0) {
        const char* ev = getenv("WFD_TG_DBG");
        tg_dbg = (ev && ev[0] == '1') ? 1 : 0;
#ifndef WFD_TG_DBG_BUILD
        if (tg_dbg || getenv("WFD_TG_SKIP"))
            fprintf(stderr, "wfd: WFD_TG_DBG / WFD_TG_SKIP need a library built with -DWFD_TG_DBG_BUILD=1 (counters read zero)\n");
#endif
        if (tg_dbg && cudaMalloc(&tg_dbg_buf, 16 * sizeof(long long)) != cudaSuccess) tg_dbg = 0;
    }
    static int tg_skip = -1;  // WFD_TG_SKIP=<bits>: see TapGemmParams::dbg_skip (row-reuse launches only)
    if (tg_skip < 0) {
        const char* ev = getenv("WFD_TG_SKIP");
        tg_skip = ev ? atoi(ev) : 0;
    }
    TapGemmOp dbg_op;
    if (tg_skip && p.reuse) {
        dbg_op = *op;
        dbg_op.p.dbg_skip = tg_skip;
        op = &dbg_op;
    }
    if (tg_dbg && !p.dbg) {
        dbg_op = *op;
        dbg_op.p.dbg = tg_dbg_buf;
        cudaMemsetAsync(tg_dbg_buf, 0, 16 * sizeof(long long), stream);
        op = &dbg_op;
    }
    e = cudaErrorInvalidValue;
#define WFD_TG_LAUNCH(XFV, EPIV)                                                                                            \
    do {                                                                                                                   \
        if (p.is_bf16) e = (cg == 2) ? launch_variant<true, 2, XFV, EPIV>(op, grid, stream) : launch_variant<true, 1, XFV, EPIV>(op, grid, stream); \
        else e = (cg == 2) ? launch_variant<false, 2, XFV, EPIV>(op, grid, stream) : launch_variant<false, 1, XFV, EPIV>(op, grid, stream);       \
    } while (0)
    if (p.ax_stats) {  // operand transform: plain conv launches only (tapgemm_prepare)
        if (p.epi_fast == 1) WFD_TG_LAUNCH(true, 1);
        else if (p.epi_fast == 0) WFD_TG_LAUNCH(true, 0);
        else set_error("tapgemm launch: operand transform with epilogue form %d", p.epi_fast);
    } else {
        switch (p.epi_fast) {
            case 0: WFD_TG_LAUNCH(false, 0); break;
            case 1: WFD_TG_LAUNCH(false, 1); break;
            case 2: WFD_TG_LAUNCH(false, 2); break;
            case 3: WFD_TG_LAUNCH(false, 3); break;
            default: set_error("tapgemm launch: unknown epilogue form %d", p.epi_fast); break;
        }
    }
#undef WFD_TG_LAUNCH
    if (tg_dbg && op == &dbg_op && e == cudaSuccess) {
        long long h[16];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, tg_dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
        const TapGemmParams& q = op->p;
        fprintf(stderr,
                "tgdbg N=%d BN=%d taps=%d cpt=%d aC=%d reuse=%d cg=%d side=%d tiles=%d grid=%d stages=%d | A %lld we %lld is %lld | "
                "B %lld we %lld is %lld | MMA %lld wte %lld wf %lld | EPI %lld wtf %lld busy %lld | XF %lld wait %lld math %lld\n",
                q.N, q.BN, q.n_taps, q.cpt, q.a_C, q.reuse, cg, q.residual ? 1 : (q.rowdot ? 2 : (q.gnb_red ? 3 : 0)), total, grid,
                op->stages, h[0], h[1], h[10], h[8], h[9], h[11], h[2], h[3], h[4], h[5], h[6], h[7], h[12], h[13], h[14]);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("tapgemm launch: %s", cudaGetErrorString(e));
        return -7;
    }
    ++g_launch_count;
    if (debug_sync_enabled()) {
        char what[256];
        snprintf(what, sizeof(what),
                 "tapgemm(m_tiles=%d n_tiles=%d z=%d BN=%d N=%d taps=%d cpt=%d box=%dx%d stages=%d smem=%zu aC=%d aW=%d aH=%d aB=%d reuse=%d cg=%d)",
                 p.m_tiles, p.n_tiles, p.z_count, p.BN, p.N, p.n_taps, p.cpt, p.w_box, p.h_box, op->stages,
                 op->smem_bytes, p.a_C, p.a_W, p.a_H, p.a_B, p.reuse, cg);
        return debug_sync_check(what, stream);
    }
    return 0;
}

// checking-kernel counterpart of the row-statistics epilogue (plain one-tap GEMMs only): one thread per (row, N tile)
template <bool BF16>
__global__ void tapgemm_ref_rowstats_kernel(const TapGemmParams p) {
    const long long total_rows = static_cast<long long>(p.m_tiles) * TG_BM;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total_rows * p.n_tiles * p.z_count) return;
    const int n_tile = static_cast<int>(idx % p.n_tiles);
    const long long zr = idx / p.n_tiles;
    const int z = static_cast<int>(zr / total_rows);
    const long long pixel = zr - z * total_rows;
    const long long rows_per_sample = static_cast<long long>(p.H) * p.W;
    const int bb = static_cast<int>(pixel / rows_per_sample);
    const int x = static_cast<int>(pixel - bb * rows_per_sample);
    const int ab = bb + z * p.a_zmul, bz = z * p.b_zmul + bb * p.b_bbmul;
    const uint16_t* arow = reinterpret_cast<const uint16_t*>(p.a_ptr) + (static_cast<long long>(ab) * p.a_W + x) * p.a_ld;
    float m = -INFINITY, l = 0.f;
    for (int n = n_tile * p.BN; n < min(p.N, (n_tile + 1) * p.BN); ++n) {
        const uint16_t* brow = reinterpret_cast<const uint16_t*>(p.b_ptr) + bz * p.b_zstride + static_cast<long long>(n) * p.ldb;
        float acc = 0.f;
        for (int k = 0; k < p.cpt * TG_KC && k < p.a_C; ++k) acc = fmaf(lo16_to_f<BF16>(arow[k]), lo16_to_f<BF16>(brow[k]), acc);
        const float v = acc * p.alpha;
        if (v > m) {
            l *= __expf(m - v);
            m = v;
        }
        l += __expf(v - m);
    }
    const long long slot = ((pixel + z * total_rows) * p.n_tiles + n_tile) * 2;
    p.rs_out[slot] = make_float2(m, l);
    p.rs_out[slot + 1] = make_float2(-INFINITY, 0.f);
}

int tapgemm_launch_ref(const TapGemmOp* op, cudaStream_t stream) {
    const TapGemmParams& p = op->p;
    if (p.rs_out) {
        if (p.n_taps != 1 || p.a_mn_major || p.a_c0) {
            set_error("tapgemm_ref: the row-statistics epilogue is checked for plain one-tap GEMMs only");
            return -3;
        }
        const long long n = static_cast<long long>(p.m_tiles) * TG_BM * p.n_tiles * p.z_count;
        if (n <= 0) return 0;
        const unsigned nb = static_cast<unsigned>((n + 127) / 128);
        if (p.is_bf16) tapgemm_ref_rowstats_kernel<true><<<nb, 128, 0, stream>>>(op->p);
        else tapgemm_ref_rowstats_kernel<false><<<nb, 128, 0, stream>>>(op->p);
        if (cudaGetLastError() != cudaSuccess) {
            set_error("tapgemm_ref rowstats launch failed");
            return -7;
        }
        return 0;
    }
    const long long total = static_cast<long long>(p.m_tiles) * TG_BM * p.N * p.z_count;
    if (total <= 0) return 0;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    if (p.is_bf16)
        tapgemm_ref_kernel<true><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(op->p);
    else
        tapgemm_ref_kernel<false><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(op->p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("tapgemm_ref launch: %s", cudaGetErrorString(e));
        return -7;
    }
    return 0;
}

}  // namespace wfd
