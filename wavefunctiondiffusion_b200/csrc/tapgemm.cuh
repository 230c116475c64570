// tapgemm: the one tensor-core kernel of the WFC-guidance path.
//
//   D[m, n] = alpha * sum_{tap} sum_{k} A[pixel(m) + offset(tap), c0(n_tile) + k] * Bm[n, tap*Kt + k]  (+bias[n]) (+res[m,n])
//
// A is a channels-last activation tensor (B, H, W, C) read through a 4-D TMA tensor map: a 128-row M tile is a
// (w_box x h_box) pixel patch of one sample, a tap is a (dx, dy) shift of the TMA start coordinate, and TMA out-of-bounds
// zero fill provides both the conv padding and the map border of the WFC neighbour contraction. Bm is a K-major matrix
// (packed conv weights / adjacency matrices / another activation) read through a 3-D TMA map. Accumulation happens in
// TMEM via tcgen05.mma (kind::f16, fp32 accumulate), double buffered so the epilogue of tile i overlaps the MMAs of
// tile i+1. The same kernel serves: grouped 3x3 / 1x1 convs forward and backward-data (reference resnet.py:409,425,456),
// the attention linears and QK^T / PV products (attention.py:339-374), and the WFC wave-by-adjacency contraction
// (wfdf/losses.py:63-75).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace wfd {

constexpr int TG_BM = 128;      // rows per M tile (UMMA M)
constexpr int TG_KC = 64;       // K elements per pipeline chunk (128 bytes = one SW128 row)
constexpr int TG_MAX_TAPS = 9;
// Cross-CTA reductions (GroupNorm sums, backward reductions, WFC row dots) are accumulated as 64-bit FIXED-POINT integers:
// integer atomics are order independent, so results are bitwise reproducible run to run and independent of the batch slot.
typedef long long acc_t;
constexpr float WFD_FX_STATS = 1048576.f;          // 2^20: forward statistics (sum x, sum x^2)
constexpr float WFD_FX_GRAD = 1099511627776.f;     // 2^40: backward reductions and row dots (small magnitudes)
constexpr int TG_STAT_SLOTS = 1024;  // per-CTA shared-memory slots for fused GroupNorm partial sums ((sample, group) x 2)
constexpr int TG_EPI_THREADS = 256;                 // warps 2-9: epilogue (two warps per TMEM lane quarter)
constexpr int TG_THREADS = 64 + TG_EPI_THREADS + 32; // warp 0 TMA (activations), warp 1 MMA, warp 10 TMA (weights)
// Launches with the operand transform (ax_stats set) use a warpgroup-aligned layout so that setmaxnreg can move registers
// between the roles: warps 0-3 = TMA (activations), MMA, TMA (weights), idle; warps 4-11 = epilogue; warps 12-19 = transform
constexpr int TG_XF_THREADS = 256;                  // transform warps (two per scheduler: a single in-order warp cannot overlap
                                                    // its quarter-rate tanh with its FMAs)
constexpr int TG_THREADS_XF = 128 + TG_EPI_THREADS + TG_XF_THREADS;  // 640

struct TapGemmParams {
    // ---- tiling
    int m_tiles, n_tiles, z_count;
    int n_group;            // >1: tile ids sweep bands of n_group N tiles (L2 reuse of A for large-K GEMMs); 0/1: M fastest
    int BN;                 // N tile, multiple of 16, <= 256
    int tiles_x, tiles_y;   // M tiles per sample along x / y
    int w_box, h_box;       // w_box * h_box == 128
    int W, H;               // output plane (pixel index = (bb*H + y)*W + x)
    int in_stride;          // input coordinate = out coordinate * in_stride + tap offset
    int a_y_off;            // added to every input row coordinate (row-band plans: the A view starts one halo row early)
    int a_mn_major;         // plain one-tap GEMMs only: A is given TRANSPOSED in memory, [a_C = K rows][a_W = M columns] per
                            // sample with row pitch a_ld (M contiguous) - D = A^T-as-stored x Bm^T without a transpose pass
    int wide_io;            // set by tapgemm_prepare: output and side-input rows are 32-byte aligned (256-bit accesses)
    int epi_split;          // set by tapgemm_prepare: narrow tiles - the two epilogue warps of a TMEM lane quarter take alternate
                            // TILES (all chunks of them) instead of alternate 16-column chunks of every tile
    // ---- K loop
    int n_taps, cpt;        // taps, 64-wide chunks per tap
    int kk_last;            // 16-channel MMA steps needed in the last chunk of a tap (0 = all four)
    int tap_dx[TG_MAX_TAPS], tap_dy[TG_MAX_TAPS];
    const int* a_c0;        // [n_tiles] first input channel of each N tile (nullptr => 0)
    // optional K-chunk skip list (plain schedule, kk_last == 0, taps within +-1): for N tile n only the chunks named by
    // kc_list[n * n_taps*cpt + j], j < kc_cnt[n] (>= 1), are visited - the others hold nothing but zeros in Bm
    // (block-sparse adjacency operand). Entry = dx+1 | (dy+1) << 2 | chunk-in-tap << 4 | chunk << 16.
    const int* kc_list;
    const int* kc_cnt;
    int cg;                 // CTA group: 0 = let tapgemm_prepare choose, 1 = single CTA, 2 = CTA pair (cta_group::2)
    int b_resident;         // set by tapgemm_prepare: weights of the current N tile stay resident in shared memory
    int reuse;              // 1: row-reuse schedule (3x3 stride-1 taps; Bm chunks in (dx, cc, dy) order)
                            // 2: patch schedule (3x3 stride-1 taps; Bm chunks in (cc, dy, dx) order): ONE activation box per
                            //    64-channel chunk carries the whole (h_box+2) x (w_box+2) halo patch (w_box = 8) and serves all
                            //    nine taps through descriptor start offsets (a row pitch of w_box+2 pixels = the SBO field)
    int stat_private;       // set by tapgemm_prepare: the fused statistics use per-warp private shared-memory slots
    int epi_fast;           // set by tapgemm_prepare: the launch qualifies for the specialised conv epilogue
    int acc_stages;         // set by tapgemm_prepare: TMEM accumulator stages (2, 4 or 8; 512 / acc_stages columns each)
    int a_stages;           // set by tapgemm_prepare (patch schedule): depth of the activation ring (weights have their own)
    int b_taps;             // set by tapgemm_prepare (patch schedule): taps per weight stage, 3 (one vertical offset) or 9 (all)
    int a_prefetch;         // patch schedule: activation boxes of the tile this many tile slots ahead are prefetched into L2 (0 = off)
    int tile_chunked;       // set by tapgemm_prepare: every CTA (pair) walks ONE contiguous range of tile ids instead of a
                            // strided one, so consecutive tiles share the N tile and resident weights are loaded once
    int a_zmul;             // A dim-3 coordinate = bb + z * a_zmul
    int b_zmul, b_bbmul;    // B dim-2 coordinate = z * b_zmul + bb * b_bbmul
    // ---- epilogue
    int N;                  // valid output columns (multiple of 16)
    void* out;              // 16-bit (operand dtype) or fp32
    int out_fp32;
    long long out_sx, out_sy, out_sb;  // element strides of the output per pixel x / row y / sample (dense: ldc, W*ldc, H*W*ldc)
    long long c_zstride;
    const float* bias;      // [N] or nullptr
    const void* residual;   // same layout/dtype(16-bit) as out, or nullptr
    float alpha;
    // optional softmax-backward epilogue (attention.py:367 differentiated): `residual` then points at the probabilities P
    // (layout of the output) and the stored value is P[m,n] * (D[m,n] - sb_d[row m]) * sb_scale, i.e. dS from dP = D without
    // dP ever reaching memory; sb_d[row] = sum_n dP[row,n] P[row,n] is given as sum_c dO[row,c] O[row,c]
    const float* sb_d;
    float sb_scale;
    // optional row-softmax epilogues (attention.py:367 without the fp32 score matrix in memory). Pass 1, rs_out set: nothing
    // is stored; per (row, N tile, epilogue half) the running {max, sum exp(. - max)} of the tile's columns goes to
    // rs_out[(row * n_tiles + n_tile) * 2 + half]. Pass 2, rs_in set: the same product again, stored as
    // exp(D[m,n] - rs_in[row].x) * rs_in[row].y, i.e. the probabilities, with rs_in = {row max, 1 / row sum} combined from
    // the partials of pass 1 in a fixed order.
    float2* rs_out;
    const float2* rs_in;
    acc_t* rowdot;          // optional: rowdot[pixel] += sum_n D[m,n] * dotsrc[pixel*ld_dot + n] (fixed point 2^40)
    const void* dotsrc;
    long long ld_dot;
    acc_t* gn_sums;         // optional: per (sample, gn_group) {sum, sumsq} of the stored output (fixed point 2^20)
    int gn_cpg, gn_G;       // channels per GroupNorm group, number of groups
    // optional fused GroupNorm(+SiLU) backward reduction over (stored output dy, saved activation x with the output's
    // layout): red[b, g] += {sum(du*gamma), sum(du*gamma*xhat)}; needs gnb_cpg % 16 == 0
    acc_t* gnb_red;           // fixed point 2^40
    const void* gnb_x;
    const acc_t* gnb_stats;   // forward {sum, sumsq} per (sample, group), fixed point 2^20
    const float* gnb_gamma;
    const float* gnb_beta;
    int gnb_cpg, gnb_silu;
    float gnb_inv_n, gnb_eps;
    // optional A-operand transform (patch schedule only): the activation tensor holds the RAW input x of a GroupNorm(+SiLU)
    // layer (reference resnet.py:461-462, 483-489; wf_vae.py:228-229). Four transform warps normalise every landed halo
    // box in shared memory - y = act((x - mean) * rstd * gamma + beta), zero outside the image (the conv pads y, not x) -
    // before the MMAs read it, so the normalised activation never exists in HBM.
    const acc_t* ax_stats;    // forward {sum, sumsq} per (sample, GroupNorm group), fixed point 2^20; nullptr = no transform
    const float* ax_gamma;    // [a_C]
    const float* ax_beta;     // [a_C]
    int ax_cpg, ax_silu;      // channels per GroupNorm group (multiple of 8); 1 = SiLU after the affine map
    int ax_hw;                // pixels of the whole plane (statistics are over ax_cpg x ax_hw elements)
    float ax_eps;
    int ax_y0, ax_y1;         // rows [ax_y0, ax_y1) of the A view are image rows (row-band plans: halo rows of the neighbour
                              // bands are image rows too, the zeroed halo at the map border is not)
    int is_bf16;            // 1: bf16 operands/outputs, 0: fp16
    long long* dbg;         // optional [8]: per-role cycle counters of CTA 0 (test hook only)
    int dbg_skip;           // diagnosis only (WFD_TG_SKIP, results are garbage): bit 0 no activation TMA, bit 1 no weight TMA,
                            // bit 2 no MMA, bit 3 no epilogue loads/stores - isolates which engine bounds a launch
    // ---- raw views (used only by the CUDA-core checking kernel)
    const void* a_ptr; int a_C, a_W, a_H, a_B; long long a_ld;
    const void* b_ptr; long long ldb, b_zstride; int b_rows, b_Z;
};

struct TapGemmOp {
    CUtensorMap tmA, tmB;
    TapGemmParams p;
    int stages;
    int threads;   // block size: TG_THREADS, + TG_XF_THREADS when the operand transform is on
    size_t smem_bytes;
    char tag[12];  // launch class for the per-launch profile (dense, g4, g16, g64, conv_out, attn, wfc, down); may be empty
};

// Encodes the two tensor maps and fills stages/smem. Returns 0 or a negative error (message via set_error).
int tapgemm_prepare(TapGemmOp* op);
// whether the row-reuse schedule can serve this contraction (tap set, tile and shared-memory budget)
bool tapgemm_reuse_ok(const TapGemmParams& p);
// Launches on `stream` (persistent grid of min(#tiles, #SMs) CTAs).
int tapgemm_launch(const TapGemmOp* op, cudaStream_t stream);
// CUDA-core restatement of the same contraction; test/debug only (WFD_DEBUG_REF_GEMM=1 or the selftest entry).
int tapgemm_launch_ref(const TapGemmOp* op, cudaStream_t stream);

void set_error(const char* fmt, ...);
const char* last_error();
extern long long g_launch_count;
void profile_enable(int on);
int profile_report(char* buf, size_t cap);
struct ProfScope {
    cudaStream_t stream;
    int idx;
    ProfScope(const char* name, cudaStream_t s);
    ~ProfScope();
};
int debug_sync_enabled();
int debug_sync_check(const char* what, cudaStream_t stream);
int num_sms();

}  // namespace wfd
