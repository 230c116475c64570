// HBM-bound kernels of the WFC-guidance path (everything that is not a tensor-core contraction):
// GroupNorm(+SiLU) forward/backward, row softmax, WFC finishing passes, argmax, layout conversion, the 4-channel
// latent-side convs. All tensors are channels-last: [B, H*W, C] with 16-bit elements unless stated otherwise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tapgemm.cuh"  // acc_t, fixed-point scales

namespace wfd {

// ---- GroupNorm (8 groups in the tile VAE; reference resnet.py:407,423, attention.py:278, wf_vae.py:212), eps 1e-6
// sums: [B, G, 2] fixed-point (2^20) {sum, sumsq}, zeroed by the caller before accumulation.
int gn_stats(const void* x, acc_t* sums, int B, int HW, int C, int G, int bf16, cudaStream_t s);
// y = act((x - mean) * rstd * gamma + beta), act = SiLU if silu else identity
// HW_stat: pixels per sample the statistics cover (0 = HW; larger when the plane is split into row bands)
int gn_apply(const void* x, const acc_t* sums, const float* gamma, const float* beta, void* y, int B, int HW, int C,
             int G, float eps, int silu, int bf16, cudaStream_t s, int HW_stat = 0);
// red: [B, G, 2] fixed-point (2^40) {sum(du*gamma), sum(du*gamma*xhat)}, zeroed by the caller; du = dy * act'(u)
int gn_bwd_reduce(const void* dy, const void* x, const acc_t* sums, const float* gamma, const float* beta, acc_t* red,
                  int B, int HW, int C, int G, float eps, int silu, int bf16, cudaStream_t s, int HW_stat = 0);
// dx = rstd * (du*gamma - (red0 + xhat*red1)/n) (+ add)
int gn_bwd_apply(const void* dy, const void* x, const acc_t* sums, const float* gamma, const float* beta,
                 const acc_t* red, const void* add, void* dx, int B, int HW, int C, int G, float eps, int silu,
                 int bf16, cudaStream_t s, int HW_stat = 0);

// reduce (optional: skipped when the producing contraction fused it) + apply, in L2-sized blocks of samples
int gn_bwd(const void* dy, const void* x, const acc_t* sums, const float* gamma, const float* beta, acc_t* red,
           const void* add, void* dx, int B, int HW, int C, int G, float eps, int silu, int bf16, int do_reduce,
           cudaStream_t s);

// ---- latent side (reference wf_vae.py:593 post_quant_conv 1x1 4->4, wf_vae.py:218 conv_in 3x3 4->Cm)
// z: NCHW [B,4,h,w] fp32 (z_f16=0) or fp16 (z_f16=1); out: [B, h*w, Cm] 16-bit. pq_w [4,4], pq_b [4], w [Cm,4,3,3], b [Cm]
int latent_in(const void* z, int z_f16, float in_scale, const float* pq_w, const float* pq_b, const float* w,
              const float* b, void* out, int B, int h, int wd, int Cm, int bf16, cudaStream_t s, int y_off = 0,
              int h_total = 0);  // row bands: z is the whole [B,4,h_total,w] plane, rows [y_off, y_off+h) are produced
// dz[b,c,y,x] (fp32 NCHW) = in_scale * pq_w^T conv_in^T(dy); dy: [B, h*w, Cm] 16-bit; scale_b: optional per-sample scale;
// w must be the conv_in weight re-ordered to [ky][kx][ci][co]
int latent_out(const void* dy, float in_scale, const float* scale_b, const float* pq_w, const float* w, float* dz,
               int B, int h, int wd, int Cm, int bf16, cudaStream_t s, int y_off = 0, int h_total = 0);
// gathers the nine taps of conv_in^T from the tensor-core product du[pixel][tap*4 + ci] (row pitch ld floats), applies
// post_quant_conv^T, in_scale and the per-sample scale
int latent_gather(const float* du, int ld, float in_scale, const float* scale_b, const float* pq_w, float* dz, int B, int h,
                  int wd, cudaStream_t s);

// ---- wave softmax over all T_pad channels (reference pipeline_wfd.py:405); rowsum[n] = sum_{i<T} p~[n,i] of the
// rounded 16-bit values that feed the contraction.
int wave_softmax(const void* logits, void* wave, float* rowsum, long long rows, int T_pad, int T, int bf16,
                 cudaStream_t s);
// rowsum only (wave given, e.g. WFCLossEinsum(wave) called directly)
int wave_rowsum(const void* wave, float* rowsum, long long rows, int T_pad, int T, int bf16, cudaStream_t s);
// loss[b] = coef * 0.5 * sum_n (S_n * c_n - rowdot_a[n]),  c_n = sum of S over the <=4 in-map neighbours
int wfc_loss_finish(const float* rowsum, const acc_t* rowdot, float* loss, int B, int H, int W, float coef,
                    cudaStream_t s, int y_off = 0, int H_total = 0);
// grad wrt wave (no softmax): dwave[n,i] = coef*dloss[b] * ([i<T]*c_n - a[n,i]);
// grad wrt logits (softmax bwd): dlogits[n,i] = coef*dloss[b] * p[n,i] * (([i<T]*c_n - a[n,i]) - (S_n*c_n - rowdot_a[n]))
int wfc_grad_finish(const void* wave, const void* a, const float* rowsum, const acc_t* rowdot, const float* dloss,
                    void* dout, int B, int H, int W, int T_pad, int T, float coef, int through_softmax, int bf16,
                    cudaStream_t s, int y_off = 0, int H_total = 0);

// ---- attention helpers (reference attention.py:354-368: scores -> fp32 softmax -> probs)
int attn_softmax(const float* scores, void* probs, long long rows, int n, int bf16, cudaStream_t s);
// ds = p * (dp - sum_j dp*p) * scale
int attn_softmax_bwd(const void* probs, const float* dp, void* ds, long long rows, int n, float scale, int bf16,
                     cudaStream_t s);
// d[row] = sum_c a[row, c] * b[row, c] (fp32) over two 16-bit [rows, C] matrices
int rowdot16(const void* a, const void* b, float* d, long long rows, int C, int bf16, cudaStream_t s);
// merges the per-(N tile, half) {max, sum exp} partials of every row into {row max, 1 / row sum} (float2 in, float2 out)
int rowstats_combine(const void* part, void* out, long long rows, int n_part, cudaStream_t s);
// batched 16-bit transpose: in [Z, R, C] (row stride ld_in) -> out [Z, C, R]
int transpose16(const void* in, void* out, int Z, int R, int C, long long ld_in, long long zstride_in, cudaStream_t s);

// ---- argmax over the last axis, numpy semantics (first max wins, NaN is max); dtype: 0 fp32, 1 fp16, 2 bf16
int argmax_rows(const void* x, int dtype, long long rows, int n, long long* out, cudaStream_t s);

// ---- rule violations of an integer tile map (the one-hot case of the loss): out[b] = {pairs (x, x+1) with
// mat_x[left, right] == 0, pairs (y, y+1) with mat_y[top, bottom] == 0}; tile ids outside [0, T) always violate.
// tiles: int64 [B, H, W]; mat_x / mat_y: uint8 [T, T] on the device; out: int64 [B, 2], overwritten.
int count_violations(const long long* tiles, const uint8_t* mat_x, const uint8_t* mat_y, int T, int B, int H, int W,
                     long long* out, cudaStream_t s);

// ---- what the reference's callers do with the argmax tiles on the host (ui.py:228-243): optional mirror symmetry
// [tiles | symm[tiles reversed along x]] (data_processing.py:191-196), then sc = gen_to_sc[shrink_to_gen[tile]]
// (map_display.py:55). tiles int64 [B, H, W]; tiles_out int64 [B, H, W'] / sc_out uint16 [B, H, W'] (either may be
// null), W' = 2 W with symmetry; *bad counts ids that fall outside a table (padded channels): they map to 0xFFFF.
int tiles_finish(const long long* tiles, int B, int H, int W, const int* symm, int n_symm, const int* shrink_to_gen,
                 int n_shrink, const int* gen_to_sc, int n_gen, long long* tiles_out, uint16_t* sc_out, long long* bad,
                 cudaStream_t s);

// ---- encoder ends: conv_in on a one-hot input as a gather (table [C_in][9][cout_g] fp32), moments export
int onehot_conv_in(const long long* tiles, const float* table, const float* bias, void* out, int B, int H, int W,
                   int C_in, int C0, int cin_g, int cout_g, int bf16, cudaStream_t s);
int moments_out(const float* in, float* out, int B, int C, int HW, int ld, cudaStream_t s);

// ---- elementwise glue of the guidance loop for affine schedulers (pipeline_wfd.py:601-603, 611-614, 630-637):
// z = a lat + b (eps_u + g (eps_t - eps_u)) (eps_t may be null: eps = eps_u), and lat -= a dz
int guidance_affine(const float* lat, const float* eps_u, const float* eps_t, float g, float a, float b, float* z,
                    long long n, cudaStream_t s);
int guidance_update(float* lat, const float* dz, float a, long long n, cudaStream_t s);

// ---- layout conversion at the API boundary: NCHW (fp32/fp16) <-> channels-last 16-bit
int nchw_to_nhwc16(const void* in, int in_f16, void* out, int B, int C, int HW, int bf16, cudaStream_t s);
int nhwc16_to_nchw(const void* in, void* out, int out_f16, int B, int C, int HW, int bf16, cudaStream_t s);
// 16-bit -> fp32 (same layout) and generic helpers
int cvt16_to_f32(const void* in, float* out, long long n, int bf16, cudaStream_t s);
int cvt_f32_to_16(const float* in, void* out, long long n, int bf16, cudaStream_t s);
// strided rows: out[r, c] (row stride ld_out, 16-bit) = in[r, c] (row stride ld_in, fp32)
int cvt_f32_to_16_2d(const float* in, long long ld_in, void* out, long long ld_out, long long rows, int cols, int bf16,
                     cudaStream_t s);
int add16(const void* a, const void* b, void* out, long long n, int bf16, cudaStream_t s);

}  // namespace wfd
