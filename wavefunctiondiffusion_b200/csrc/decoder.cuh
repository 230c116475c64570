// Decoder plan and WFC handle (internal C++ side of the C ABI in include/wfd_b200.h).
#pragma once
#include "../../include/wfd_b200.h"
#include "comm.cuh"
#include "tapgemm.cuh"
#include <map>
#include <string>
#include <vector>

namespace wfd {

void debug_set_ref_gemm(int on);
int debug_ref_gemm();

// A K-major B operand for tapgemm: [rows = n_tiles*BN][ldb = n_taps*cpt*64] 16-bit, zero outside each group's block.
struct PackedB {
    uint16_t* Bm = nullptr;
    int* a_c0 = nullptr;  // first A channel per N tile
    int BN = 0, n_tiles = 0, cpt = 0, n_taps = 0, N = 0, rows = 0, in_stride = 1, reuse = 0, kk_last = 0;
    long long ldb = 0;
    int dx[TG_MAX_TAPS] = {0}, dy[TG_MAX_TAPS] = {0};
    void release();
};
struct TapSel {  // one tap of a contraction: which weight tap it reads and where it shifts the input
    int w_tap, dx, dy;
};
struct GnBwdFuse {  // GroupNorm(+SiLU) backward reduction fused into the producing contraction's epilogue
    acc_t* red;
    const void* x;
    const acc_t* stats;
    const float* gamma;
    const float* beta;
    int cpg, silu;
    float inv_n, eps;
};
struct AxForm {  // GroupNorm(+SiLU) applied to the A operand inside the consuming contraction (TapGemmParams::ax_*)
    const acc_t* stats;
    const float* gamma;
    const float* beta;
    int cpg, silu, hw;
    float eps;
};
struct OutStrides {  // element strides of a strided output plane (downsample backward writes every other pixel)
    long long sx, sy, sb;
};
// reuse: chunk order of the packed weights = K-loop schedule of the contraction (0 plain, 1 row reuse, 2 patch; tapgemm.cuh)
int pack_conv(const float* w, int cout, int cin, int groups, int k, int mode, bool bf16, int bn_override, PackedB* out,
              const std::vector<TapSel>* tap_sel, int reuse);

struct ConvW {
    PackedB fwd, bwd;
    float* bias = nullptr;
    int cin = 0, cout = 0;
};

struct ResnetDesc {
    std::string prefix;
    int cin, cout, groups, h, w;
};

struct Exec {
    bool prepare = false;
    int B = 0;
    cudaStream_t stream = 0;
    std::vector<TapGemmOp>* ops = nullptr;
    size_t cursor = 0;
};

struct PlainGemm;
int upload_f32(const std::vector<float>& v, float** dev);
bool pick_box(int W, int H, int* w_box, int* h_box);

struct Decoder {
    wfd_decoder_config cfg;
    int h = 0, w = 0, B_max = 0, bf16 = 1, G = 8;
    int Cm = 0, C_last = 0, H_out = 0, W_out = 0, T_pad = 0;
    int fuse_gn = 1;
    int conv_n_group = 0;  // tile rasterisation of the convs: 0/1 = M fastest (default), k = bands of k N tiles (WFD_CONV_NGROUP)
    int fuse_bwd_max_c = 768;  // largest channel count whose GroupNorm-backward reduction is fused into the conv epilogue
    bool mmajor_a = true;  // attention backward reads P / dS as stored (M-major A operand) instead of transposed copies
    int reuse_hbox = 0;    // preferred patch height of the row-reuse contractions (WFD_REUSE_HBOX: 0 = widest patch, 2, 4, 8, 16)
    int use_reuse = 1;  // row-reuse schedule for 3x3 convs whose stage fits (WFD_NO_REUSE=1 disables, for A/B runs)
    int fuse_attn_fwd = 1;  // row softmax inside two passes of the score product (WFD_NO_FUSE_ATTN_FWD=1: fp32 scores + kernel)
    int fuse_attn_bwd = 1;  // softmax backward in the epilogue of dP = dO V^T (WFD_NO_FUSE_ATTN_BWD=1: separate fp32 dP + kernel)
    int tc_latent = 1;  // conv_in^T contracts its Cm channels on the tensor cores (WFD_NO_TC_LATENT=1: the CUDA-core kernel)
    int use_patch = 1;  // patch schedule for 3x3 convs on planes that tile into 16 x 8 patches (WFD_NO_PATCH=1 disables)
    int fuse_apply = 1; // GroupNorm+SiLU applied in the consuming patch-schedule conv's operand path instead of a standalone
                        // gn_apply pass (WFD_FUSE_APPLY): 0 = never, 1 = where a box's MMAs outlast its transform (N tiles of
                        // >= 192 columns: the dense and groups=4 layers), 2 = every patch-schedule conv. Measured on B200
                        // (profiles/r02_summary.md): the transform costs ~2000 issue cycles per 18 x 10 x 64 box whatever the
                        // number of transform warps (tanh and the 16-bit pack go through the quarter-rate special-function
                        // unit), more than the MMAs of a narrow-N box take, so mode 2 is slower than the standalone pass
    // row bands (one map split along y over band_R ranks): h is the local band height, h_total the map's
    bool band = false;
    int band_R = 1, band_r = 0, h_total = 0;
    Comm* comm = nullptr;   // not owned
    size_t halo_pad = 0;    // bytes kept free before and after every workspace buffer for the halo rows
    size_t o_dkv = 0;       // fp32 [Nk, 2*Cm] partial dK | dV of this band's queries (reduce-scattered over the bands)
    std::vector<ResnetDesc> resnets;
    std::vector<int> down_after;
    std::map<std::string, std::vector<float>> host;
    std::map<std::string, ConvW> convs;
    std::map<std::string, float*> vecs;
    bool finalized = false, prepared = false;
    // workspace
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;
    size_t o_x0 = 0, o_attn_out = 0, o_xn = 0, o_qkv = 0, o_vt = 0, o_S = 0, o_P = 0, o_O = 0;
    size_t o_RS = 0;        // float2 [B, N] {row max, 1 / row sum} of the fused forward softmax
    size_t o_D = 0;         // fp32 [B, N] row dots sum_c dO * O of the softmax backward
    size_t o_PT = 0, o_dS = 0, o_dST = 0, o_dO = 0, o_dOT = 0, o_KT = 0, o_QT = 0, o_dqkv = 0;
    size_t o_act = 0, o_act2 = 0, o_sc = 0, o_g[4] = {0, 0, 0, 0}, o_logits = 0, o_dlogits = 0, o_gn = 0;
    std::vector<size_t> o_res_h1, o_res_out, o_down_out;
    int n_gn = 0;
    uint8_t* x_last = nullptr;
    float last_in_scale = 1.f;
    int last_B = 0;
    std::vector<TapGemmOp> fwd_ops, bwd_ops;
    bool next_out_fp32 = false;  // the contraction prepared next stores fp32 (the encoder's moments); reset by conv_gemm
    bool encoder_side = false;   // set_weight keeps encoder.* / quant_conv.* instead of decoder.* / post_quant_conv.*
    const char* cur_tag = "";  // launch class stamped on the contractions prepared next (per-launch profile of bench.py)

    ~Decoder();
    int init(const wfd_decoder_config& c, int h, int w, int B_max, int dtype16, int band_rank = -1, int band_nranks = 1);
    void layout_workspace();
    int set_weight(const char* key, const float* data, long long numel);
    int finalize_weights();
    int bind_workspace(void* p, size_t bytes);
    int forward(const void* z, int z_dtype, int B, float in_scale, cudaStream_t s);
    int backward(const void* dlogits, int B, const float* scale_b, float* dz, cudaStream_t s);

    // internals
    int need(const std::string& key, long long numel, const std::vector<float>** out);
    int pack_named_conv(const std::string& name, int cout, int cin, int groups, int k, bool want_bwd, int fwd_mode, int ph,
                        int pw);
    int upload_vec(const std::string& key, long long numel);
    int pack_attention(const std::string& prefix, bool want_bwd);
    int prepare_ops();
    int run_gemm(Exec& ex, TapGemmOp& tmpl, bool z_is_batch);
    int conv_gemm(Exec& ex, const PackedB& pk, const void* A, int a_C, int Hin, int Win, int Hout, int Wout, void* out,
                  long long ldc, const float* bias, const void* residual, acc_t* gn_sums, int gn_cpg,
                  const OutStrides* os, const GnBwdFuse* gb, const AxForm* ax = nullptr);
    // whether the GroupNorm in front of this 3x3 conv is applied inside the conv (patch schedule) or needs gn_apply
    bool apply_fused(const PackedB& pk) const { return pk.reuse == 2 && (fuse_apply == 2 || (fuse_apply == 1 && pk.BN >= 192)); }
    int plain_gemm(Exec& ex, const PlainGemm& g);
    acc_t* gn_sums_ptr(int idx) const;
    acc_t* gn_red_ptr(int idx) const;
    int sync_sums(Exec& ex, acc_t* slot);
    int halo_rows(Exec& ex, void* buf, int rows, int W, int C);
    int gn_backward(Exec& ex, const void* dy, const void* x, int gn_idx, const float* gamma, const float* beta,
                    const void* add, void* dx, int HW, int C, int silu, int do_reduce);
    int resnet_forward(Exec& ex, int idx, uint8_t* xin, bool stats_ready, int next_gn);
    int attention_forward(Exec& ex, uint8_t* x, const std::string& attn_prefix, int next_gn);
    int forward_impl(Exec& ex, const void* z, int z_f16, float in_scale);
    int backward_impl(Exec& ex, const void* dlogits, const float* scale_b, float* dz);
};

// Tile-VAE ENCODER plan (reference wf_vae.py:68-148 EncoderTile + quant_conv :582-590), forward only: the img2img entry of
// the pipelines (pipeline_wfd_img2img.py:478-486) encodes a one-hot tile map into latent moments. Built on the decoder's
// machinery (same contractions, GroupNorm kernels and attention block, in the encoder's order); only configurations whose
// encoder blocks keep the resolution ("EvenEncoderBlock2D") are supported.
struct Encoder : Decoder {
    int C_in = 0, C0 = 0;            // wave channels (T_pad), channels after conv_in
    int cin_g = 0, cout_g = 0;       // conv_in channels per group (in / out)
    size_t o_wave = 0, o_mom = 0, o_ping[2] = {0, 0}, o_h1 = 0;
    float* gather = nullptr;         // [C_in][9][cout_g] fp32: conv_in weights by (input tile, tap) for one-hot input
    ~Encoder();
    int init_enc(const wfd_decoder_config& c, int H, int W, int B_max, int dtype16);
    int finalize_enc();
    int bind_workspace_enc(void* p, size_t bytes);
    int prepare_enc();
    int forward_enc(Exec& ex, const void* wave, int wave_dtype, const long long* tiles, float* moments);
    int encode_wave(const void* wave_nchw, int dtype, int B, float* moments, cudaStream_t s);
    int encode_tiles(const long long* tiles, int B, float* moments, cudaStream_t s);
};

// WFC loss handle: packed adjacency operand [T_pad rows][4 * T_pad] = [Ax | Ax^T | Ay | Ay^T] (0/1), see wfc.cu
struct WfcSavedHeader {  // host-side record of the last forward (what backward needs to know)
    int B = 0, H = 0, W = 0, is_logits = 0;
    float coef = 0.f;
    const void* wave = nullptr;  // the wave operand (inside saved, or the caller's pointer when is_logits == 0)
};

// sliced-ELL lists of the allowed pairs for the CUDA-core form of the contraction (csrc/wfc_sparse.cu)
struct WfcSparse {
    int T = 0, T_pad = 0, n_slices = 0;
    long long nnz = 0, padded = 0;  // allowed pairs over the four directions; list entries including the ELL padding
    uint16_t* ent = nullptr;
    int* slice_off = nullptr;
    uint16_t* row_of = nullptr;
    uint16_t* len_of = nullptr;
    ~WfcSparse();
    int init(const uint8_t* mat_x, const uint8_t* mat_y, int T, int T_pad);
    size_t smem_bytes() const;
    int contract(const void* wave, void* a, acc_t* rowdot, int B, int H, int W, int bf16, cudaStream_t s) const;
};

struct Wfc {
    int T = 0, T_pad = 0, bf16 = 1;
    WfcSparse sparse;         // built on the first set_sparse(1)
    bool use_sparse = false;  // wfd_wfc_set_sparse / WFD_WFC_SPARSE=1
    int set_sparse(int on);
    WfcSavedHeader hdr_host;         // of the last forward
    void* hdr_saved = nullptr;
    // one record per `saved` block that a forward has filled: several forwards may be outstanding before their backwards
    // run (loss(a) + loss(b)), each backward must see the shapes of ITS forward
    std::map<void*, WfcSavedHeader> saved_hdrs;
    uint16_t* Bm = nullptr;
    uint8_t* rules = nullptr;  // device copy of the 0/1 rule matrices [2][T][T] (x then y) for count_violations
    int* kc_list = nullptr;  // K chunks of each N tile that hold at least one allowed pair (see Wfc::init)
    int* kc_cnt = nullptr;
    double kc_kept = 1.0;    // fraction of the K chunks that are visited
    // last prepared contraction (tensor maps depend on the buffers of the call)
    TapGemmOp op;
    const void* op_wave = nullptr;
    void* op_a = nullptr;
    int op_B = 0, op_H = 0, op_W = 0;
    // row bands: this handle scores rows [band_r * H, (band_r + 1) * H) of a map of H_total rows
    bool band = false;
    int band_R = 1, band_r = 0, H_total = 0;
    Comm* comm = nullptr;
    int set_band(Comm* c, int H_total);
    ~Wfc();
    int init(const uint8_t* mat_x, const uint8_t* mat_y, int T, int T_pad, int dtype16);
    size_t saved_bytes(int B, int H, int W) const;
    int forward(const void* in, int is_logits, int B, int H, int W, float strength, void* saved, float* loss,
                cudaStream_t s);
    int backward(void* saved, const float* dloss, void* dout, cudaStream_t s);
};
}  // namespace wfd
