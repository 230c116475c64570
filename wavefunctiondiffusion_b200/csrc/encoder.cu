// Tile-VAE encoder plan (forward only): AutoencoderTile.encode = quant_conv(EncoderTile(x)) (reference wf_vae.py:582-590,
// EncoderTile :68-148), the img2img entry of the pipelines for a tile map given as a wave (pipeline_wfd_img2img.py:68-74,
// 478-486). Same building blocks as the decoder (decoder.cu): grouped 3x3 convs and linears on the tapgemm kernel,
// GroupNorm(8)+SiLU passes, the single-head attention of the mid block - in the encoder's order:
//   conv_in (grouped, T_pad -> C0; on a one-hot input: a gather) -> down blocks of `layers_per_block` resnets ->
//   mid block (resnet, attention, resnet) -> GroupNorm + SiLU -> conv_out (3x3, -> 2 * latent) -> quant_conv (1x1)
// conv_out and quant_conv are both linear with nothing in between: they are folded into ONE 3x3 conv on the host.
#include "decoder.cuh"
#include "kernels.cuh"
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace wfd {

#define CK(expr)                      \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ < 0) return rc__;    \
    } while (0)
#define CUDA_CK(expr)                                                           \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) {                                               \
            set_error("%s: %s", #expr, cudaGetErrorString(e__));                \
            return -8;                                                          \
        }                                                                       \
    } while (0)

Encoder::~Encoder() {
    if (gather) cudaFree(gather);
}

int Encoder::init_enc(const wfd_decoder_config& c, int H_, int W_, int Bmax, int dtype16) {
    cfg = c;
    encoder_side = true;
    h = H_;
    w = W_;
    h_total = H_;
    B_max = Bmax;
    bf16 = dtype16 ? 1 : 0;
    G = c.norm_num_groups;
    {
        const char* e = getenv("WFD_NO_REUSE");
        use_reuse = (e && e[0] == '1') ? 0 : 1;
        const char* f = getenv("WFD_NO_FUSE_GN");
        fuse_gn = (f && f[0] == '1') ? 0 : 1;
        const char* np = getenv("WFD_NO_PATCH");
        use_patch = (np && np[0] == '1') ? 0 : 1;
        const char* na = getenv("WFD_FUSE_APPLY");
        fuse_apply = na ? atoi(na) : 1;
        const char* ff = getenv("WFD_NO_FUSE_ATTN_FWD");
        fuse_attn_fwd = (ff && ff[0] == '1') ? 0 : 1;
    }
    const int nb = c.n_blocks;
    if (c.latent_channels != 4 || nb < 1 || nb > 8 || Bmax < 1 || h < 1 || w < 1) {
        set_error("encoder: bad config (latent_channels=%d n_blocks=%d B_max=%d plane %dx%d)", c.latent_channels, nb, Bmax, h, w);
        return -3;
    }
    for (int i = 0; i < nb; ++i)
        if (c.block_down[i]) {
            set_error("encoder: only configurations whose encoder blocks keep the resolution are supported");
            return -3;
        }
    C_in = c.out_channels;  // the wave's channel count (in_channels == out_channels == T_pad for the tile VAE)
    C0 = c.block_out_channels[0];
    Cm = c.block_out_channels[nb - 1];
    cin_g = C_in / c.conv_init_groups;
    cout_g = C0 / c.conv_init_groups;
    if (C_in % c.conv_init_groups || C0 % c.conv_init_groups || C_in % 8 || cout_g > 128) {
        set_error("encoder: conv_in %d -> %d with %d groups is not supported", C_in, C0, c.conv_init_groups);
        return -3;
    }
    resnets.clear();
    int ch = C0;
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[i], groups = c.conv_groups[i];
        for (int j = 0; j < c.layers_per_block; ++j) {
            ResnetDesc r;
            char buf[128];
            snprintf(buf, sizeof(buf), "encoder.down_blocks.%d.resnets.%d", i, j);
            r.prefix = buf;
            r.cin = j == 0 ? ch : cout;
            r.cout = cout;
            r.groups = groups;
            r.h = h;
            r.w = w;
            resnets.push_back(r);
        }
        ch = cout;
    }
    for (int j = 0; j < 2; ++j) {
        ResnetDesc r;
        r.prefix = "encoder.mid_block.resnets." + std::to_string(j);
        r.cin = r.cout = Cm;
        r.groups = 1;
        r.h = h;
        r.w = w;
        resnets.push_back(r);
    }
    for (const ResnetDesc& r : resnets)
        if (r.cin % 8 || r.cout % 16 || (r.cin / G) % 8 || (r.cout / G) % 8 || r.cin % r.groups || r.cout % r.groups) {
            set_error("encoder: unsupported channel counts in %s (%d -> %d, groups %d)", r.prefix.c_str(), r.cin, r.cout, r.groups);
            return -3;
        }
    C_last = Cm;
    H_out = h;
    W_out = w;
    T_pad = C_in;
    int wb, hb;
    if (!pick_box(w, h, &wb, &hb) || (static_cast<long long>(h) * w) % 128) {
        set_error("encoder: plane %dx%d cannot be tiled into 128-pixel patches", h, w);
        return -3;
    }
    // workspace: forward only, so the resnet outputs ping-pong between two buffers and share one conv1 buffer
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~size_t(255);
        return o;
    };
    const size_t e = 2, B = static_cast<size_t>(B_max), hw = static_cast<size_t>(h) * w;
    int maxC = std::max(C0, Cm);
    for (const ResnetDesc& r : resnets) maxC = std::max(maxC, std::max(r.cin, r.cout));
    o_wave = take(B * hw * C_in * e);
    o_x0 = take(B * hw * C0 * e);
    o_h1 = take(B * hw * maxC * e);
    o_ping[0] = take(B * hw * maxC * e);
    o_ping[1] = take(B * hw * maxC * e);
    o_res_h1.assign(resnets.size(), o_h1);
    o_res_out.clear();
    for (size_t i = 0; i < resnets.size(); ++i) o_res_out.push_back(o_ping[i & 1]);
    o_attn_out = take(B * hw * Cm * e);
    o_xn = take(B * hw * Cm * e);
    o_qkv = take(B * hw * 3 * Cm * e);
    o_vt = take(B * Cm * hw * e);
    o_S = take(B * hw * hw * 4);
    o_RS = take(B * hw * 8);
    o_P = take(B * hw * hw * e);
    o_O = take(B * hw * Cm * e);
    o_act = take(B * hw * maxC * e);
    o_act2 = take(B * hw * maxC * e);
    o_sc = take(B * hw * maxC * e);
    o_mom = take(B * hw * 16 * 4);
    n_gn = static_cast<int>(resnets.size()) * 2 + 2;
    o_gn = take(static_cast<size_t>(n_gn) * 2 * B * G * 2 * sizeof(acc_t));
    ws_bytes = off;
    down_after.assign(nb, 0);
    return 0;
}

int Encoder::finalize_enc() {
    // conv_in: the tapgemm form (dense wave input) and the gather table (one-hot input)
    CK(pack_named_conv("encoder.conv_in", C0, C_in, cfg.conv_init_groups, 3, false, 0, h, w));
    {
        const std::vector<float>* wv;
        CK(need("encoder.conv_in.weight", static_cast<long long>(C0) * cin_g * 9, &wv));
        std::vector<float> tab(static_cast<size_t>(C_in) * 9 * cout_g);
        for (int t = 0; t < C_in; ++t) {
            const int g = t / cin_g, ci = t % cin_g;
            for (int tap = 0; tap < 9; ++tap)
                for (int j = 0; j < cout_g; ++j)
                    tab[(static_cast<size_t>(t) * 9 + tap) * cout_g + j] =
                        (*wv)[(static_cast<size_t>(g * cout_g + j) * cin_g + ci) * 9 + tap];
        }
        CK(upload_f32(tab, &gather));
    }
    for (const ResnetDesc& r : resnets) {
        CK(upload_vec(r.prefix + ".norm1.weight", r.cin));
        CK(upload_vec(r.prefix + ".norm1.bias", r.cin));
        CK(upload_vec(r.prefix + ".norm2.weight", r.cout));
        CK(upload_vec(r.prefix + ".norm2.bias", r.cout));
        CK(pack_named_conv(r.prefix + ".conv1", r.cout, r.cin, r.groups, 3, false, 0, r.h, r.w));
        CK(pack_named_conv(r.prefix + ".conv2", r.cout, r.cout, r.groups, 3, false, 0, r.h, r.w));
        if (r.cin != r.cout) CK(pack_named_conv(r.prefix + ".conv_shortcut", r.cout, r.cin, r.groups, 1, false, 0, 0, 0));
    }
    CK(pack_attention("encoder.mid_block.attentions.0.", false));
    CK(upload_vec("encoder.conv_norm_out.weight", Cm));
    CK(upload_vec("encoder.conv_norm_out.bias", Cm));
    {
        // moments = quant_conv(conv_out(x)): W'[o] = sum_m Wq[o, m] W[m], b' = Wq b + bq; padded to 16 output rows
        const int M = 2 * cfg.latent_channels;
        const std::vector<float>*wo, *bo, *wq, *bq;
        CK(need("encoder.conv_out.weight", static_cast<long long>(M) * Cm * 9, &wo));
        CK(need("encoder.conv_out.bias", M, &bo));
        CK(need("quant_conv.weight", static_cast<long long>(M) * M, &wq));
        CK(need("quant_conv.bias", M, &bq));
        const size_t per = static_cast<size_t>(Cm) * 9;
        std::vector<float> wf(16 * per, 0.f), bf(16, 0.f);
        for (int o = 0; o < M; ++o) {
            double bb = (*bq)[o];
            for (int m = 0; m < M; ++m) {
                const double q = (*wq)[o * M + m];
                bb += q * (*bo)[m];
                for (size_t i = 0; i < per; ++i) wf[o * per + i] += static_cast<float>(q * (*wo)[m * per + i]);
            }
            bf[o] = static_cast<float>(bb);
        }
        ConvW& cw = convs["encoder.conv_out+quant_conv"];
        cw.cin = Cm;
        cw.cout = 16;
        // same K-loop schedule rule as every other 3x3 conv of the plan
        const int sched = (use_patch && w % 8 == 0 && h % 16 == 0) ? 2 : 0;
        CK(pack_conv(wf.data(), 16, Cm, 1, 3, 0, bf16 != 0, 0, &cw.fwd, nullptr, sched));
        CK(upload_f32(bf, &cw.bias));
    }
    finalized = true;
    host.clear();
    if (ws) return prepare_enc();
    return 0;
}

int Encoder::bind_workspace_enc(void* p, size_t bytes) {
    if (bytes < ws_bytes || (reinterpret_cast<uintptr_t>(p) & 255)) {
        set_error("encoder: workspace too small or not 256-byte aligned (%zu < %zu)", bytes, ws_bytes);
        return -3;
    }
    ws = static_cast<uint8_t*>(p);
    if (finalized) return prepare_enc();
    return 0;
}

int Encoder::prepare_enc() {
    fwd_ops.clear();
    Exec ex;
    ex.prepare = true;
    ex.B = B_max;
    ex.stream = 0;
    ex.ops = &fwd_ops;
    CK(forward_enc(ex, nullptr, 0, nullptr, nullptr));
    prepared = true;
    return 0;
}

// wave != nullptr: dense input (NCHW fp32 / fp16); tiles != nullptr: one-hot input given as integer ids [B, H, W]
int Encoder::forward_enc(Exec& ex, const void* wave, int wave_dtype, const long long* tiles, float* moments) {
    const int B = ex.B;
    cudaStream_t s = ex.stream;
    const float eps = 1e-6f;
    const int hw = h * w;
    uint8_t* x = ws + o_x0;
    if (!ex.prepare) CUDA_CK(cudaMemsetAsync(ws + o_gn, 0, static_cast<size_t>(n_gn) * B_max * G * 2 * sizeof(acc_t), s));
    const ConvW& ci = convs["encoder.conv_in"];
    cur_tag = "conv_in";
    bool stats0 = false;
    if (ex.prepare || wave) {
        if (!ex.prepare) CK(nchw_to_nhwc16(wave, wave_dtype, ws + o_wave, B, C_in, hw, bf16, s));
        CK(conv_gemm(ex, ci.fwd, ws + o_wave, C_in, h, w, h, w, x, C0, ci.bias, nullptr, gn_sums_ptr(0), C0 / G, nullptr, nullptr));
        stats0 = true;
    } else {
        ex.cursor = 1;  // the recorded launch sequence starts with the dense conv_in: skip it
        CK(onehot_conv_in(tiles, gather, ci.bias, x, B, h, w, C_in, C0, cin_g, cout_g, bf16, s));
    }
    const int n_res = static_cast<int>(resnets.size());
    for (int idx = 0; idx < n_res - 2; ++idx) {
        CK(resnet_forward(ex, idx, x, idx == 0 ? stats0 : true, 2 * (idx + 1)));
        x = ws + o_res_out[idx];
    }
    CK(resnet_forward(ex, n_res - 2, x, true, n_gn - 2));
    x = ws + o_res_out[n_res - 2];
    CK(attention_forward(ex, x, "encoder.mid_block.attentions.0.", 2 * (n_res - 1)));
    x = ws + o_attn_out;
    CK(resnet_forward(ex, n_res - 1, x, true, n_gn - 1));
    x = ws + o_res_out[n_res - 1];
    auto fused_ok = [&](int C) { return fuse_gn && (C / G) % 16 == 0; };
    if (!ex.prepare) {
        if (!fused_ok(Cm)) CK(gn_stats(x, gn_sums_ptr(n_gn - 1), B, hw, Cm, G, bf16, s));
        CK(gn_apply(x, gn_sums_ptr(n_gn - 1), vecs["encoder.conv_norm_out.weight"], vecs["encoder.conv_norm_out.bias"],
                    ws + o_act, B, hw, Cm, G, eps, 1, bf16, s, hw));
    }
    const ConvW& co = convs["encoder.conv_out+quant_conv"];
    cur_tag = "enc_out";
    next_out_fp32 = true;
    CK(conv_gemm(ex, co.fwd, ws + o_act, Cm, h, w, h, w, ws + o_mom, 16, co.bias, nullptr, nullptr, 0, nullptr, nullptr));
    if (!ex.prepare)
        CK(moments_out(reinterpret_cast<const float*>(ws + o_mom), moments, B, 2 * cfg.latent_channels, hw, 16, s));
    return 0;
}

int Encoder::encode_wave(const void* wave_nchw, int dtype, int B, float* moments, cudaStream_t s) {
    if (!prepared || B < 1 || B > B_max) {
        set_error("encoder: not prepared or batch %d outside [1, %d]", B, B_max);
        return -12;
    }
    Exec ex;
    ex.B = B;
    ex.stream = s;
    ex.ops = &fwd_ops;
    return forward_enc(ex, wave_nchw, dtype, nullptr, moments);
}

int Encoder::encode_tiles(const long long* tiles, int B, float* moments, cudaStream_t s) {
    if (!prepared || B < 1 || B > B_max) {
        set_error("encoder: not prepared or batch %d outside [1, %d]", B, B_max);
        return -12;
    }
    Exec ex;
    ex.B = B;
    ex.stream = s;
    ex.ops = &fwd_ops;
    return forward_enc(ex, nullptr, 0, tiles, moments);
}

}  // namespace wfd
