// Tile-VAE decoder plan: weight packing for the tapgemm kernel, forward, backward-data.
// Mirrors AutoencoderTile._decode (reference wf_vae.py:592-599) = post_quant_conv -> DecoderTile.forward (:216-232):
// conv_in, UNetMidBlock2D (resnet, single-head attention, resnet), 4 decoder blocks of 3 ResnetBlock2D each
// (resnet.py:458-499; grouped 3x3 convs, GroupNorm(8)+SiLU, 1x1 grouped shortcut), GroupNorm+SiLU, grouped conv_out.
// Backward is the data gradient only (torch.autograd.grad w.r.t. latents, pipeline_wfd.py:613).
#include "decoder.cuh"
#include "kernels.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace wfd {

static int g_debug_ref_gemm = 0;
void debug_set_ref_gemm(int on) { g_debug_ref_gemm = on; }
int debug_ref_gemm() { return g_debug_ref_gemm; }

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

// ------------------------------------------------------------------------------------------------ host 16-bit
static uint16_t host_f_to_16(float f, bool bf16) {
    if (bf16) {
        uint32_t u;
        memcpy(&u, &f, 4);
        if ((u & 0x7FFFFFFFu) > 0x7F800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
        u += 0x7FFFu + ((u >> 16) & 1u);
        return static_cast<uint16_t>(u >> 16);
    }
    __half h = __float2half_rn(f);
    uint16_t r;
    memcpy(&r, &h, 2);
    return r;
}

// ------------------------------------------------------------------------------------------------ weight packing
static int choose_bn(int rpg) {
    if (rpg % 16 == 0) {
        if (rpg <= 256) {
            if (rpg >= 48) return rpg;
            return rpg == 16 ? 48 : 64;  // 16 -> 3 groups, 32 -> 2 groups
        }
        for (int d = 256; d >= 16; d -= 16)
            if (rpg % d == 0) return d;
    }
    if ((2 * rpg) % 16 == 0 && 2 * rpg <= 256) return 2 * rpg;
    return 64;
}

// Few-tile launches (batch 1, or one row band of a large map): a 128-pixel x BN tiling that yields fewer tiles than
// there are SMs leaves most of the GPU idle, so BN is halved (down to 48..96) until the launch has ~100 tiles.
static int fit_bn(int BN, long long m_tiles, int N) {
    while (BN >= 96 && (BN / 2) % 16 == 0 && N % (BN / 2) == 0 && m_tiles * ((N + BN - 1) / BN) < 100) BN /= 2;
    return BN;
}

void PackedB::release() {
    if (Bm) cudaFree(Bm);
    if (a_c0) cudaFree(a_c0);
    Bm = nullptr;
    a_c0 = nullptr;
}

// w: [cout][cin/groups][k][k] fp32 (torch Conv2d / Linear layout). mode 0: forward, pad (k-1)/2; mode 1: backward-data of
// mode 0; mode 2: forward of the stride-2 downsample conv applied after F.pad(x, (0,1,0,1)) (resnet.py:185-190).
int pack_conv(const float* w, int cout, int cin, int groups, int k, int mode, bool bf16, int bn_override, PackedB* out,
              const std::vector<TapSel>* tap_sel, int reuse) {
    const bool bwd = (mode == 1);
    const int cin_g = cin / groups, cout_g = cout / groups;
    const int N = bwd ? cin : cout;
    const int rpg = bwd ? cin_g : cout_g;   // output rows per group
    const int kpg = bwd ? cout_g : cin_g;   // contraction channels per group
    const int BN = bn_override > 0 ? bn_override : choose_bn(rpg);
    const int n_tiles = (N + BN - 1) / BN;
    std::vector<int> c0(n_tiles);
    int cpt = 1, span_max = 1;
    for (int t = 0; t < n_tiles; ++t) {
        const int n0 = t * BN, n1 = std::min(N, n0 + BN);
        const int g_lo = n0 / rpg, g_hi = (n1 - 1) / rpg;
        c0[t] = (g_lo * kpg) & ~7;  // TMA box start must be 16-byte aligned
        const int span = (g_hi + 1) * kpg - c0[t];
        span_max = std::max(span_max, span);
        cpt = std::max(cpt, (span + TG_KC - 1) / TG_KC);
    }
    // useful 16-channel steps in the last chunk of a tap: the rest only meets zero-padded weights
    const int kk_last = (span_max - (cpt - 1) * TG_KC + 15) / 16;
    const int wtaps = k * k;                                          // taps stored in the weight tensor
    const int taps = tap_sel ? static_cast<int>(tap_sel->size()) : wtaps;  // taps of this contraction
    const long long ldb = static_cast<long long>(taps) * cpt * TG_KC;
    const long long rows = static_cast<long long>(n_tiles) * BN;
    // tap offsets
    int tdx[TG_MAX_TAPS] = {0}, tdy[TG_MAX_TAPS] = {0};
    if (tap_sel) {
        for (int t = 0; t < taps; ++t) {
            tdx[t] = (*tap_sel)[t].dx;
            tdy[t] = (*tap_sel)[t].dy;
        }
    } else {
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) {
                const int tap = ky * k + kx;
                int dy = ky - (k - 1) / 2, dx = kx - (k - 1) / 2;
                if (mode == 2) { dy = ky; dx = kx; }
                if (bwd) { dy = -dy; dx = -dx; }
                tdx[tap] = dx;
                tdy[tap] = dy;
            }
    }
    if (reuse && (taps != 9 || mode == 2 || k != 3)) {
        set_error("pack_conv: the row-reuse chunk order needs a 3x3 stride-1 tap set");
        return -3;
    }
    std::vector<uint16_t> hb(static_cast<size_t>(rows * ldb), 0);
    for (int n = 0; n < N; ++n) {
        const int g = n / rpg, t = n / BN;
        for (int tap = 0; tap < taps; ++tap) {
            const int wt = tap_sel ? (*tap_sel)[tap].w_tap : tap;
            for (int kk = 0; kk < kpg; ++kk) {
                const int c = g * kpg + kk;
                const int pos = c - c0[t];
                const int cc = pos / TG_KC, within = pos % TG_KC;
                // chunk order: plain = (tap, cc); row reuse = (dx, cc, dy); patch = (cc, dy, dx), see b_chunk_index()
                const long long chunk = reuse == 2 ? static_cast<long long>(cc) * 9 + (tdy[tap] + 1) * 3 + (tdx[tap] + 1)
                                        : reuse ? (static_cast<long long>(tdx[tap] + 1) * cpt + cc) * 3 + (tdy[tap] + 1)
                                                : static_cast<long long>(tap) * cpt + cc;
                float v;
                if (!bwd) {
                    v = w[(static_cast<long long>(n) * cin_g + kk) * wtaps + wt];
                } else {
                    const int co = g * cout_g + kk, ci_local = n - g * cin_g;
                    v = w[(static_cast<long long>(co) * cin_g + ci_local) * wtaps + wt];
                }
                hb[static_cast<size_t>(n * ldb + chunk * TG_KC + within)] = host_f_to_16(v, bf16);
            }
        }
    }
    out->release();
    out->BN = BN;
    out->n_tiles = n_tiles;
    out->cpt = cpt;
    out->n_taps = taps;
    out->N = N;
    out->ldb = ldb;
    out->rows = static_cast<int>(rows);
    out->in_stride = (mode == 2) ? 2 : 1;
    out->reuse = reuse;
    out->kk_last = kk_last;
    for (int t = 0; t < taps; ++t) {
        out->dx[t] = tdx[t];
        out->dy[t] = tdy[t];
    }
    CUDA_CK(cudaMalloc(&out->Bm, hb.size() * 2));
    CUDA_CK(cudaMemcpy(out->Bm, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CUDA_CK(cudaMalloc(&out->a_c0, n_tiles * sizeof(int)));
    CUDA_CK(cudaMemcpy(out->a_c0, c0.data(), n_tiles * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

int upload_f32(const std::vector<float>& v, float** dev) {
    if (*dev) cudaFree(*dev);
    *dev = nullptr;
    CUDA_CK(cudaMalloc(dev, std::max<size_t>(v.size(), 1) * sizeof(float)));
    CUDA_CK(cudaMemcpy(*dev, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// ------------------------------------------------------------------------------------------------ plan construction
bool pick_box(int W, int H, int* w_box, int* h_box) {
    for (int wb = 128; wb >= 1; wb >>= 1) {
        if (W % wb != 0) continue;
        const int hb = TG_BM / wb;
        if (H % hb != 0) continue;
        *w_box = wb;
        *h_box = hb;
        return true;
    }
    return false;
}

Decoder::~Decoder() {
    for (auto& kv : convs) {
        kv.second.fwd.release();
        kv.second.bwd.release();
        if (kv.second.bias) cudaFree(kv.second.bias);
    }
    for (auto& kv : vecs)
        if (kv.second) cudaFree(kv.second);
}

int Decoder::init(const wfd_decoder_config& c, int h_, int w_, int Bmax, int dtype16, int band_rank, int band_nranks) {
    cfg = c;
    h = h_;
    w = w_;
    if (band_rank >= 0) {
        // row-band plan: this handle decodes rows [band_rank * h/R, (band_rank + 1) * h/R) of one h x w latent plane
        if (band_nranks < 1 || band_rank >= band_nranks || h_ % band_nranks != 0 || Bmax != 1) {
            set_error("decoder: row bands need batch 1 and a plane height (%d) divisible by the rank count (%d)", h_,
                      band_nranks);
            return -3;
        }
        for (int i = 0; i < c.n_blocks; ++i)
            if (c.block_down[i]) {
                set_error("decoder: row bands do not support DownDecoderBlock2D configs");
                return -3;
            }
        band = true;
        band_R = band_nranks;
        band_r = band_rank;
        h_total = h_;
        h = h_ / band_nranks;
    } else {
        h_total = h_;
    }
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
        const char* fa = getenv("WFD_NO_FUSE_ATTN_BWD");
        fuse_attn_bwd = (fa && fa[0] == '1') ? 0 : 1;
        const char* tl = getenv("WFD_NO_TC_LATENT");
        tc_latent = (tl && tl[0] == '1') ? 0 : 1;
        const char* ng = getenv("WFD_CONV_NGROUP");
        conv_n_group = ng ? atoi(ng) : 0;
        const char* m = getenv("WFD_FUSE_BWD_MAX_C");
        fuse_bwd_max_c = m ? atoi(m) : 768;
        const char* mm = getenv("WFD_NO_MMAJOR");
        mmajor_a = !(mm && mm[0] == '1');
        const char* rh = getenv("WFD_REUSE_HBOX");
        reuse_hbox = rh ? atoi(rh) : 0;  // 0: keep the widest patch pick_box() finds
    }
    if (c.latent_channels != 4) {
        set_error("decoder: latent_channels must be 4 (got %d)", c.latent_channels);
        return -3;
    }
    if (c.n_blocks < 1 || c.n_blocks > 8 || Bmax < 1 || h < 1 || w < 1) {
        set_error("decoder: bad config (n_blocks=%d B_max=%d h=%d w=%d)", c.n_blocks, Bmax, h, w);
        return -3;
    }
    const int nb = c.n_blocks;
    Cm = c.block_out_channels[nb - 1];
    // resnet list in execution order: mid.0, [attention], mid.1, then blocks
    resnets.clear();
    auto add_res = [&](const std::string& prefix, int cin, int cout, int groups, int hh, int ww) {
        ResnetDesc r;
        r.prefix = prefix;
        r.cin = cin;
        r.cout = cout;
        r.groups = groups;
        r.h = hh;
        r.w = ww;
        resnets.push_back(r);
    };
    add_res("decoder.mid_block.resnets.0", Cm, Cm, 1, h, w);
    add_res("decoder.mid_block.resnets.1", Cm, Cm, 1, h, w);
    int ch = Cm, hh = h, ww = w;
    down_after.assign(nb, 0);
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[nb - 1 - i];
        const int groups = c.conv_groups[nb - 1 - i];
        for (int j = 0; j < c.layers_per_block + 1; ++j) {
            char buf[128];
            snprintf(buf, sizeof(buf), "decoder.up_blocks.%d.resnets.%d", i, j);
            add_res(buf, j == 0 ? ch : cout, cout, groups, hh, ww);
        }
        ch = cout;
        if (c.block_down[i]) {
            down_after[i] = 1;
            if (hh % 2 || ww % 2) {
                set_error("decoder: DownDecoderBlock2D needs an even plane (%dx%d)", hh, ww);
                return -3;
            }
            hh /= 2;
            ww /= 2;
        }
    }
    C_last = ch;
    H_out = hh;
    W_out = ww;
    T_pad = c.out_channels;
    for (const ResnetDesc& r : resnets) {
        if (r.cin % 8 || r.cout % 16 || (r.cin / G) % 8 || (r.cout / G) % 8 || r.cin % r.groups || r.cout % r.groups) {
            set_error("decoder: unsupported channel counts in %s (%d -> %d, groups %d)", r.prefix.c_str(), r.cin, r.cout,
                      r.groups);
            return -3;
        }
    }
    if (T_pad % 16 || T_pad % c.conv_init_groups || C_last % c.conv_init_groups) {
        set_error("decoder: out_channels=%d must be a multiple of 16 and of conv_init_groups", T_pad);
        return -3;
    }
    int wb, hb;
    if (!pick_box(w, h, &wb, &hb) || !pick_box(W_out, H_out, &wb, &hb) || (static_cast<long long>(h) * w) % 128) {
        set_error("decoder: plane %dx%d cannot be tiled into 128-pixel patches", h, w);
        return -3;
    }
    layout_workspace();
    return 0;
}

// Workspace layout (all offsets 256-byte aligned); sizes for B_max.
void Decoder::layout_workspace() {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += ((bytes + 255) & ~size_t(255)) + halo_pad;
        return o;
    };
    const size_t e = 2;  // bytes per 16-bit element
    const size_t B = static_cast<size_t>(B_max);
    const size_t hw = static_cast<size_t>(h) * w;
    int maxC = Cm;
    for (const ResnetDesc& r : resnets) maxC = std::max(maxC, std::max(r.cin, r.cout));
    if (band) {
        // every buffer keeps one row of its widest tenant free on both sides: the halo rows of the 3x3 contractions
        halo_pad = (static_cast<size_t>(w) * std::max(std::max(maxC, T_pad), 3 * Cm) * e + 255) & ~size_t(255);
        off = halo_pad;
    }
    o_x0 = take(B * hw * Cm * e);
    o_res_h1.clear();
    o_res_out.clear();
    for (const ResnetDesc& r : resnets) {
        const size_t phw = static_cast<size_t>(r.h) * r.w;
        o_res_h1.push_back(take(B * phw * r.cout * e));
        o_res_out.push_back(take(B * phw * r.cout * e));
    }
    o_down_out.assign(cfg.n_blocks, 0);
    {
        int hh = h, ww = w;
        for (int i = 0; i < cfg.n_blocks; ++i) {
            if (down_after[i]) {
                hh /= 2;
                ww /= 2;
                o_down_out[i] = take(B * hh * ww * cfg.block_out_channels[cfg.n_blocks - 1 - i] * e);
            }
        }
    }
    // attention
    // N queries (this band's pixels) attend over Nk keys (the whole plane)
    const size_t N = hw, Nk = hw * band_R;
    o_attn_out = take(B * N * Cm * e);
    o_xn = take(B * N * Cm * e);
    o_qkv = take(B * Nk * 3 * Cm * e);
    o_vt = take(B * Cm * Nk * e);
    {
        // fp32 N x Nk scratch of the unfused attention (scores, then dP). With both softmax directions fused into the
        // products' epilogues only the row-statistics partials (16 bytes per row and N tile) and the 48-column product of
        // conv_in^T live here: 17 GB less at 256 x 256
        const size_t bn_n = Nk % 256 == 0 ? 256 : 128;
        const size_t fused = std::max(B * N * (Nk / bn_n) * 16, B * hw * 48 * 4);
        o_S = take((fuse_attn_fwd && fuse_attn_bwd) ? fused : std::max(fused, B * N * Nk * 4));
    }
    o_RS = take(B * N * 8);
    o_P = take(B * N * Nk * e);
    o_O = take(B * N * Cm * e);
    // backward-only attention scratch
    o_PT = mmajor_a ? 0 : take(B * N * Nk * e);  // transposed copies only without the M-major A operand
    o_dS = take(B * N * Nk * e);
    o_dST = mmajor_a ? 0 : take(B * N * Nk * e);
    o_dO = take(B * N * Cm * e);
    o_D = take(B * N * 4);
    o_dOT = take(B * Cm * N * e);
    o_KT = take(B * Cm * Nk * e);
    o_QT = take(B * Cm * N * e);
    o_dqkv = take(B * N * 3 * Cm * e);
    o_dkv = band_R > 1 ? take(Nk * 2 * Cm * 4) : 0;
    // scratch
    o_act = take(B * hw * maxC * e);
    o_act2 = take(B * hw * maxC * e);
    o_sc = take(B * hw * maxC * e);
    for (int i = 0; i < 4; ++i) o_g[i] = take(B * hw * maxC * e);
    const size_t ohw = static_cast<size_t>(H_out) * W_out;
    o_logits = take(B * ohw * T_pad * e);
    o_dlogits = take(B * ohw * T_pad * e);
    n_gn = static_cast<int>(resnets.size()) * 2 + 2;  // + attention group_norm + conv_norm_out
    o_gn = take(static_cast<size_t>(n_gn) * 2 * B * G * 2 * sizeof(acc_t));  // sums then reds (fixed point)
    ws_bytes = off;
}

int Decoder::set_weight(const char* key, const float* data, long long numel) {
    std::string k(key);
    if (encoder_side ? (k.rfind("encoder.", 0) != 0 && k.rfind("quant_conv.", 0) != 0)
                     : (k.rfind("decoder.", 0) != 0 && k.rfind("post_quant_conv.", 0) != 0))
        return 1;
    host[k] = std::vector<float>(data, data + numel);
    finalized = false;
    return 0;
}

int Decoder::need(const std::string& key, long long numel, const std::vector<float>** out) {
    auto it = host.find(key);
    if (it == host.end()) {
        set_error("decoder: missing weight '%s'", key.c_str());
        return -9;
    }
    if (static_cast<long long>(it->second.size()) != numel) {
        set_error("decoder: weight '%s' has %lld elements, expected %lld", key.c_str(),
                  static_cast<long long>(it->second.size()), numel);
        return -9;
    }
    *out = &it->second;
    return 0;
}

int Decoder::pack_named_conv(const std::string& name, int cout, int cin, int groups, int k, bool want_bwd, int fwd_mode,
                             int ph, int pw) {
    const std::vector<float>*wv, *bv;
    CK(need(name + ".weight", static_cast<long long>(cout) * (cin / groups) * k * k, &wv));
    CK(need(name + ".bias", cout, &bv));
    ConvW& cw = convs[name];
    cw.cin = cin;
    cw.cout = cout;
    // the row-reuse schedule pays off when a stage (patch with halo rows + three weight boxes) leaves >= 3 stages
    // K-loop schedule of a 3x3 stride-1 conv: the patch schedule whenever the plane tiles into 16 x 8 pixel patches (one
    // activation box per 64-channel chunk serves all nine taps); else row reuse when a stage leaves room for 3 stages
    auto want_reuse = [&](int rows_per_group) -> int {
        if (k != 3 || fwd_mode != 0 || ph <= 0) return 0;
        if (use_patch && pw % 8 == 0 && ph % 16 == 0) return 2;
        if (!use_reuse) return 0;
        int wb = 0, hb = 0;
        if (!pick_box(pw, ph, &wb, &hb) || wb < 8) return 0;
        const int BN = choose_bn(rows_per_group);
        const size_t per_stage = static_cast<size_t>(hb + 2) * wb * 128 + 3 * static_cast<size_t>(BN) * 128;
        return per_stage * 3 <= 216 * 1024 ? 1 : 0;
    };
    // dense (groups == 1) layers may use a narrower N tile when the plan has few 128-pixel tiles
    const long long m_tiles = ph > 0 ? static_cast<long long>(B_max) * ((static_cast<long long>(ph) * pw) / TG_BM)
                                     : static_cast<long long>(B_max) * ((static_cast<long long>(h) * w) / TG_BM);
    const int bn_f = groups == 1 ? fit_bn(choose_bn(cout), m_tiles, cout) : 0;
    const int bn_b = groups == 1 ? fit_bn(choose_bn(cin), m_tiles, cin) : 0;
    CK(pack_conv(wv->data(), cout, cin, groups, k, fwd_mode, bf16 != 0, bn_f, &cw.fwd, nullptr, want_reuse(cout / groups)));
    if (want_bwd && fwd_mode == 0)
        CK(pack_conv(wv->data(), cout, cin, groups, k, 1, bf16 != 0, bn_b, &cw.bwd, nullptr, want_reuse(cin / groups)));
    CK(upload_f32(*bv, &cw.bias));
    return 0;
}

int Decoder::upload_vec(const std::string& key, long long numel) {
    const std::vector<float>* v;
    CK(need(key, numel, &v));
    float*& dev = vecs[key];
    CK(upload_f32(*v, &dev));
    return 0;
}

// attention weights: q, k, v merged into one [3C, C] matrix (+ backward form when want_bwd), proj
int Decoder::pack_attention(const std::string& a, bool want_bwd) {
    CK(upload_vec(a + "group_norm.weight", Cm));
    CK(upload_vec(a + "group_norm.bias", Cm));
    const std::vector<float>*q, *k, *v, *qb, *kb, *vb;
    const long long cc = static_cast<long long>(Cm) * Cm;
    CK(need(a + "query.weight", cc, &q));
    CK(need(a + "key.weight", cc, &k));
    CK(need(a + "value.weight", cc, &v));
    CK(need(a + "query.bias", Cm, &qb));
    CK(need(a + "key.bias", Cm, &kb));
    CK(need(a + "value.bias", Cm, &vb));
    std::vector<float> wq(3 * cc), bq(3 * Cm);
    memcpy(wq.data(), q->data(), cc * 4);
    memcpy(wq.data() + cc, k->data(), cc * 4);
    memcpy(wq.data() + 2 * cc, v->data(), cc * 4);
    memcpy(bq.data(), qb->data(), Cm * 4);
    memcpy(bq.data() + Cm, kb->data(), Cm * 4);
    memcpy(bq.data() + 2 * Cm, vb->data(), Cm * 4);
    ConvW& cq = convs[a + "qkv"];
    cq.cin = Cm;
    cq.cout = 3 * Cm;
    CK(pack_conv(wq.data(), 3 * Cm, Cm, 1, 1, 0, bf16 != 0, Cm % 192 == 0 ? 192 : 0, &cq.fwd, nullptr, 0));
    if (want_bwd) CK(pack_conv(wq.data(), 3 * Cm, Cm, 1, 1, 1, bf16 != 0, 0, &cq.bwd, nullptr, 0));
    CK(upload_f32(bq, &cq.bias));
    CK(pack_named_conv(a + "proj_attn", Cm, Cm, 1, 1, want_bwd, 0, 0, 0));
    return 0;
}

int Decoder::finalize_weights() {
    CK(upload_vec("post_quant_conv.weight", 16));
    CK(upload_vec("post_quant_conv.bias", 4));
    CK(upload_vec("decoder.conv_in.weight", static_cast<long long>(Cm) * 36));
    CK(upload_vec("decoder.conv_in.bias", Cm));
    {   // [co][ci][ky][kx] -> [ky][kx][ci][co] for the coalesced reads of latent_out
        const std::vector<float>* wv;
        CK(need("decoder.conv_in.weight", static_cast<long long>(Cm) * 36, &wv));
        std::vector<float> wt(static_cast<size_t>(Cm) * 36);
        for (int co = 0; co < Cm; ++co)
            for (int ci = 0; ci < 4; ++ci)
                for (int t = 0; t < 9; ++t) wt[(static_cast<size_t>(t) * 4 + ci) * Cm + co] = (*wv)[(co * 4 + ci) * 9 + t];
        CK(upload_f32(wt, &vecs["decoder.conv_in.weight.T"]));
    }
    if (!encoder_side && Cm % 64 == 0) {
        // conv_in^T as a tensor-core product: row tap*4 + ci of a 48-row matrix holds w[co, ci, tap] over co (rows 36..47
        // are zero), i.e. a 1x1 "conv" from the Cm gradient channels to the 36 (tap, latent channel) pairs
        const std::vector<float>* wv;
        CK(need("decoder.conv_in.weight", static_cast<long long>(Cm) * 36, &wv));
        std::vector<float> wt(static_cast<size_t>(48) * Cm, 0.f);
        for (int co = 0; co < Cm; ++co)
            for (int ci = 0; ci < 4; ++ci)
                for (int t = 0; t < 9; ++t) wt[static_cast<size_t>(t * 4 + ci) * Cm + co] = (*wv)[(co * 4 + ci) * 9 + t];
        ConvW& cw = convs["decoder.conv_in.T36"];
        cw.cin = Cm;
        cw.cout = 48;
        CK(pack_conv(wt.data(), 48, Cm, 1, 1, 0, bf16 != 0, 48, &cw.fwd, nullptr, 0));
    }
    for (const ResnetDesc& r : resnets) {
        CK(upload_vec(r.prefix + ".norm1.weight", r.cin));
        CK(upload_vec(r.prefix + ".norm1.bias", r.cin));
        CK(upload_vec(r.prefix + ".norm2.weight", r.cout));
        CK(upload_vec(r.prefix + ".norm2.bias", r.cout));
        CK(pack_named_conv(r.prefix + ".conv1", r.cout, r.cin, r.groups, 3, true, 0, r.h, r.w));
        CK(pack_named_conv(r.prefix + ".conv2", r.cout, r.cout, r.groups, 3, true, 0, r.h, r.w));
        if (r.cin != r.cout) CK(pack_named_conv(r.prefix + ".conv_shortcut", r.cout, r.cin, r.groups, 1, true, 0, 0, 0));
    }
    for (int i = 0; i < cfg.n_blocks; ++i) {
        if (!down_after[i]) continue;
        char buf[128];
        snprintf(buf, sizeof(buf), "decoder.up_blocks.%d.downsamplers.0.conv", i);
        const int cch = cfg.block_out_channels[cfg.n_blocks - 1 - i];
        const int grp = cfg.conv_groups[cfg.n_blocks - 1 - i];
        CK(pack_named_conv(buf, cch, cch, grp, 3, false, 2, 0, 0));
        // backward-data of the stride-2 conv: one contraction per output parity (py, px). For row parity 0 the taps
        // ky = 0 (source row y') and ky = 2 (source row y' - 1) contribute, for parity 1 only ky = 1; same along x.
        const std::vector<float>* wv;
        CK(need(std::string(buf) + ".weight", static_cast<long long>(cch) * (cch / grp) * 9, &wv));
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                std::vector<TapSel> sel;
                for (int ky = py; ky < 3; ky += 2)
                    for (int kx = px; kx < 3; kx += 2) sel.push_back(TapSel{ky * 3 + kx, -((kx - px) / 2), -((ky - py) / 2)});
                ConvW& cw = convs[std::string(buf) + ".bwd" + std::to_string(py * 2 + px)];
                cw.cin = cch;
                cw.cout = cch;
                CK(pack_conv(wv->data(), cch, cch, grp, 3, 1, bf16 != 0, 0, &cw.bwd, &sel, 0));
            }
    }
    CK(pack_attention("decoder.mid_block.attentions.0.", true));
    CK(upload_vec("decoder.conv_norm_out.weight", C_last));
    CK(upload_vec("decoder.conv_norm_out.bias", C_last));
    CK(pack_named_conv("decoder.conv_out", T_pad, C_last, cfg.conv_init_groups, 3, true, 0, H_out, W_out));
    finalized = true;
    host.clear();
    if (ws) return prepare_ops();
    return 0;
}

int Decoder::bind_workspace(void* p, size_t bytes) {
    if (bytes < ws_bytes) {
        set_error("decoder: workspace too small (%zu < %zu)", bytes, ws_bytes);
        return -3;
    }
    if (reinterpret_cast<uintptr_t>(p) & 255) {
        set_error("decoder: workspace must be 256-byte aligned");
        return -3;
    }
    ws = static_cast<uint8_t*>(p);
    if (finalized) return prepare_ops();
    return 0;
}

int Decoder::prepare_ops() {
    fwd_ops.clear();
    bwd_ops.clear();
    Exec ex;
    ex.prepare = true;
    ex.B = B_max;
    ex.stream = 0;
    ex.ops = &fwd_ops;
    CK(forward_impl(ex, nullptr, 0, 1.f));
    ex.ops = &bwd_ops;
    ex.cursor = 0;
    CK(backward_impl(ex, nullptr, nullptr, nullptr));
    prepared = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ gemm helpers
int Decoder::run_gemm(Exec& ex, TapGemmOp& tmpl, bool z_is_batch) {
    if (ex.prepare) {
        CK(tapgemm_prepare(&tmpl));
        snprintf(tmpl.tag, sizeof(tmpl.tag), "%s", cur_tag);
        ex.ops->push_back(tmpl);
        return 0;
    }
    if (ex.cursor >= ex.ops->size()) {
        set_error("decoder: internal op cursor overrun");
        return -10;
    }
    TapGemmOp op = (*ex.ops)[ex.cursor++];
    if (z_is_batch) op.p.z_count = ex.B;
    else op.p.m_tiles = ex.B * op.p.tiles_x * op.p.tiles_y;
    return g_debug_ref_gemm ? tapgemm_launch_ref(&op, ex.stream) : tapgemm_launch(&op, ex.stream);
}

// conv-style contraction over a channels-last activation
int Decoder::conv_gemm(Exec& ex, const PackedB& pk, const void* A, int a_C, int Hin, int Win, int Hout, int Wout,
                       void* out, long long ldc, const float* bias, const void* residual, acc_t* gn_sums, int gn_cpg,
                       const OutStrides* os, const GnBwdFuse* gb, const AxForm* ax) {
    TapGemmOp op;
    memset(&op, 0, sizeof(op));
    TapGemmParams& p = op.p;
    int wb = 0, hb = 0;
    bool vertical = false;  // any tap that leaves the output row: its input needs the neighbour bands' edge rows
    for (int t = 0; t < pk.n_taps; ++t) vertical |= pk.dy[t] != 0;
    if (band && vertical) CK(halo_rows(ex, const_cast<void*>(A), Hin, Win, a_C));
    if (ex.prepare) {
        if (!pick_box(Wout, Hout, &wb, &hb)) {
            set_error("decoder: cannot tile %dx%d plane", Hout, Wout);
            return -3;
        }
        // Row-reuse contractions fetch (h_box + 2) image rows per horizontal shift: a taller, narrower patch re-reads
        // fewer halo rows from L2 (3 x 2.0 of the activation bytes at 64 x 2, 3 x 1.25 at 16 x 8). Measured on B200
        // (WFD_REUSE_HBOX = 2 / 4 / 8: 1.00 / 1.14 / 1.12 ms for the five 3072-channel groups=64 launches): no gain, so
        // L2 delivery is not what bounds them and the default stays the widest patch.
        if (pk.reuse == 2) {  // patch schedule: 16 rows x 8 pixels
            wb = 8;
            hb = 16;
        } else if (pk.reuse && reuse_hbox > hb) {
            const int hb2 = reuse_hbox, wb2 = TG_BM / reuse_hbox;
            if (wb2 >= 8 && wb2 * hb2 == TG_BM && Wout % wb2 == 0 && Hout % hb2 == 0) {
                wb = wb2;
                hb = hb2;
            }
        }
        p.BN = pk.BN;
        p.n_tiles = pk.n_tiles;
        p.z_count = 1;
        // N fastest: the CTAs of a wave work on the same pixels across channel groups, so activations are read and
        // written as whole channels-last pixel rows (DRAM page locality) and dense layers fetch each A tile once
        p.n_group = conv_n_group > 1 ? std::min(conv_n_group, pk.n_tiles) : 0;  // default: M fastest (measured best)
        p.w_box = wb;
        p.h_box = hb;
        p.tiles_x = Wout / wb;
        p.tiles_y = Hout / hb;
        p.m_tiles = B_max * p.tiles_x * p.tiles_y;
        p.W = Wout;
        p.H = Hout;
        p.in_stride = pk.in_stride;
        p.n_taps = pk.n_taps;
        p.cpt = pk.cpt;
        for (int t = 0; t < pk.n_taps; ++t) {
            p.tap_dx[t] = pk.dx[t];
            p.tap_dy[t] = pk.dy[t];
        }
        p.a_c0 = pk.a_c0;
        p.reuse = pk.reuse;
        p.kk_last = pk.kk_last;
        p.N = pk.N;
        p.out = out;
        p.out_fp32 = next_out_fp32 ? 1 : 0;
        next_out_fp32 = false;
        p.out_sx = os ? os->sx : ldc;
        p.out_sy = os ? os->sy : ldc * Wout;
        p.out_sb = os ? os->sb : ldc * Wout * Hout;
        p.bias = bias;
        p.residual = residual;
        p.alpha = 1.f;
        p.gn_sums = (fuse_gn && gn_cpg > 0 && gn_cpg % 16 == 0) ? gn_sums : nullptr;
        p.gn_cpg = gn_cpg;
        p.gn_G = G;
        if (gb && gb->red) {
            p.gnb_red = gb->red;
            p.gnb_x = gb->x;
            p.gnb_stats = gb->stats;
            p.gnb_gamma = gb->gamma;
            p.gnb_beta = gb->beta;
            p.gnb_cpg = gb->cpg;
            p.gnb_silu = gb->silu;
            p.gnb_inv_n = gb->inv_n;
            p.gnb_eps = gb->eps;
        }
        p.is_bf16 = bf16;
        p.a_ptr = A;
        p.a_C = a_C;
        p.a_W = Win;
        p.a_H = Hin;
        if (band && vertical) {  // the view starts at the top halo row
            p.a_ptr = static_cast<const uint8_t*>(A) - static_cast<size_t>(Win) * a_C * 2;
            p.a_H = Hin + 2;
            p.a_y_off = 1;
        }
        if (ax) {
            p.ax_stats = ax->stats;
            p.ax_gamma = ax->gamma;
            p.ax_beta = ax->beta;
            p.ax_cpg = ax->cpg;
            p.ax_silu = ax->silu;
            p.ax_hw = ax->hw;
            p.ax_eps = ax->eps;
            // image rows of the A view: everything but the zeroed halo rows at the top / bottom border of the map
            p.ax_y0 = (band && vertical && band_r == 0) ? 1 : 0;
            p.ax_y1 = p.a_H - ((band && vertical && band_r == band_R - 1) ? 1 : 0);
        }
        p.a_B = B_max;
        p.a_ld = a_C;
        p.b_ptr = pk.Bm;
        p.ldb = pk.ldb;
        p.b_zstride = 0;
        p.b_rows = pk.rows;
        p.b_Z = 1;
    }
    return run_gemm(ex, op, false);
}

// plain batched GEMM: D[rows, N] = alpha * A[rows, K] * Bm[N, K]^T (+bias)(+residual); rows are tokens of sample bb
struct PlainGemm {
    const void* A; int K; long long a_ld; int rows_per_sample; int a_B;   // A: [a_B, rows_per_sample, a_ld]
    bool a_mn_major;                                                      // A stored transposed: [a_B, K, a_ld >= rows]
    bool a_shared;                                                        // A identical for every z (weights as A)
    const void* Bm; int b_rows; long long ldb, b_zstride; int b_Z;        // Bm: [b_Z][b_rows][ldb]
    bool b_per_sample;                                                    // Bm slice = sample index
    int BN, N;
    void* out; int out_fp32; long long ldc, c_zstride;
    const float* bias; const void* residual; float alpha;
    acc_t* gn_sums; int gn_cpg;
    const GnBwdFuse* gb;
    const float* sb_d; float sb_scale;   // softmax-backward epilogue: out = residual (= P) * (acc - sb_d[row]) * sb_scale
    float2* rs_out; const float2* rs_in; // row-softmax passes (TapGemmParams::rs_out / rs_in)
};
int Decoder::plain_gemm(Exec& ex, const PlainGemm& g) {
    TapGemmOp op;
    memset(&op, 0, sizeof(op));
    if (ex.prepare) {
        TapGemmParams& p = op.p;
        if (g.rows_per_sample % TG_BM || g.K % TG_KC) {
            set_error("decoder: plain gemm needs rows %% 128 == 0 and K %% 64 == 0 (rows=%d K=%d)", g.rows_per_sample, g.K);
            return -3;
        }
        p.BN = g.BN;
        p.n_tiles = (g.N + g.BN - 1) / g.BN;
        p.w_box = TG_BM;
        p.h_box = 1;
        p.tiles_x = g.rows_per_sample / TG_BM;
        p.tiles_y = 1;
        p.W = g.rows_per_sample;
        p.H = 1;
        p.in_stride = 1;
        p.n_taps = 1;
        p.cpt = g.K / TG_KC;
        p.tap_dx[0] = p.tap_dy[0] = 0;
        p.a_c0 = nullptr;
        if (g.a_shared) {  // z enumerates samples, A is shared
            p.z_count = B_max;
            p.m_tiles = p.tiles_x;
            p.a_zmul = 0;
            p.b_zmul = 1;
            p.b_bbmul = 0;
        } else {
            p.z_count = 1;
            p.m_tiles = B_max * p.tiles_x;
            p.a_zmul = 0;
            p.b_zmul = 0;
            p.b_bbmul = g.b_per_sample ? 1 : 0;
        }
        p.N = g.N;
        p.out = g.out;
        p.out_fp32 = g.out_fp32;
        p.out_sx = g.ldc;
        p.out_sy = 0;
        p.out_sb = g.ldc * g.rows_per_sample;
        p.c_zstride = g.c_zstride;
        p.bias = g.bias;
        p.residual = g.residual;
        p.alpha = g.alpha;
        p.sb_d = g.sb_d;
        p.sb_scale = g.sb_scale;
        p.rs_out = g.rs_out;
        p.rs_in = g.rs_in;
        p.gn_sums = (fuse_gn && g.gn_cpg > 0 && g.gn_cpg % 16 == 0) ? g.gn_sums : nullptr;
        p.gn_cpg = g.gn_cpg;
        p.gn_G = G;
        if (g.gb && g.gb->red) {
            p.gnb_red = g.gb->red;
            p.gnb_x = g.gb->x;
            p.gnb_stats = g.gb->stats;
            p.gnb_gamma = g.gb->gamma;
            p.gnb_beta = g.gb->beta;
            p.gnb_cpg = g.gb->cpg;
            p.gnb_silu = g.gb->silu;
            p.gnb_inv_n = g.gb->inv_n;
            p.gnb_eps = g.gb->eps;
        }
        p.is_bf16 = bf16;
        p.a_mn_major = g.a_mn_major ? 1 : 0;
        p.a_ptr = g.A;
        p.a_C = g.K;
        p.a_W = g.rows_per_sample;
        p.a_H = 1;
        p.a_B = g.a_B;
        p.a_ld = g.a_ld;
        p.b_ptr = g.Bm;
        p.ldb = g.ldb;
        p.b_zstride = g.b_zstride;
        p.b_rows = g.b_rows;
        p.b_Z = g.b_Z;
    }
    return run_gemm(ex, op, g.a_shared);
}

// ---- row-band exchanges (no-ops outside band plans)
int Decoder::sync_sums(Exec& ex, acc_t* slot) {
    if (!band || ex.prepare || band_R == 1) return 0;
    if (!comm) {
        set_error("decoder: row-band plan without an attached communicator");
        return -12;
    }
    return comm->allreduce_i64(slot, static_cast<size_t>(G) * 2, ex.stream);
}
int Decoder::halo_rows(Exec& ex, void* buf, int rows, int W, int C) {
    if (!band || ex.prepare) return 0;
    const size_t rb = static_cast<size_t>(W) * C * 2;
    uint8_t* b = static_cast<uint8_t*>(buf);
    uint8_t* top = b - rb;
    uint8_t* bottom = b + static_cast<size_t>(rows) * rb;
    if (band_R == 1) {
        CUDA_CK(cudaMemsetAsync(top, 0, rb, ex.stream));
        CUDA_CK(cudaMemsetAsync(bottom, 0, rb, ex.stream));
        return 0;
    }
    if (!comm) {
        set_error("decoder: row-band plan without an attached communicator");
        return -12;
    }
    return comm->halo(top, b, b + static_cast<size_t>(rows - 1) * rb, bottom, rb, ex.stream);
}
// GroupNorm(+SiLU) backward; in a band plan the two reduction sums are completed over the bands before the apply pass
int Decoder::gn_backward(Exec& ex, const void* dy, const void* x, int gn_idx, const float* gamma, const float* beta,
                         const void* add, void* dx, int HW, int C, int silu, int do_reduce) {
    const float eps = 1e-6f;
    if (!band)
        return gn_bwd(dy, x, gn_sums_ptr(gn_idx), gamma, beta, gn_red_ptr(gn_idx), add, dx, ex.B, HW, C, G, eps, silu, bf16,
                      do_reduce, ex.stream);
    const int HW_stat = HW * band_R;
    if (do_reduce)
        CK(gn_bwd_reduce(dy, x, gn_sums_ptr(gn_idx), gamma, beta, gn_red_ptr(gn_idx), ex.B, HW, C, G, eps, silu, bf16,
                         ex.stream, HW_stat));
    CK(sync_sums(ex, gn_red_ptr(gn_idx)));
    return gn_bwd_apply(dy, x, gn_sums_ptr(gn_idx), gamma, beta, gn_red_ptr(gn_idx), add, dx, ex.B, HW, C, G, eps, silu, bf16,
                        ex.stream, HW_stat);
}

acc_t* Decoder::gn_sums_ptr(int idx) const {
    return reinterpret_cast<acc_t*>(ws + o_gn) + static_cast<size_t>(idx) * B_max * G * 2;
}
acc_t* Decoder::gn_red_ptr(int idx) const {
    return reinterpret_cast<acc_t*>(ws + o_gn) + static_cast<size_t>(n_gn + idx) * B_max * G * 2;
}

// ------------------------------------------------------------------------------------------------ shared forward blocks
static const char* group_tag_of(int groups) {
    return groups == 1 ? "dense" : groups == 4 ? "g4" : groups == 16 ? "g16" : groups == 64 ? "g64" : "grouped";
}

// ResnetBlock2D.forward with temb = None (reference resnet.py:458-499). next_gn: index of the GroupNorm that consumes the
// output (its statistics ride in conv2's epilogue), or -1 when the consumer computes them itself.
int Decoder::resnet_forward(Exec& ex, int idx, uint8_t* xin, bool stats_ready, int next_gn) {
    const int B = ex.B;
    cudaStream_t s = ex.stream;
    const float eps = 1e-6f;
    auto fused_ok = [&](int C) { return fuse_gn && (C / G) % 16 == 0; };
    auto tag_of = [&](int groups) { return group_tag_of(groups); };
    {
        const ResnetDesc& r = resnets[idx];
        cur_tag = tag_of(r.groups);
        const int phw = r.h * r.w;
        uint8_t* act = ws + o_act;
        uint8_t* act2 = ws + o_act2;
        uint8_t* h1 = ws + o_res_h1[idx];
        uint8_t* xout = ws + o_res_out[idx];
        const int gi1 = 2 * idx, gi2 = 2 * idx + 1;
        int next_cpg = r.cout / G;
        const ConvW& c1 = convs[r.prefix + ".conv1"];
        const ConvW& c2 = convs[r.prefix + ".conv2"];
        // GroupNorm + SiLU in front of a patch-schedule conv runs inside that conv (transform warps rewrite the landed
        // activation box in shared memory): the conv then reads the RAW tensor and the standalone gn_apply pass is gone
        const bool ax1 = apply_fused(c1.fwd), ax2 = apply_fused(c2.fwd);
        AxForm a1 = {gn_sums_ptr(gi1), vecs[r.prefix + ".norm1.weight"], vecs[r.prefix + ".norm1.bias"], r.cin / G, 1,
                     phw * band_R, eps};
        AxForm a2 = {gn_sums_ptr(gi2), vecs[r.prefix + ".norm2.weight"], vecs[r.prefix + ".norm2.bias"], r.cout / G, 1,
                     phw * band_R, eps};
        if (!ex.prepare) {
            if (!stats_ready || !fused_ok(r.cin)) CK(gn_stats(xin, gn_sums_ptr(gi1), B, phw, r.cin, G, bf16, s));
            CK(sync_sums(ex, gn_sums_ptr(gi1)));
            if (!ax1)
                CK(gn_apply(xin, a1.stats, a1.gamma, a1.beta, act, B, phw, r.cin, G, eps, 1, bf16, s, phw * band_R));
        }
        CK(conv_gemm(ex, c1.fwd, ax1 ? xin : act, r.cin, r.h, r.w, r.h, r.w, h1, r.cout, c1.bias, nullptr, gn_sums_ptr(gi2),
                     r.cout / G, nullptr, nullptr, ax1 ? &a1 : nullptr));
        if (!ex.prepare) {
            if (!fused_ok(r.cout)) CK(gn_stats(h1, gn_sums_ptr(gi2), B, phw, r.cout, G, bf16, s));
            CK(sync_sums(ex, gn_sums_ptr(gi2)));
            if (!ax2)
                CK(gn_apply(h1, a2.stats, a2.gamma, a2.beta, act2, B, phw, r.cout, G, eps, 1, bf16, s, phw * band_R));
        }
        const void* residual = xin;
        if (r.cin != r.cout) {
            const ConvW& cs = convs[r.prefix + ".conv_shortcut"];
            CK(conv_gemm(ex, cs.fwd, xin, r.cin, r.h, r.w, r.h, r.w, ws + o_sc, r.cout, cs.bias, nullptr, nullptr, 0, nullptr, nullptr));
            residual = ws + o_sc;
        }
        CK(conv_gemm(ex, c2.fwd, ax2 ? h1 : act2, r.cout, r.h, r.w, r.h, r.w, xout, r.cout, c2.bias, residual,
                     next_gn < 0 ? nullptr : gn_sums_ptr(next_gn), next_cpg, nullptr, nullptr, ax2 ? &a2 : nullptr));
        return 0;
    }
}

// AttentionBlock.forward (reference attention.py:329-380): one head of Cm channels over the h x w plane; x -> ws + o_attn_out.
// attn_prefix: state_dict prefix of the block ("decoder.mid_block.attentions.0."); next_gn as in resnet_forward.
int Decoder::attention_forward(Exec& ex, uint8_t* x, const std::string& attn_prefix, int next_gn) {
    const int B = ex.B;
    cudaStream_t s = ex.stream;
    const float eps = 1e-6f;
    const int hw = h * w;
    auto fused_ok = [&](int C) { return fuse_gn && (C / G) % 16 == 0; };
    {
        // AttentionBlock.forward (attention.py:329-380), one head of Cm channels. N queries (this band's pixels)
        // attend over Nk keys; in a band plan the q|k|v rows of all bands are gathered into one [Nk, 3*Cm] block
        const int N = hw, Nk = hw * band_R;
        cur_tag = "attn";
        uint8_t* xn = ws + o_xn;
        uint8_t* qkv_all = ws + o_qkv;
        uint8_t* qkv = qkv_all + static_cast<size_t>(band_r) * N * 3 * Cm * 2;  // this band's rows
        const std::string a = attn_prefix;
        if (!ex.prepare) {
            if (!fused_ok(Cm)) CK(gn_stats(x, gn_sums_ptr(n_gn - 2), B, N, Cm, G, bf16, s));
            CK(sync_sums(ex, gn_sums_ptr(n_gn - 2)));
            CK(gn_apply(x, gn_sums_ptr(n_gn - 2), vecs[a + "group_norm.weight"], vecs[a + "group_norm.bias"], xn, B, N,
                        Cm, G, eps, 0, bf16, s, N * band_R));
        }
        const ConvW& cq = convs[attn_prefix + "qkv"];
        const int bn_c = fit_bn(Cm % 192 == 0 ? 192 : (Cm % 128 == 0 ? 128 : (Cm % 64 == 0 ? 64 : 16)),
                                static_cast<long long>(B_max) * (N / TG_BM), Cm);
        const int bn_n = Nk % 256 == 0 ? 256 : 128;
        PlainGemm g;
        memset(&g, 0, sizeof(g));
        // q | k | v = xn W^T + b
        g.A = xn; g.K = Cm; g.a_ld = Cm; g.rows_per_sample = N; g.a_B = B_max; g.a_shared = false;
        g.Bm = cq.fwd.Bm; g.b_rows = cq.fwd.rows; g.ldb = cq.fwd.ldb; g.b_zstride = 0; g.b_Z = 1; g.b_per_sample = false;
        g.BN = cq.fwd.BN; g.N = 3 * Cm; g.out = qkv; g.out_fp32 = 0; g.ldc = 3 * Cm; g.c_zstride = 0;
        g.bias = cq.bias; g.residual = nullptr; g.alpha = 1.f;
        CK(plain_gemm(ex, g));
        if (!ex.prepare && band_R > 1) CK(comm->allgather(qkv_all, static_cast<size_t>(N) * 3 * Cm * 2, s));
        // v^T per sample (the B operand of P V must be K-major along the key tokens)
        if (!ex.prepare)
            CK(transpose16(qkv_all + static_cast<size_t>(2 * Cm) * 2, ws + o_vt, B, Nk, Cm, 3 * Cm,
                           static_cast<long long>(Nk) * 3 * Cm, s));
        // scores = q k^T / sqrt(C) in fp32
        memset(&g, 0, sizeof(g));
        g.A = qkv; g.K = Cm; g.a_ld = 3 * Cm; g.rows_per_sample = N; g.a_B = B_max; g.a_shared = false;
        g.Bm = qkv_all + static_cast<size_t>(Cm) * 2; g.b_rows = Nk; g.ldb = 3 * Cm;
        g.b_zstride = static_cast<long long>(Nk) * 3 * Cm; g.b_Z = B_max; g.b_per_sample = true;
        g.BN = bn_n; g.N = Nk; g.alpha = 1.f / sqrtf(static_cast<float>(Cm));
        if (fuse_attn_fwd) {
            // fp32 row softmax without the N x Nk score matrix in memory: the product runs twice, first keeping only
            // {max, sum exp} per (row, N tile), then writing exp(s - max) / sum as the 16-bit probabilities
            g.rs_out = reinterpret_cast<float2*>(ws + o_S);
            CK(plain_gemm(ex, g));
            if (!ex.prepare)
                CK(rowstats_combine(ws + o_S, ws + o_RS, static_cast<long long>(B) * N, 2 * ((Nk + bn_n - 1) / bn_n), s));
            g.rs_out = nullptr;
            g.rs_in = reinterpret_cast<const float2*>(ws + o_RS);
            g.out = ws + o_P; g.out_fp32 = 0; g.ldc = Nk;
            CK(plain_gemm(ex, g));
        } else {
            g.out = ws + o_S; g.out_fp32 = 1; g.ldc = Nk;
            CK(plain_gemm(ex, g));
            if (!ex.prepare) CK(attn_softmax(reinterpret_cast<const float*>(ws + o_S), ws + o_P, static_cast<long long>(B) * N, Nk, bf16, s));
        }
        // o = p v
        memset(&g, 0, sizeof(g));
        g.A = ws + o_P; g.K = Nk; g.a_ld = Nk; g.rows_per_sample = N; g.a_B = B_max; g.a_shared = false;
        g.Bm = ws + o_vt; g.b_rows = Cm; g.ldb = Nk; g.b_zstride = static_cast<long long>(Cm) * Nk; g.b_Z = B_max;
        g.b_per_sample = true; g.BN = bn_c; g.N = Cm; g.out = ws + o_O; g.out_fp32 = 0; g.ldc = Cm;
        g.alpha = 1.f;
        CK(plain_gemm(ex, g));
        // proj + residual
        const ConvW& cp = convs[a + "proj_attn"];
        memset(&g, 0, sizeof(g));
        g.A = ws + o_O; g.K = Cm; g.a_ld = Cm; g.rows_per_sample = N; g.a_B = B_max; g.a_shared = false;
        g.Bm = cp.fwd.Bm; g.b_rows = cp.fwd.rows; g.ldb = cp.fwd.ldb; g.b_Z = 1; g.BN = cp.fwd.BN; g.N = Cm;
        g.out = ws + o_attn_out; g.out_fp32 = 0; g.ldc = Cm; g.bias = cp.bias; g.residual = x; g.alpha = 1.f;
        g.gn_sums = next_gn < 0 ? nullptr : gn_sums_ptr(next_gn); g.gn_cpg = Cm / G;
        CK(plain_gemm(ex, g));
        }
    return 0;
}

// ------------------------------------------------------------------------------------------------ forward
int Decoder::forward_impl(Exec& ex, const void* z, int z_f16, float in_scale) {
    const int B = ex.B;
    cudaStream_t s = ex.stream;
    const float eps = 1e-6f;
    const int hw = h * w;
    uint8_t* x = ws + o_x0;
    if (!ex.prepare) {
        CUDA_CK(cudaMemsetAsync(ws + o_gn, 0, static_cast<size_t>(n_gn) * B_max * G * 2 * sizeof(acc_t), s));
        CK(latent_in(z, z_f16, in_scale, vecs["post_quant_conv.weight"], vecs["post_quant_conv.bias"],
                     vecs["decoder.conv_in.weight.T"], vecs["decoder.conv_in.bias"], x, B, h, w, Cm, bf16, s, band_r * h,
                     h_total));
    }
    int cur_h = h, cur_w = w;
    int ri = 0;
    auto fused_ok = [&](int C) { return fuse_gn && (C / G) % 16 == 0; };
    static const char* const group_tag[] = {"dense", "g4", "g16", "g64", "grouped"};
    auto tag_of = [&](int groups) { return group_tag[groups == 1 ? 0 : groups == 4 ? 1 : groups == 16 ? 2 : groups == 64 ? 3 : 4]; };
    auto resnet = [&](int idx, uint8_t* xin, bool stats_ready) -> int {
        const bool last = idx + 1 == static_cast<int>(resnets.size());
        // stats of the resnet output feed the next GroupNorm: the next resnet's norm1, the attention norm after
        // mid.resnets.0, or conv_norm_out after the last resnet; a block that ends in a downsample conv re-computes
        // them after it
        int next_gn = last ? n_gn - 1 : (idx == 0 ? n_gn - 2 : 2 * (idx + 1));
        if (idx >= 2) {
            const int blk = (idx - 2) / (cfg.layers_per_block + 1), j = (idx - 2) % (cfg.layers_per_block + 1);
            if (down_after[blk] && j == cfg.layers_per_block) next_gn = -1;
        }
        return resnet_forward(ex, idx, xin, stats_ready, next_gn);
    };
    // mid block: resnet, attention, resnet (wf_unet_2d_blocks.py:457-464)
    CK(resnet(0, x, false));  // conv_in is not a tapgemm: its output statistics need the standalone pass
    x = ws + o_res_out[0];
    CK(attention_forward(ex, x, "decoder.mid_block.attentions.0.", 2));
    x = ws + o_attn_out;
    CK(resnet(1, x, true));
    x = ws + o_res_out[1];
    ri = 2;
    for (int i = 0; i < cfg.n_blocks; ++i) {
        for (int j = 0; j < cfg.layers_per_block + 1; ++j) {
            // the first resnet after a downsample has no fused stats for its input
            const bool ready = !(j == 0 && i > 0 && down_after[i - 1]);
            CK(resnet(ri, x, ready));
            x = ws + o_res_out[ri];
            ++ri;
        }
        if (down_after[i]) {
            char buf[128];
            snprintf(buf, sizeof(buf), "decoder.up_blocks.%d.downsamplers.0.conv", i);
            const ConvW& cd = convs[buf];
            cur_tag = "down";
            CK(conv_gemm(ex, cd.fwd, x, cd.cin, cur_h, cur_w, cur_h / 2, cur_w / 2, ws + o_down_out[i], cd.cout,
                         cd.bias, nullptr, nullptr, 0, nullptr, nullptr));
            x = ws + o_down_out[i];
            cur_h /= 2;
            cur_w /= 2;
        }
    }
    x_last = x;
    // conv_norm_out + SiLU + conv_out (wf_vae.py:228-230)
    const bool last_down = down_after[cfg.n_blocks - 1] != 0;
    const ConvW& co = convs["decoder.conv_out"];
    const bool axo = apply_fused(co.fwd);
    AxForm ao = {gn_sums_ptr(n_gn - 1), vecs["decoder.conv_norm_out.weight"], vecs["decoder.conv_norm_out.bias"], C_last / G, 1,
                 H_out * W_out * band_R, eps};
    if (!ex.prepare) {
        if (!fused_ok(C_last) || last_down) CK(gn_stats(x, gn_sums_ptr(n_gn - 1), B, H_out * W_out, C_last, G, bf16, s));
        CK(sync_sums(ex, gn_sums_ptr(n_gn - 1)));
        if (!axo)
            CK(gn_apply(x, ao.stats, ao.gamma, ao.beta, ws + o_act, B, H_out * W_out, C_last, G, eps, 1, bf16, s,
                        H_out * W_out * band_R));
    }
    cur_tag = "conv_out";
    CK(conv_gemm(ex, co.fwd, axo ? x : ws + o_act, C_last, H_out, W_out, H_out, W_out, ws + o_logits, T_pad, co.bias, nullptr,
                 nullptr, 0, nullptr, nullptr, axo ? &ao : nullptr));
    return 0;
}

// ------------------------------------------------------------------------------------------------ backward
int Decoder::backward_impl(Exec& ex, const void* dlogits_in, const float* scale_b, float* dz) {
    const int B = ex.B;
    cudaStream_t s = ex.stream;
    const float eps = 1e-6f;
    const int hw = h * w;
    const void* dlog = dlogits_in ? dlogits_in : ws + o_dlogits;
    if (!ex.prepare)
        CUDA_CK(cudaMemsetAsync(gn_red_ptr(0), 0, static_cast<size_t>(n_gn) * B_max * G * 2 * sizeof(acc_t), s));
    // GroupNorm-backward reduction fused into the epilogue of the contraction that produces dy (when the channel
    // groups are 16-column aligned); otherwise the standalone gn_bwd_reduce pass runs
    // (measured on B200: the fused form wins for the <= 768-channel layers whose tiles carry enough MMA work to hide the
    // extra epilogue math; for the 1536/3072-channel grouped layers the standalone pass is cheaper)
    auto fuse_ok = [&](int C) { return fuse_gn && (C / G) % 16 == 0 && C <= fuse_bwd_max_c; };
    auto make_gb = [&](int gn_idx, const void* xsaved, const std::string& norm, int C, int px, int silu) {
        GnBwdFuse gb;
        memset(&gb, 0, sizeof(gb));
        if (!fuse_ok(C)) return gb;
        gb.red = gn_red_ptr(gn_idx);
        gb.x = xsaved;
        gb.stats = gn_sums_ptr(gn_idx);
        gb.gamma = vecs[norm + ".weight"];
        gb.beta = vecs[norm + ".bias"];
        gb.cpg = C / G;
        gb.silu = silu;
        gb.inv_n = 1.f / (static_cast<float>(C / G) * static_cast<float>(px) * static_cast<float>(band_R));
        gb.eps = eps;
        return gb;
    };
    // gradient buffers rotate: `cur` holds d(out) of the layer being differentiated
    int cur = 0;
    auto gbuf = [&](int i) { return ws + o_g[i & 3]; };
    // conv_out^T, then conv_norm_out + SiLU backward
    const ConvW& co = convs["decoder.conv_out"];
    cur_tag = "conv_out";
    static const char* const group_tag[] = {"dense", "g4", "g16", "g64", "grouped"};
    auto tag_of = [&](int groups) { return group_tag[groups == 1 ? 0 : groups == 4 ? 1 : groups == 16 ? 2 : groups == 64 ? 3 : 4]; };
    {
        const GnBwdFuse gb = make_gb(n_gn - 1, x_last, "decoder.conv_norm_out", C_last, H_out * W_out, 1);
        CK(conv_gemm(ex, co.bwd, dlog, T_pad, H_out, W_out, H_out, W_out, gbuf(cur), C_last, nullptr, nullptr, nullptr, 0,
                     nullptr, &gb));
    }
    if (!ex.prepare) {
        CK(gn_backward(ex, gbuf(cur), x_last, n_gn - 1, vecs["decoder.conv_norm_out.weight"],
                       vecs["decoder.conv_norm_out.bias"], nullptr, gbuf(cur), H_out * W_out, C_last, 1,
                       fuse_ok(C_last) ? 0 : 1));
    }
    auto resnet_bwd = [&](int idx, const uint8_t* xin) -> int {
        const ResnetDesc& r = resnets[idx];
        cur_tag = tag_of(r.groups);
        const int phw = r.h * r.w;
        uint8_t* dout = gbuf(cur);
        uint8_t* t1 = gbuf(cur + 1);
        uint8_t* t2 = gbuf(cur + 2);
        uint8_t* t3 = gbuf(cur + 3);
        const uint8_t* h1 = ws + o_res_h1[idx];
        const int gi1 = 2 * idx, gi2 = 2 * idx + 1;
        const ConvW& c2 = convs[r.prefix + ".conv2"];
        {
            const GnBwdFuse gb = make_gb(gi2, h1, r.prefix + ".norm2", r.cout, phw, 1);
            CK(conv_gemm(ex, c2.bwd, dout, r.cout, r.h, r.w, r.h, r.w, t1, r.cout, nullptr, nullptr, nullptr, 0, nullptr, &gb));
        }
        if (!ex.prepare) {
            const float* ga = vecs[r.prefix + ".norm2.weight"];
            const float* be = vecs[r.prefix + ".norm2.bias"];
            CK(gn_backward(ex, t1, h1, gi2, ga, be, nullptr, t1, phw, r.cout, 1, fuse_ok(r.cout) ? 0 : 1));
        }
        const ConvW& c1 = convs[r.prefix + ".conv1"];
        {
            const GnBwdFuse gb = make_gb(gi1, xin, r.prefix + ".norm1", r.cin, phw, 1);
            CK(conv_gemm(ex, c1.bwd, t1, r.cout, r.h, r.w, r.h, r.w, t2, r.cin, nullptr, nullptr, nullptr, 0, nullptr, &gb));
        }
        const void* add = dout;
        if (r.cin != r.cout) {
            const ConvW& cs = convs[r.prefix + ".conv_shortcut"];
            CK(conv_gemm(ex, cs.bwd, dout, r.cout, r.h, r.w, r.h, r.w, t3, r.cin, nullptr, nullptr, nullptr, 0, nullptr, nullptr));
            add = t3;
        }
        if (!ex.prepare) {
            const float* ga = vecs[r.prefix + ".norm1.weight"];
            const float* be = vecs[r.prefix + ".norm1.bias"];
            CK(gn_backward(ex, t2, xin, gi1, ga, be, add, t2, phw, r.cin, 1, fuse_ok(r.cin) ? 0 : 1));
        }
        cur = (cur + 2) & 3;
        return 0;
    };
    // decoder blocks in reverse; a DownDecoderBlock2D first undoes its stride-2 conv
    {
        const int n_res = cfg.layers_per_block + 1;
        for (int i = cfg.n_blocks - 1; i >= 0; --i) {
            const int last_idx = 2 + i * n_res + n_res - 1;
            if (down_after[i]) {
                const ResnetDesc& r = resnets[last_idx];  // its plane is the conv's input plane
                cur_tag = "down";
                char buf[128];
                uint8_t* dout = gbuf(cur);
                uint8_t* dxin = gbuf(cur + 1);
                for (int py = 0; py < 2; ++py)
                    for (int px = 0; px < 2; ++px) {
                        snprintf(buf, sizeof(buf), "decoder.up_blocks.%d.downsamplers.0.conv.bwd%d", i, py * 2 + px);
                        const ConvW& cb = convs[buf];
                        OutStrides os;
                        os.sx = 2LL * r.cout;
                        os.sy = 2LL * r.w * r.cout;
                        os.sb = static_cast<long long>(r.h) * r.w * r.cout;
                        uint8_t* base = dxin + (static_cast<size_t>(py) * r.w + px) * r.cout * 2;
                        CK(conv_gemm(ex, cb.bwd, dout, r.cout, r.h / 2, r.w / 2, r.h / 2, r.w / 2, base, r.cout, nullptr,
                                     nullptr, nullptr, 0, &os, nullptr));
                    }
                cur = (cur + 1) & 3;
            }
            for (int j = n_res - 1; j >= 0; --j) {
                const int idx = 2 + i * n_res + j;
                const uint8_t* xin;
                if (j > 0) xin = ws + o_res_out[idx - 1];
                else if (i == 0) xin = ws + o_res_out[1];
                else if (down_after[i - 1]) xin = ws + o_down_out[i - 1];
                else xin = ws + o_res_out[idx - 1];
                CK(resnet_bwd(idx, xin));
            }
        }
    }
    CK(resnet_bwd(1, ws + o_attn_out));
    {
        // attention backward. In a band plan (N local queries, Nk keys) dQ is complete locally, while dK and dV collect a
        // term from every band's queries: the fp32 partials [Nk, dK | dV] are reduce-scattered over the bands.
        const int N = hw, Nk = hw * band_R;
        cur_tag = "attn";
        const bool split = band_R > 1;
        const std::string a = "decoder.mid_block.attentions.0.";
        uint8_t* dout = gbuf(cur);
        uint8_t* dxn = gbuf(cur + 1);
        const uint8_t* xin = ws + o_res_out[0];
        const uint8_t* qkv_all = ws + o_qkv;
        const uint8_t* qkv = qkv_all + static_cast<size_t>(band_r) * N * 3 * Cm * 2;
        uint8_t* dqkv = ws + o_dqkv;
        uint8_t* dkv = ws + o_dkv;
        const ConvW& cp = convs[a + "proj_attn"];
        const int bn_c = fit_bn(Cm % 192 == 0 ? 192 : (Cm % 128 == 0 ? 128 : (Cm % 64 == 0 ? 64 : 16)),
                                static_cast<long long>(B_max) * (N / TG_BM), Cm);
        const int bn_n = Nk % 256 == 0 ? 256 : 128;
        PlainGemm g;
        // dO = dout Wp
        memset(&g, 0, sizeof(g));
        g.A = dout; g.K = Cm; g.a_ld = Cm; g.rows_per_sample = N; g.a_B = B_max;
        g.Bm = cp.bwd.Bm; g.b_rows = cp.bwd.rows; g.ldb = cp.bwd.ldb; g.b_Z = 1; g.BN = cp.bwd.BN; g.N = Cm;
        g.out = ws + o_dO; g.ldc = Cm; g.alpha = 1.f;
        CK(plain_gemm(ex, g));
        if (!ex.prepare) {
            CK(transpose16(ws + o_dO, ws + o_dOT, B, N, Cm, Cm, static_cast<long long>(N) * Cm, s));
            if (!mmajor_a) CK(transpose16(ws + o_P, ws + o_PT, B, N, Nk, Nk, static_cast<long long>(N) * Nk, s));
            CK(transpose16(qkv, ws + o_QT, B, N, Cm, 3 * Cm, static_cast<long long>(N) * 3 * Cm, s));
            CK(transpose16(qkv_all + static_cast<size_t>(Cm) * 2, ws + o_KT, B, Nk, Cm, 3 * Cm,
                           static_cast<long long>(Nk) * 3 * Cm, s));
        }
        // dV = P^T dO  -> dqkv[:, 2C:3C] (band plan: partial over this band's queries, fp32)
        memset(&g, 0, sizeof(g));
        // (P is read as stored, M-major: no transposed copy of the N x Nk matrix)
        if (mmajor_a) { g.A = ws + o_P; g.a_ld = Nk; g.a_mn_major = true; }
        else { g.A = ws + o_PT; g.a_ld = N; }
        g.K = N; g.rows_per_sample = Nk; g.a_B = B_max;
        g.Bm = ws + o_dOT; g.b_rows = Cm; g.ldb = N; g.b_zstride = static_cast<long long>(Cm) * N; g.b_Z = B_max;
        g.b_per_sample = true; g.BN = bn_c; g.N = Cm; g.alpha = 1.f;
        if (split) { g.out = dkv + static_cast<size_t>(Cm) * 4; g.out_fp32 = 1; g.ldc = 2 * Cm; }
        else { g.out = dqkv + static_cast<size_t>(2 * Cm) * 2; g.ldc = 3 * Cm; }
        CK(plain_gemm(ex, g));
        // dP = dO V^T and the softmax backward dS = P * (dP - rowdot) / sqrt(C). Fused form (default): rowdot[i] =
        // sum_j dP[i,j] P[i,j] = sum_c dO[i,c] O[i,c] comes from a small [N, C] pass and the product's epilogue writes dS
        // directly, so the fp32 N x Nk matrix dP is neither written nor read (WFD_NO_FUSE_ATTN_BWD=1: the two-kernel form)
        memset(&g, 0, sizeof(g));
        g.A = ws + o_dO; g.K = Cm; g.a_ld = Cm; g.rows_per_sample = N; g.a_B = B_max;
        g.Bm = qkv_all + static_cast<size_t>(2 * Cm) * 2; g.b_rows = Nk; g.ldb = 3 * Cm;
        g.b_zstride = static_cast<long long>(Nk) * 3 * Cm; g.b_Z = B_max; g.b_per_sample = true;
        g.BN = bn_n; g.N = Nk; g.alpha = 1.f;
        if (fuse_attn_bwd) {
            if (!ex.prepare)
                CK(rowdot16(ws + o_dO, ws + o_O, reinterpret_cast<float*>(ws + o_D), static_cast<long long>(B) * N, Cm, bf16, s));
            g.out = ws + o_dS; g.out_fp32 = 0; g.ldc = Nk; g.residual = ws + o_P;
            g.sb_d = reinterpret_cast<const float*>(ws + o_D); g.sb_scale = 1.f / sqrtf(static_cast<float>(Cm));
            CK(plain_gemm(ex, g));
        } else {
            g.out = ws + o_S; g.out_fp32 = 1; g.ldc = Nk;
            CK(plain_gemm(ex, g));
            if (!ex.prepare)
                CK(attn_softmax_bwd(ws + o_P, reinterpret_cast<const float*>(ws + o_S), ws + o_dS, static_cast<long long>(B) * N,
                                    Nk, 1.f / sqrtf(static_cast<float>(Cm)), bf16, s));
        }
        if (!ex.prepare && !mmajor_a) CK(transpose16(ws + o_dS, ws + o_dST, B, N, Nk, Nk, static_cast<long long>(N) * Nk, s));
        // dQ = dS K -> dqkv[:, 0:C]
        memset(&g, 0, sizeof(g));
        g.A = ws + o_dS; g.K = Nk; g.a_ld = Nk; g.rows_per_sample = N; g.a_B = B_max;
        g.Bm = ws + o_KT; g.b_rows = Cm; g.ldb = Nk; g.b_zstride = static_cast<long long>(Cm) * Nk; g.b_Z = B_max;
        g.b_per_sample = true; g.BN = bn_c; g.N = Cm; g.out = dqkv; g.ldc = 3 * Cm; g.alpha = 1.f;
        CK(plain_gemm(ex, g));
        // dK = dS^T Q -> dqkv[:, C:2C]
        memset(&g, 0, sizeof(g));
        if (mmajor_a) { g.A = ws + o_dS; g.a_ld = Nk; g.a_mn_major = true; }
        else { g.A = ws + o_dST; g.a_ld = N; }
        g.K = N; g.rows_per_sample = Nk; g.a_B = B_max;
        g.Bm = ws + o_QT; g.b_rows = Cm; g.ldb = N; g.b_zstride = static_cast<long long>(Cm) * N; g.b_Z = B_max;
        g.b_per_sample = true; g.BN = bn_c; g.N = Cm; g.alpha = 1.f;
        if (split) { g.out = dkv; g.out_fp32 = 1; g.ldc = 2 * Cm; }
        else { g.out = dqkv + static_cast<size_t>(Cm) * 2; g.ldc = 3 * Cm; }
        CK(plain_gemm(ex, g));
        if (!ex.prepare && split) {
            float* part = reinterpret_cast<float*>(dkv);
            const size_t per_rank = static_cast<size_t>(N) * 2 * Cm;
            CK(comm->reduce_scatter_f32(part, per_rank, s));
            CK(cvt_f32_to_16_2d(part + static_cast<size_t>(band_r) * per_rank, 2 * Cm, dqkv + static_cast<size_t>(Cm) * 2,
                                3 * Cm, N, 2 * Cm, bf16, s));
        }
        // d xn = [dQ | dK | dV] Wqkv
        const ConvW& cq = convs[a + "qkv"];
        memset(&g, 0, sizeof(g));
        g.A = dqkv; g.K = 3 * Cm; g.a_ld = 3 * Cm; g.rows_per_sample = N; g.a_B = B_max;
        g.Bm = cq.bwd.Bm; g.b_rows = cq.bwd.rows; g.ldb = cq.bwd.ldb; g.b_Z = 1; g.BN = cq.bwd.BN; g.N = Cm;
        g.out = dxn; g.ldc = Cm; g.alpha = 1.f;
        const GnBwdFuse gba = make_gb(n_gn - 2, xin, a + "group_norm", Cm, N, 0);
        g.gb = &gba;
        CK(plain_gemm(ex, g));
        if (!ex.prepare) {
            const float* ga = vecs[a + "group_norm.weight"];
            const float* be = vecs[a + "group_norm.bias"];
            CK(gn_backward(ex, dxn, xin, n_gn - 2, ga, be, dout, dxn, N, Cm, 0, fuse_ok(Cm) ? 0 : 1));
        }
        cur = (cur + 1) & 3;
    }
    CK(resnet_bwd(0, ws + o_x0));
    // conv_in^T + post_quant_conv^T. Single plans contract the Cm gradient channels against the 36 (tap, latent channel)
    // weight rows on the tensor cores (the fp32 product lands in the idle score buffer) and a small kernel gathers the
    // nine taps per latent pixel; row-band plans keep the CUDA-core kernel that reads the neighbour bands' halo rows.
    if (tc_latent && !band && convs.count("decoder.conv_in.T36") != 0) {
        const ConvW& ct = convs["decoder.conv_in.T36"];
        cur_tag = "dense";
        PlainGemm g;
        memset(&g, 0, sizeof(g));
        g.A = gbuf(cur); g.K = Cm; g.a_ld = Cm; g.rows_per_sample = hw; g.a_B = B_max;
        g.Bm = ct.fwd.Bm; g.b_rows = ct.fwd.rows; g.ldb = ct.fwd.ldb; g.b_Z = 1; g.BN = ct.fwd.BN; g.N = 48;
        g.out = ws + o_S; g.out_fp32 = 1; g.ldc = 48; g.alpha = 1.f;
        CK(plain_gemm(ex, g));
        if (!ex.prepare)
            CK(latent_gather(reinterpret_cast<const float*>(ws + o_S), 48, last_in_scale, scale_b, vecs["post_quant_conv.weight"],
                             dz, B, h, w, s));
        return 0;
    }
    if (!ex.prepare)
    {
        if (band) CK(halo_rows(ex, gbuf(cur), h, w, Cm));  // conv_in^T reads the neighbour bands' edge rows
        CK(latent_out(gbuf(cur), last_in_scale, scale_b, vecs["post_quant_conv.weight"], vecs["decoder.conv_in.weight.T"],
                      dz, B, h, w, Cm, bf16, s, band_r * h, h_total));
    }
    return 0;
}

int Decoder::forward(const void* z, int z_dtype, int B, float in_scale, cudaStream_t s) {
    if (!prepared) {
        set_error("decoder: weights not finalized or workspace not bound");
        return -12;
    }
    if (B < 1 || B > B_max) {
        set_error("decoder: batch %d outside [1, %d]", B, B_max);
        return -3;
    }
    Exec ex;
    ex.prepare = false;
    ex.B = B;
    ex.stream = s;
    ex.ops = &fwd_ops;
    last_in_scale = in_scale;
    last_B = B;
    return forward_impl(ex, z, z_dtype, in_scale);
}

int Decoder::backward(const void* dlogits, int B, const float* scale_b, float* dz, cudaStream_t s) {
    if (!prepared) {
        set_error("decoder: weights not finalized or workspace not bound");
        return -12;
    }
    if (B != last_B) {
        set_error("decoder: backward batch %d does not match the last forward (%d)", B, last_B);
        return -3;
    }
    Exec ex;
    ex.prepare = false;
    ex.B = B;
    ex.stream = s;
    ex.ops = &bwd_ops;
    return backward_impl(ex, dlogits, scale_b, dz);
}

}  // namespace wfd
