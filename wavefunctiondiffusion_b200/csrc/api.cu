// extern "C" surface of libwfd_b200.so (declared in include/wfd_b200.h).
#include "decoder.cuh"
#include "kernels.cuh"
#include "../../include/wfd_b200_test.h"
#include <stdlib.h>
#include <string.h>
#include <new>

using namespace wfd;

struct wfd_decoder {
    Decoder d;
};
struct wfd_wfc {
    Wfc w;
};
struct wfd_encoder {
    Encoder e;
};
struct wfd_comm {
    Comm* c = nullptr;
    ~wfd_comm() { delete c; }
};
struct wfd_loop_group {
    LoopGroup* g = nullptr;
};

#define API_CUDA_CK(expr)                                                       \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) {                                               \
            set_error("%s: %s", #expr, cudaGetErrorString(e__));                \
            return -8;                                                          \
        }                                                                       \
    } while (0)

static inline cudaStream_t S(void* stream) { return static_cast<cudaStream_t>(stream); }

// include/wfd_b200_test.h: the hooks answer only in a process started with WFD_ENABLE_TEST_HOOKS=1
static int test_hooks_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("WFD_ENABLE_TEST_HOOKS");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    if (!v) set_error("test hooks are disabled (start the process with WFD_ENABLE_TEST_HOOKS=1)");
    return v;
}


extern "C" {

int wfd_version(void) { return WFD_ABI_VERSION; }
const char* wfd_last_error(void) { return last_error(); }
long long wfd_launch_count(void) { return g_launch_count; }

int wfd_decoder_create(const wfd_decoder_config* cfg, int h, int w, int B_max, int dtype16, wfd_decoder** out) {
    if (!cfg || !out) {
        set_error("wfd_decoder_create: null argument");
        return -1;
    }
    wfd_decoder* p = new (std::nothrow) wfd_decoder();
    if (!p) {
        set_error("wfd_decoder_create: out of host memory");
        return -1;
    }
    int rc = p->d.init(*cfg, h, w, B_max, dtype16);
    if (rc < 0) {
        delete p;
        return rc;
    }
    *out = p;
    return 0;
}
void wfd_decoder_destroy(wfd_decoder* d) { delete d; }
int wfd_decoder_out_hw(const wfd_decoder* d, int* H, int* W) {
    if (!d) return -1;
    if (H) *H = d->d.H_out;
    if (W) *W = d->d.W_out;
    return 0;
}
int wfd_decoder_set_weight(wfd_decoder* d, const char* key, const float* host_data, long long numel) {
    if (!d || !key || !host_data) {
        set_error("wfd_decoder_set_weight: null argument");
        return -1;
    }
    return d->d.set_weight(key, host_data, numel);
}
int wfd_decoder_finalize_weights(wfd_decoder* d) { return d ? d->d.finalize_weights() : -1; }
size_t wfd_decoder_workspace_bytes(const wfd_decoder* d) { return d ? d->d.ws_bytes : 0; }
int wfd_decoder_bind_workspace(wfd_decoder* d, void* ws, size_t bytes) { return d ? d->d.bind_workspace(ws, bytes) : -1; }
int wfd_decoder_forward(wfd_decoder* d, const void* z, int z_dtype, int B, float in_scale, void* stream) {
    if (!d || !z) {
        set_error("wfd_decoder_forward: null argument");
        return -1;
    }
    return d->d.forward(z, z_dtype, B, in_scale, S(stream));
}
void* wfd_decoder_logits(wfd_decoder* d) { return d && d->d.ws ? d->d.ws + d->d.o_logits : nullptr; }
void* wfd_decoder_dlogits(wfd_decoder* d) { return d && d->d.ws ? d->d.ws + d->d.o_dlogits : nullptr; }

int wfd_decoder_export_logits(wfd_decoder* d, int B, void* out, int out_dtype, int layout, void* stream) {
    if (!d || !out || !d->d.ws) {
        set_error("wfd_decoder_export_logits: null argument / unbound workspace");
        return -1;
    }
    Decoder& D = d->d;
    const int HW = D.H_out * D.W_out;
    const void* src = D.ws + D.o_logits;
    if (layout == 0) {
        if (out_dtype > 1) {
            set_error("export_logits: NCHW export supports fp32/fp16");
            return -3;
        }
        return nhwc16_to_nchw(src, out, out_dtype, B, D.T_pad, HW, D.bf16, S(stream));
    }
    const long long n = static_cast<long long>(B) * HW * D.T_pad;
    if (out_dtype == 0) return cvt16_to_f32(src, static_cast<float*>(out), n, D.bf16, S(stream));
    if ((out_dtype == 2 && D.bf16) || (out_dtype == 1 && !D.bf16)) {
        API_CUDA_CK(cudaMemcpyAsync(out, src, n * 2, cudaMemcpyDeviceToDevice, S(stream)));
        return 0;
    }
    set_error("export_logits: channels-last 16-bit export must match the plan dtype");
    return -3;
}

int wfd_decoder_import_dlogits(wfd_decoder* d, int B, const void* in, int in_dtype, int layout, void* stream) {
    if (!d || !in || !d->d.ws) {
        set_error("wfd_decoder_import_dlogits: null argument / unbound workspace");
        return -1;
    }
    Decoder& D = d->d;
    const int HW = D.H_out * D.W_out;
    void* dst = D.ws + D.o_dlogits;
    if (layout == 0) {
        if (in_dtype > 1) {
            set_error("import_dlogits: NCHW import supports fp32/fp16");
            return -3;
        }
        return nchw_to_nhwc16(in, in_dtype, dst, B, D.T_pad, HW, D.bf16, S(stream));
    }
    const long long n = static_cast<long long>(B) * HW * D.T_pad;
    if (in_dtype == 0) return cvt_f32_to_16(static_cast<const float*>(in), dst, n, D.bf16, S(stream));
    if ((in_dtype == 2 && D.bf16) || (in_dtype == 1 && !D.bf16)) {
        API_CUDA_CK(cudaMemcpyAsync(dst, in, n * 2, cudaMemcpyDeviceToDevice, S(stream)));
        return 0;
    }
    set_error("import_dlogits: channels-last 16-bit import must match the plan dtype");
    return -3;
}

int wfd_decoder_backward(wfd_decoder* d, const void* dlogits, int B, const float* scale_b, float* dz, void* stream) {
    if (!d || !dz) {
        set_error("wfd_decoder_backward: null argument");
        return -1;
    }
    Decoder& D = d->d;
    if (dlogits && D.ws && dlogits != D.ws + D.o_dlogits) {
        const size_t n = static_cast<size_t>(B) * D.H_out * D.W_out * D.T_pad * 2;
        API_CUDA_CK(cudaMemcpyAsync(D.ws + D.o_dlogits, dlogits, n, cudaMemcpyDeviceToDevice, S(stream)));
    }
    return D.backward(nullptr, B, scale_b, dz, S(stream));
}

int wfd_encoder_create(const wfd_decoder_config* cfg, int H, int W, int B_max, int dtype16, wfd_encoder** out) {
    if (!cfg || !out) {
        set_error("wfd_encoder_create: null argument");
        return -1;
    }
    wfd_encoder* p = new (std::nothrow) wfd_encoder();
    if (!p) {
        set_error("wfd_encoder_create: out of host memory");
        return -1;
    }
    int rc = p->e.init_enc(*cfg, H, W, B_max, dtype16);
    if (rc < 0) {
        delete p;
        return rc;
    }
    *out = p;
    return 0;
}
void wfd_encoder_destroy(wfd_encoder* e) { delete e; }
int wfd_encoder_set_weight(wfd_encoder* e, const char* key, const float* host_data, long long numel) {
    if (!e || !key || !host_data) {
        set_error("wfd_encoder_set_weight: null argument");
        return -1;
    }
    return e->e.set_weight(key, host_data, numel);
}
int wfd_encoder_finalize_weights(wfd_encoder* e) { return e ? e->e.finalize_enc() : -1; }
size_t wfd_encoder_workspace_bytes(const wfd_encoder* e) { return e ? e->e.ws_bytes : 0; }
int wfd_encoder_bind_workspace(wfd_encoder* e, void* ws, size_t bytes) { return e ? e->e.bind_workspace_enc(ws, bytes) : -1; }
int wfd_encoder_forward_wave(wfd_encoder* e, const void* wave, int dtype, int B, float* moments, void* stream) {
    if (!e || !wave || !moments) {
        set_error("wfd_encoder_forward_wave: null argument");
        return -1;
    }
    return e->e.encode_wave(wave, dtype, B, moments, S(stream));
}
int wfd_encoder_forward_tiles(wfd_encoder* e, const long long* tiles, int B, float* moments, void* stream) {
    if (!e || !tiles || !moments) {
        set_error("wfd_encoder_forward_tiles: null argument");
        return -1;
    }
    return e->e.encode_tiles(tiles, B, moments, S(stream));
}

int wfd_wfc_create(const uint8_t* mat_x, const uint8_t* mat_y, int T, int T_pad, int dtype16, wfd_wfc** out) {
    if (!mat_x || !mat_y || !out) {
        set_error("wfd_wfc_create: null argument");
        return -1;
    }
    wfd_wfc* p = new (std::nothrow) wfd_wfc();
    if (!p) {
        set_error("wfd_wfc_create: out of host memory");
        return -1;
    }
    int rc = p->w.init(mat_x, mat_y, T, T_pad, dtype16);
    if (rc < 0) {
        delete p;
        return rc;
    }
    *out = p;
    return 0;
}
void wfd_wfc_destroy(wfd_wfc* h) { delete h; }
size_t wfd_wfc_saved_bytes(const wfd_wfc* h, int B, int H, int W) { return h ? h->w.saved_bytes(B, H, W) : 0; }
int wfd_wfc_forward(wfd_wfc* h, const void* in, int is_logits, int B, int H, int W, float strength, void* saved,
                    float* loss, void* stream) {
    if (!h || !in || !saved || !loss) {
        set_error("wfd_wfc_forward: null argument");
        return -1;
    }
    return h->w.forward(in, is_logits, B, H, W, strength, saved, loss, S(stream));
}
int wfd_wfc_backward(wfd_wfc* h, void* saved, const float* dloss, void* dout, void* stream) {
    if (!h || !saved || !dout) {
        set_error("wfd_wfc_backward: null argument");
        return -1;
    }
    return h->w.backward(saved, dloss, dout, S(stream));
}
void* wfd_wfc_saved_wave(const wfd_wfc* h, void* saved, int B, int H, int W) {
    (void)B; (void)H;
    if (!saved) return nullptr;
    size_t halo = 0;
    if (h && h->w.band) halo = (static_cast<size_t>(W) * h->w.T_pad * 2 + 255) & ~size_t(255);
    return static_cast<uint8_t*>(saved) + 256 + halo;
}

int wfd_comm_unique_id(const char* libnccl_path, void* id128) {
    if (!id128) return -1;
    return nccl_unique_id(libnccl_path, id128);
}
int wfd_comm_create_nccl(const char* libnccl_path, const void* id128, int rank, int nranks, wfd_comm** out) {
    if (!id128 || !out) {
        set_error("wfd_comm_create_nccl: null argument");
        return -1;
    }
    Comm* c = nullptr;
    int rc = nccl_comm_create(libnccl_path, id128, rank, nranks, &c);
    if (rc < 0) return rc;
    *out = new wfd_comm();
    (*out)->c = c;
    return 0;
}
wfd_loop_group* wfd_loop_group_create(int nranks) {
    LoopGroup* g = loop_group_create(nranks);
    if (!g) return nullptr;
    wfd_loop_group* p = new wfd_loop_group();
    p->g = g;
    return p;
}
void wfd_loop_group_destroy(wfd_loop_group* g) {
    if (!g) return;
    loop_group_destroy(g->g);
    delete g;
}
int wfd_comm_create_loopback(wfd_loop_group* g, int rank, wfd_comm** out) {
    if (!g || !out) {
        set_error("wfd_comm_create_loopback: null argument");
        return -1;
    }
    Comm* c = nullptr;
    int rc = loop_comm_create(g->g, rank, &c);
    if (rc < 0) return rc;
    *out = new wfd_comm();
    (*out)->c = c;
    return 0;
}
void wfd_comm_destroy(wfd_comm* c) { delete c; }
int wfd_comm_peer_handle(wfd_comm* c, size_t max_row_bytes, void* handle64) {
    if (!c || !c->c || !handle64) {
        set_error("wfd_comm_peer_handle: null argument");
        return -1;
    }
    return c->c->peer_local_handle(max_row_bytes, handle64);
}
int wfd_comm_peer_attach(wfd_comm* c, const void* handles) {
    if (!c || !c->c || !handles) {
        set_error("wfd_comm_peer_attach: null argument");
        return -1;
    }
    return c->c->peer_attach(handles);
}

int wfd_decoder_create_band(const wfd_decoder_config* cfg, int h, int w, int rank, int nranks, int dtype16,
                            wfd_decoder** out) {
    if (!cfg || !out) {
        set_error("wfd_decoder_create_band: null argument");
        return -1;
    }
    if (rank < 0) {
        set_error("wfd_decoder_create_band: negative rank");
        return -3;
    }
    wfd_decoder* p = new (std::nothrow) wfd_decoder();
    if (!p) {
        set_error("wfd_decoder_create_band: out of host memory");
        return -1;
    }
    int rc = p->d.init(*cfg, h, w, 1, dtype16, rank, nranks);
    if (rc < 0) {
        delete p;
        return rc;
    }
    *out = p;
    return 0;
}
int wfd_decoder_attach_comm(wfd_decoder* d, wfd_comm* c) {
    if (!d || !c || !c->c) {
        set_error("wfd_decoder_attach_comm: null argument");
        return -1;
    }
    if (!d->d.band || c->c->nranks != d->d.band_R || c->c->rank != d->d.band_r) {
        set_error("wfd_decoder_attach_comm: communicator rank/size does not match the band plan");
        return -3;
    }
    d->d.comm = c->c;
    return 0;
}
int wfd_wfc_attach_comm(wfd_wfc* h, wfd_comm* c, int H_total) {
    if (!h || !c || !c->c) {
        set_error("wfd_wfc_attach_comm: null argument");
        return -1;
    }
    return h->w.set_band(c->c, H_total);
}

int wfd_guidance_step(wfd_decoder* d, wfd_wfc* h, const void* z, int z_dtype, int B, float in_scale, float strength,
                      void* wfc_saved, float* loss, float* dz, void* stream) {
    if (!d || !h || !z || !wfc_saved || !loss || !dz) {
        set_error("wfd_guidance_step: null argument");
        return -1;
    }
    Decoder& D = d->d;
    if (h->w.T_pad != D.T_pad || h->w.bf16 != D.bf16) {
        set_error("wfd_guidance_step: decoder (T_pad=%d) and wfc handle (T_pad=%d) disagree", D.T_pad, h->w.T_pad);
        return -3;
    }
    int rc = D.forward(z, z_dtype, B, in_scale, S(stream));
    if (rc < 0) return rc;
    rc = h->w.forward(D.ws + D.o_logits, 1, B, D.H_out, D.W_out, strength, wfc_saved, loss, S(stream));
    if (rc < 0) return rc;
    rc = h->w.backward(wfc_saved, nullptr, D.ws + D.o_dlogits, S(stream));
    if (rc < 0) return rc;
    return D.backward(nullptr, B, nullptr, dz, S(stream));
}

double wfd_wfc_kc_kept(const wfd_wfc* h) { return h ? h->w.kc_kept : 0.0; }
int wfd_guidance_loop(wfd_decoder* d, wfd_wfc* h, float* lat, const float* eps_uncond, const float* eps_text, float cfg_scale,
                      float a, float b, int B, float in_scale, float strength, int n_steps, void* wfc_saved, float* loss_out,
                      float* z_scratch, float* dz_scratch, void* stream) {
    if (!d || !h || !lat || !eps_uncond || !wfc_saved || !loss_out || !z_scratch || !dz_scratch) {
        set_error("wfd_guidance_loop: null argument");
        return -1;
    }
    if (d->d.band) {
        set_error("wfd_guidance_loop: row-band plans take the whole plane per call; use wfd_guidance_step");
        return -3;
    }
    const long long n = static_cast<long long>(B) * 4 * d->d.h * d->d.w;
    for (int s = 0; s < n_steps; ++s) {
        int rc = guidance_affine(lat, eps_uncond, eps_text, cfg_scale, a, b, z_scratch, n, S(stream));
        if (rc < 0) return rc;
        rc = wfd_guidance_step(d, h, z_scratch, 0, B, in_scale, strength, wfc_saved, loss_out + static_cast<size_t>(s) * B,
                               dz_scratch, stream);
        if (rc < 0) return rc;
        rc = guidance_update(lat, dz_scratch, a, n, S(stream));
        if (rc < 0) return rc;
    }
    return 0;
}

int wfd_wfc_set_sparse(wfd_wfc* h, int on) {
    if (!h) {
        set_error("wfd_wfc_set_sparse: null handle");
        return -1;
    }
    return h->w.set_sparse(on);
}

int wfd_wfc_count_violations(const wfd_wfc* h, const long long* tiles, int B, int H, int W, long long* out, void* stream) {
    if (!h || !tiles || !out) {
        set_error("wfd_wfc_count_violations: null argument");
        return -1;
    }
    const Wfc& w = h->w;
    return count_violations(tiles, w.rules, w.rules + static_cast<size_t>(w.T) * w.T, w.T, B, H, W, out, S(stream));
}

int wfd_argmax_tiles(const void* x, int dtype, long long rows, int n, long long* out, void* stream) {
    if (!x || !out) {
        set_error("wfd_argmax_tiles: null argument");
        return -1;
    }
    return argmax_rows(x, dtype, rows, n, out, S(stream));
}

int wfd_tiles_finish(const long long* tiles, int B, int H, int W, const int* symm, int n_symm, const int* shrink_to_gen,
                     int n_shrink, const int* gen_to_sc, int n_gen, long long* tiles_out, uint16_t* sc_out, long long* bad,
                     void* stream) {
    if (!tiles || !gen_to_sc || !bad) {
        set_error("wfd_tiles_finish: null argument");
        return -1;
    }
    return tiles_finish(tiles, B, H, W, symm, n_symm, shrink_to_gen, n_shrink, gen_to_sc, n_gen, tiles_out, sc_out, bad,
                        S(stream));
}

int wfd_nchw_to_nhwc16(const void* in, int in_dtype, void* out, int B, int C, int HW, int dtype16, void* stream) {
    return nchw_to_nhwc16(in, in_dtype, out, B, C, HW, dtype16, S(stream));
}
int wfd_nhwc16_to_nchw(const void* in, void* out, int out_dtype, int B, int C, int HW, int dtype16, void* stream) {
    return nhwc16_to_nchw(in, out, out_dtype, B, C, HW, dtype16, S(stream));
}
int wfd_cvt16_to_f32(const void* in, float* out, long long n, int dtype16, void* stream) {
    return cvt16_to_f32(in, out, n, dtype16, S(stream));
}

int wfd_test_tapgemm(const wfd_tapgemm_desc* t, int use_ref, void* stream) {
    if (!test_hooks_enabled()) return -13;
    if (!t) return -1;
    TapGemmOp op;
    memset(&op, 0, sizeof(op));
    TapGemmParams& p = op.p;
    p.BN = t->BN;
    p.n_tiles = (t->N + t->BN - 1) / t->BN;
    p.z_count = t->z_count > 0 ? t->z_count : 1;
    p.w_box = t->w_box;
    p.h_box = t->h_box;
    p.tiles_x = t->W / t->w_box;
    p.tiles_y = t->H / t->h_box;
    p.m_tiles = t->B * p.tiles_x * p.tiles_y;
    p.W = t->W;
    p.H = t->H;
    p.in_stride = t->in_stride > 0 ? t->in_stride : 1;
    p.n_taps = t->n_taps;
    p.cpt = t->cpt;
    for (int i = 0; i < t->n_taps && i < 9; ++i) {
        p.tap_dx[i] = t->tap_dx[i];
        p.tap_dy[i] = t->tap_dy[i];
    }
    p.a_c0 = t->a_c0;
    p.reuse = t->reuse;
    p.cg = t->cg;
    p.n_group = t->n_group;
    p.a_zmul = t->a_zmul;
    p.b_zmul = t->b_zmul;
    p.b_bbmul = t->b_bbmul;
    p.N = t->N;
    p.out = t->out;
    p.out_fp32 = t->out_fp32;
    p.out_sx = t->ldc;
    p.out_sy = t->ldc * t->W;
    p.out_sb = t->ldc * t->W * t->H;
    p.c_zstride = t->c_zstride;
    p.bias = t->bias;
    p.residual = t->residual;
    p.alpha = t->alpha;
    p.rowdot = static_cast<acc_t*>(t->rowdot);
    p.dotsrc = t->dotsrc;
    p.ld_dot = t->ld_dot;
    p.gn_sums = static_cast<acc_t*>(t->gn_sums);
    p.gn_cpg = t->gn_cpg;
    p.gn_G = 8;
    p.is_bf16 = t->dtype16;
    p.dbg = t->dbg;
    p.a_mn_major = t->a_mn_major;
    p.kc_list = t->kc_list;
    p.kc_cnt = t->kc_cnt;
    p.a_ptr = t->a;
    p.a_C = t->a_C;
    p.a_W = t->a_W;
    p.a_H = t->a_H;
    p.a_B = t->a_B;
    p.a_ld = t->a_ld;
    p.b_ptr = t->b;
    p.ldb = t->ldb;
    p.b_zstride = t->b_zstride;
    p.b_rows = t->b_rows;
    p.b_Z = t->b_Z;
    int rc = tapgemm_prepare(&op);
    if (rc < 0) return rc;
    return use_ref ? tapgemm_launch_ref(&op, S(stream)) : tapgemm_launch(&op, S(stream));
}
int wfd_profile_enable(int on) {
    profile_enable(on);
    return 0;
}
int wfd_profile_report(char* buf, size_t cap) { return profile_report(buf, cap); }
int wfd_debug_set_ref_gemm(int on) {
    if (!test_hooks_enabled()) return -13;
    debug_set_ref_gemm(on);
    return 0;
}

}  // extern "C"
