// WFC transition-rule loss (reference wfd/wfdf/losses.py:43-75, WFCLossEinsum) and its gradient.
//
// loss[b] = strength / P * ( sum_{y,x<W-1} p[y,x]^T (1-Ax) p[y,x+1] + sum_{y<H-1,x} p[y,x]^T (1-Ay) p[y+1,x] ),
// P = H(W-1) + (H-1)W, p = wave[:T] (the softmax itself runs over all T_pad channels, pipeline_wfd.py:405).
//
// With S_n = sum_{i<T} p[n,i] and the complement identity p^T (1-A) q = S_p S_q - p^T A q, the dense tensor-core
// contraction runs against the adjacency matrices themselves:
//     a[n,i] = sum_j Ax[i,j] p[right(n),j] + Ax[j,i] p[left(n),j] + Ay[i,j] p[below(n),j] + Ay[j,i] p[above(n),j]
// (one tapgemm with four taps and K = 4*T_pad), so that g'[n,i] = [i<T] c_n - a[n,i], c_n = sum of S over the in-map
// neighbours, carries no cancellation through the 16-bit operands. Every adjacent pair is seen from both of its cells:
//     loss[b] = strength/P * 1/2 * sum_n ( S_n c_n - <p[n], a[n]> ).
#include "decoder.cuh"
#include "kernels.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

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

constexpr int WFC_BN = 256;

Wfc::~Wfc() {
    if (Bm) cudaFree(Bm);
    if (rules) cudaFree(rules);
    if (kc_list) cudaFree(kc_list);
    if (kc_cnt) cudaFree(kc_cnt);
}

int Wfc::init(const uint8_t* mx, const uint8_t* my, int T_, int T_pad_, int dtype16) {
    T = T_;
    T_pad = T_pad_;
    bf16 = dtype16 ? 1 : 0;
    if (T < 1 || T_pad < T || T_pad % TG_KC != 0) {
        set_error("wfc: need 1 <= T <= T_pad and T_pad %% 64 == 0 (T=%d T_pad=%d)", T, T_pad);
        return -3;
    }
    const uint16_t one = bf16 ? 0x3F80 : 0x3C00;
    const long long rows = static_cast<long long>((T_pad + WFC_BN - 1) / WFC_BN) * WFC_BN;
    const long long ldb = 4LL * T_pad;
    std::vector<uint16_t> hb(static_cast<size_t>(rows * ldb), 0);
    for (int i = 0; i < T; ++i) {
        uint16_t* row = hb.data() + static_cast<size_t>(i) * ldb;
        for (int j = 0; j < T; ++j) {
            if (mx[static_cast<size_t>(i) * T + j]) row[j] = one;                  // right neighbour:  Ax[i,j]
            if (mx[static_cast<size_t>(j) * T + i]) row[T_pad + j] = one;          // left neighbour:   Ax[j,i]
            if (my[static_cast<size_t>(i) * T + j]) row[2 * T_pad + j] = one;      // below:            Ay[i,j]
            if (my[static_cast<size_t>(j) * T + i]) row[3 * T_pad + j] = one;      // above:            Ay[j,i]
        }
    }
    CUDA_CK(cudaMalloc(&Bm, hb.size() * 2));
    CUDA_CK(cudaMemcpy(Bm, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    {
        const size_t tt = static_cast<size_t>(T) * T;
        std::vector<uint8_t> r(2 * tt);
        for (size_t i = 0; i < tt; ++i) {
            r[i] = mx[i] ? 1 : 0;
            r[tt + i] = my[i] ? 1 : 0;
        }
        CUDA_CK(cudaMalloc(&rules, r.size()));
        CUDA_CK(cudaMemcpy(rules, r.data(), r.size(), cudaMemcpyHostToDevice));
    }
    // Adjacency is very sparse (0.2 % for the StarCraft tilesets): about half of the [256 tiles] x [64 neighbour tiles]
    // blocks of the packed operand hold no allowed pair at all. Those K chunks are skipped - they add exact zeros, so the
    // result is bit-identical to the dense contraction (WFD_WFC_DENSE=1 keeps every chunk, for A/B runs).
    const char* dense_env = getenv("WFD_WFC_DENSE");
    if (!(dense_env && dense_env[0] == '1')) {
        const int n_tiles = static_cast<int>(rows / WFC_BN), k_iters = static_cast<int>(ldb / TG_KC);
        std::vector<int> list(static_cast<size_t>(n_tiles) * k_iters, 0), cnt(n_tiles, 0);
        long long kept = 0;
        for (int nt = 0; nt < n_tiles; ++nt) {
            int c = 0;
            for (int kc = 0; kc < k_iters; ++kc) {
                bool any = false;
                for (int r = 0; r < WFC_BN && !any; ++r) {
                    const uint16_t* row = hb.data() + (static_cast<size_t>(nt) * WFC_BN + r) * ldb + static_cast<size_t>(kc) * TG_KC;
                    for (int k = 0; k < TG_KC; ++k)
                        if (row[k]) {
                            any = true;
                            break;
                        }
                }
                if (any || (dense_env && dense_env[0] == '2')) {
                    // packed for the producer warps: dx+1 | (dy+1) << 2 | chunk-in-tap << 4 | chunk << 16
                    static const int tdx[4] = {1, -1, 0, 0}, tdy[4] = {0, 0, 1, -1};  // tap order of forward()
                    const int cpt = T_pad / TG_KC, tap = kc / cpt, cc = kc - tap * cpt;
                    list[static_cast<size_t>(nt) * k_iters + c++] = (tdx[tap] + 1) | ((tdy[tap] + 1) << 2) | (cc << 4) | (kc << 16);
                }  // WFD_WFC_DENSE=2: the list mechanism with every chunk kept
            }
            if (c == 0) c = 1;  // an N tile without any allowed pair still runs one (all-zero) chunk
            cnt[nt] = c;
            kept += c;
        }
        kc_kept = static_cast<double>(kept) / (static_cast<double>(n_tiles) * k_iters);
        CUDA_CK(cudaMalloc(&kc_list, list.size() * sizeof(int)));
        CUDA_CK(cudaMalloc(&kc_cnt, cnt.size() * sizeof(int)));
        CUDA_CK(cudaMemcpy(kc_list, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_CK(cudaMemcpy(kc_cnt, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    {
        const char* e = getenv("WFD_WFC_SPARSE");  // whole-suite runs of the alternative contraction
        if (e && e[0] == '1') return set_sparse(1);
    }
    return 0;
}

int Wfc::set_sparse(int on) {
    if (on && band) {
        set_error("wfc: the sparse contraction does not serve row-band handles");
        return -3;
    }
    if (on && !sparse.ent) {
        const size_t tt = static_cast<size_t>(T) * T;
        std::vector<uint8_t> r(2 * tt);
        CUDA_CK(cudaMemcpy(r.data(), rules, r.size(), cudaMemcpyDeviceToHost));
        CK(sparse.init(r.data(), r.data() + tt, T, T_pad));
    }
    use_sparse = on != 0;
    return 0;
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// Row bands keep one halo row of the wave and of the row sums above and below the band (vertical pairs, losses.py:60-66).
size_t Wfc::saved_bytes(int B, int H, int W) const {
    const size_t rows = static_cast<size_t>(B) * H * W;
    const size_t pad = band ? 2 * (align256(static_cast<size_t>(W) * T_pad * 2) + align256(static_cast<size_t>(W) * 4)) : 0;
    return 256 + 2 * align256(rows * T_pad * 2) + align256(rows * 4) + align256(rows * sizeof(acc_t)) + pad;
}

int Wfc::set_band(Comm* c, int H_total_) {
    if (!c || H_total_ < 1 || H_total_ % c->nranks != 0) {
        set_error("wfc: row bands need a communicator and a map height divisible by the rank count");
        return -3;
    }
    comm = c;
    band = true;
    band_R = c->nranks;
    band_r = c->rank;
    H_total = H_total_;
    op_wave = nullptr;
    return 0;
}

struct SavedView {
    WfcSavedHeader* hdr;
    uint8_t* wave;
    uint8_t* a;
    float* rowsum;
    acc_t* rowdot;  // fixed point 2^40
};
static SavedView view_saved(void* saved, size_t rows, int T_pad, size_t halo_wave = 0, size_t halo_sum = 0) {
    SavedView v;
    uint8_t* p = static_cast<uint8_t*>(saved);
    v.hdr = reinterpret_cast<WfcSavedHeader*>(p);
    p += 256 + halo_wave;
    v.wave = p;
    p += align256(rows * T_pad * 2) + halo_wave;
    v.a = p;
    p += align256(rows * T_pad * 2) + halo_sum;
    v.rowsum = reinterpret_cast<float*>(p);
    p += align256(rows * 4) + halo_sum;
    v.rowdot = reinterpret_cast<acc_t*>(p);
    return v;
}

int Wfc::forward(const void* in, int is_logits, int B, int H, int W, float strength, void* saved, float* loss,
                 cudaStream_t s) {
    if (B < 1 || H < 1 || W < 1 || (H == 1 && W == 1)) {
        set_error("wfc: bad map shape B=%d H=%d W=%d", B, H, W);
        return -3;
    }
    int wb = 0, hb = 0;
    for (int c = 128; c >= 1; c >>= 1) {
        if (W % c == 0 && H % (TG_BM / c) == 0) {
            wb = c;
            hb = TG_BM / c;
            break;
        }
    }
    if (!wb && !use_sparse) {  // the sparse contraction walks strips of four cells: any map shape
        set_error("wfc: map %dx%d cannot be tiled into 128-cell patches", H, W);
        return -3;
    }
    const size_t rows = static_cast<size_t>(B) * H * W;
    const size_t halo_wave = band ? align256(static_cast<size_t>(W) * T_pad * 2) : 0;
    const size_t halo_sum = band ? align256(static_cast<size_t>(W) * 4) : 0;
    if (band && (B != 1 || !is_logits || H * band_R != H_total)) {
        set_error("wfc: a row-band handle scores logits of one map, %d rows per band (got B=%d H=%d is_logits=%d)",
                  H_total / band_R, B, H, is_logits);
        return -3;
    }
    SavedView v = view_saved(saved, rows, T_pad, halo_wave, halo_sum);
    const void* wave = is_logits ? v.wave : in;
    const int Hm = band ? H_total : H;        // rows of the whole map
    const int y_off = band ? band_r * H : 0;  // first row of this band
    // the header lives in caller memory on the device side of `saved`; keep the host copy in the handle instead
    hdr_host.B = B;
    hdr_host.H = H;
    hdr_host.W = W;
    hdr_host.is_logits = is_logits;
    hdr_host.coef = strength / static_cast<float>(Hm * (W - 1) + (Hm - 1) * W);
    hdr_host.wave = wave;
    hdr_saved = saved;
    if (saved_hdrs.size() > 256) saved_hdrs.clear();  // blocks that were never differentiated (inference-only callers)
    saved_hdrs[saved] = hdr_host;
    if (is_logits) CK(wave_softmax(in, v.wave, v.rowsum, static_cast<long long>(rows), T_pad, T, bf16, s));
    else CK(wave_rowsum(in, v.rowsum, static_cast<long long>(rows), T_pad, T, bf16, s));
    CUDA_CK(cudaMemsetAsync(v.rowdot, 0, rows * sizeof(acc_t), s));
    if (band) {
        const size_t wb_ = static_cast<size_t>(W) * T_pad * 2, sb_ = static_cast<size_t>(W) * 4;
        uint8_t* wv = v.wave;
        uint8_t* rs = reinterpret_cast<uint8_t*>(v.rowsum);
        CK(comm->halo(wv - wb_, wv, wv + static_cast<size_t>(H - 1) * wb_, wv + static_cast<size_t>(H) * wb_, wb_, s));
        CK(comm->halo(rs - sb_, rs, rs + static_cast<size_t>(H - 1) * sb_, rs + static_cast<size_t>(H) * sb_, sb_, s));
    }
    if (use_sparse) {
        if (band) {
            set_error("wfc: the sparse contraction does not serve row-band handles");
            return -3;
        }
        CK(sparse.contract(wave, v.a, v.rowdot, B, H, W, bf16, s));
        CK(wfc_loss_finish(v.rowsum, v.rowdot, loss, B, H, W, hdr_host.coef, s, y_off, Hm));
        return 0;
    }
    if (op_wave != wave || op_a != v.a || op_B != B || op_H != H || op_W != W) {
        memset(&op, 0, sizeof(op));
        TapGemmParams& p = op.p;
        p.BN = WFC_BN;
        p.n_tiles = (T_pad + WFC_BN - 1) / WFC_BN;
        p.z_count = 1;
        {
            // N tiles swept per wave of CTAs before the next M tiles (L2 reuse of the wave tiles). All of them: measured on
            // ice 64x64 batch 8 (ncu, DRAM read of the launch / time): band of 2: 2255 MB / 0.763 ms, 4: 1767 / 0.766,
            // 7: 1014 / 0.833, all 13: 1097 / 0.772 - same time, 0.6x the DRAM reads of the former default (4). An L2
            // evict_last hint on the adjacency boxes changed neither (1051 vs 1088 MB): what is re-read is the wave.
            // The persistent CTAs (pairs) stride through the tile ids by the grid size: the band width must be coprime to
            // that stride, or a CTA would only ever meet some of the N tiles (the heavy or the light ones of the skip list).
            const char* e = getenv("WFD_WFC_NGROUP");
            int ng = p.n_tiles;
            const int stride = num_sms() / 2;
            auto gcd = [](int a, int b) {
                while (b) {
                    const int t = a % b;
                    a = b;
                    b = t;
                }
                return a;
            };
            while (ng > 1 && gcd(ng, stride) != 1) --ng;
            p.n_group = e ? atoi(e) : ng;
        }
        p.w_box = wb;
        p.h_box = hb;
        p.tiles_x = W / wb;
        p.tiles_y = H / hb;
        p.m_tiles = B * p.tiles_x * p.tiles_y;
        p.W = W;
        p.H = H;
        p.in_stride = 1;
        p.n_taps = 4;
        p.cpt = T_pad / TG_KC;
        p.tap_dx[0] = 1;  p.tap_dy[0] = 0;
        p.tap_dx[1] = -1; p.tap_dy[1] = 0;
        p.tap_dx[2] = 0;  p.tap_dy[2] = 1;
        p.tap_dx[3] = 0;  p.tap_dy[3] = -1;
        p.N = T_pad;
        p.out = v.a;
        p.out_fp32 = 0;
        p.out_sx = T_pad;
        p.out_sy = static_cast<long long>(T_pad) * W;
        p.out_sb = static_cast<long long>(T_pad) * W * H;
        p.alpha = 1.f;
        p.rowdot = v.rowdot;
        p.dotsrc = wave;
        p.ld_dot = T_pad;
        p.is_bf16 = bf16;
        p.a_ptr = wave;
        p.a_C = T_pad;
        p.a_W = W;
        p.a_H = H;
        if (band) {  // the view starts at the top halo row
            p.a_ptr = static_cast<const uint8_t*>(wave) - static_cast<size_t>(W) * T_pad * 2;
            p.a_H = H + 2;
            p.a_y_off = 1;
        }
        p.a_B = B;
        p.a_ld = T_pad;
        p.b_ptr = Bm;
        p.ldb = 4LL * T_pad;
        p.b_rows = p.n_tiles * WFC_BN;
        p.b_Z = 1;
        p.kc_list = kc_list;
        p.kc_cnt = kc_cnt;
        CK(tapgemm_prepare(&op));
        snprintf(op.tag, sizeof(op.tag), "wfc");
        op_wave = wave;
        op_a = v.a;
        op_B = B;
        op_H = H;
        op_W = W;
    }
    {
        // WFD_WFC_DBG=1: per-role wait counters of CTA 0 for the contraction (clock cycles), printed once per process
        static int dbg_mode = -1;
        if (dbg_mode < 0) {
            const char* e = getenv("WFD_WFC_DBG");
            dbg_mode = (e && e[0] == '1') ? 1 : 0;
        }
        if (dbg_mode == 1) {
            long long* d = nullptr;
            CUDA_CK(cudaMalloc(&d, 16 * sizeof(long long)));
            CUDA_CK(cudaMemset(d, 0, 16 * sizeof(long long)));
            TapGemmOp o2 = op;
            o2.p.dbg = d;
            CK(tapgemm_launch(&o2, s));  // a warm-up launch, then the measured one
            CK(tapgemm_launch(&o2, s));
            CUDA_CK(cudaStreamSynchronize(s));
            long long h[16];
            CUDA_CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
            cudaFree(d);
            fprintf(stderr, "wfc dbg: A total %lld wait-empty %lld issue %lld | B total %lld wait-empty %lld issue %lld | "
                            "MMA total %lld wait-tempty %lld wait-full %lld | EPI total %lld wait-tfull %lld busy %lld\n",
                    h[0], h[1], h[10], h[8], h[9], h[11], h[2], h[3], h[4], h[5], h[6], h[7]);
            dbg_mode = 2;
        }
    }
    CK(debug_ref_gemm() ? tapgemm_launch_ref(&op, s) : tapgemm_launch(&op, s));
    CK(wfc_loss_finish(v.rowsum, v.rowdot, loss, B, H, W, hdr_host.coef, s, y_off, Hm));
    if (band) CK(comm->allreduce_f32(loss, 1, s));  // every band holds the map's loss
    return 0;
}

int Wfc::backward(void* saved, const float* dloss, void* dout, cudaStream_t s) {
    auto it = saved_hdrs.find(saved);
    if (it == saved_hdrs.end() || it->second.B == 0) {
        set_error("wfc: backward called with a `saved` block that no forward of this handle has filled");
        return -3;
    }
    const WfcSavedHeader hdr_host = it->second;  // shapes of the forward that filled THIS block
    const size_t rows = static_cast<size_t>(hdr_host.B) * hdr_host.H * hdr_host.W;
    const size_t halo_wave = band ? align256(static_cast<size_t>(hdr_host.W) * T_pad * 2) : 0;
    const size_t halo_sum = band ? align256(static_cast<size_t>(hdr_host.W) * 4) : 0;
    SavedView v = view_saved(saved, rows, T_pad, halo_wave, halo_sum);
    return wfc_grad_finish(hdr_host.wave, v.a, v.rowsum, v.rowdot, dloss, dout, hdr_host.B, hdr_host.H, hdr_host.W,
                           T_pad, T, hdr_host.coef, hdr_host.is_logits, bf16, s, band ? band_r * hdr_host.H : 0,
                           band ? H_total : hdr_host.H);
}

}  // namespace wfd
