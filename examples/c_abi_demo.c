/*
 * c_abi_demo.c - libwfd_b200.so driven from plain C (no Python, no torch): builds a small tile-VAE decoder with
 * pseudo-random weights, a random adjacency rule set, runs two whole WFC-guidance steps through wfd_guidance_step and
 * checks two size-independent properties of the result (loss linear in the strength, finite gradient).
 *
 *   gcc -std=c99 -O2 examples/c_abi_demo.c -Iinclude -I/usr/local/cuda/include \
 *       -Lwavefunctiondiffusion_b200 -lwfd_b200 -L/usr/local/cuda/lib64 -lcudart -lm -o /tmp/c_abi_demo
 *   LD_LIBRARY_PATH=wavefunctiondiffusion_b200:/usr/local/cuda/lib64 /tmp/c_abi_demo
 *
 * The weight names are the reference's state_dict keys (wfd/wf_diffusers/wf_vae.py:151-232, resnet.py:369-499,
 * attention.py:247-380): a reference-side binding hands over `model.state_dict()` entry by entry in the same way.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wfd_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ < 0) {                                                           \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, wfd_last_error()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)
#define CUDA(call)                                                               \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));          \
            return 1;                                                            \
        }                                                                        \
    } while (0)

static unsigned long long g_state = 0x9E3779B97F4A7C15ull;
static float frand(void) { /* uniform in [-1, 1) */
    g_state = g_state * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((g_state >> 40) & 0xFFFFFF) / 8388608.0f - 1.0f;
}

static int set_weight(wfd_decoder* d, const char* key, long long numel, float scale, float offset) {
    float* w = (float*)malloc((size_t)numel * sizeof(float));
    if (!w) return -1;
    for (long long i = 0; i < numel; ++i) w[i] = offset + scale * frand();
    int rc = wfd_decoder_set_weight(d, key, w, numel);
    free(w);
    return rc;
}

static int set_conv(wfd_decoder* d, const char* prefix, int cout, int cin_per_group, int k) {
    char key[192];
    const long long fan_in = (long long)cin_per_group * k * k;
    snprintf(key, sizeof(key), "%s.weight", prefix);
    if (set_weight(d, key, (long long)cout * fan_in, 1.0f / sqrtf((float)fan_in), 0.f) < 0) return -1;
    snprintf(key, sizeof(key), "%s.bias", prefix);
    return set_weight(d, key, cout, 0.05f, 0.f);
}

static int set_norm(wfd_decoder* d, const char* prefix, int c) {
    char key[192];
    snprintf(key, sizeof(key), "%s.weight", prefix);
    if (set_weight(d, key, c, 0.1f, 1.f) < 0) return -1;
    snprintf(key, sizeof(key), "%s.bias", prefix);
    return set_weight(d, key, c, 0.1f, 0.f);
}

static int set_resnet(wfd_decoder* d, const char* prefix, int cin, int cout, int groups) {
    char key[192];
    snprintf(key, sizeof(key), "%s.norm1", prefix);
    if (set_norm(d, key, cin) < 0) return -1;
    snprintf(key, sizeof(key), "%s.conv1", prefix);
    if (set_conv(d, key, cout, cin / groups, 3) < 0) return -1;
    snprintf(key, sizeof(key), "%s.norm2", prefix);
    if (set_norm(d, key, cout) < 0) return -1;
    snprintf(key, sizeof(key), "%s.conv2", prefix);
    if (set_conv(d, key, cout, cout / groups, 3) < 0) return -1;
    if (cin != cout) {
        snprintf(key, sizeof(key), "%s.conv_shortcut", prefix);
        if (set_conv(d, key, cout, cin / groups, 1) < 0) return -1;
    }
    return 0;
}

int main(void) {
    enum { NB = 4, T = 100, T_PAD = 128, H = 16, W = 16, B = 2 };
    wfd_decoder_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int boc[NB] = {512, 256, 128, 64}, grp[NB] = {8, 4, 2, 1};
    cfg.latent_channels = 4;
    cfg.out_channels = T_PAD;
    cfg.n_blocks = NB;
    for (int i = 0; i < NB; ++i) {
        cfg.block_out_channels[i] = boc[i];
        cfg.conv_groups[i] = grp[i];
        cfg.block_down[i] = 0;
    }
    cfg.conv_init_groups = 16;
    cfg.layers_per_block = 2;
    cfg.norm_num_groups = 8;

    printf("libwfd_b200 ABI version %d\n", wfd_version());
    wfd_decoder* dec = NULL;
    CHECK(wfd_decoder_create(&cfg, H, W, B, /*dtype16=bf16*/ 1, &dec));

    /* weights, by the reference's state_dict keys */
    const int Cm = boc[NB - 1];
    char key[192];
    CHECK(set_conv(dec, "post_quant_conv", 4, 4, 1));
    CHECK(set_conv(dec, "decoder.conv_in", Cm, 4, 3));
    CHECK(set_resnet(dec, "decoder.mid_block.resnets.0", Cm, Cm, 1));
    CHECK(set_resnet(dec, "decoder.mid_block.resnets.1", Cm, Cm, 1));
    CHECK(set_norm(dec, "decoder.mid_block.attentions.0.group_norm", Cm));
    CHECK(set_conv(dec, "decoder.mid_block.attentions.0.query", Cm, Cm, 1));
    CHECK(set_conv(dec, "decoder.mid_block.attentions.0.key", Cm, Cm, 1));
    CHECK(set_conv(dec, "decoder.mid_block.attentions.0.value", Cm, Cm, 1));
    CHECK(set_conv(dec, "decoder.mid_block.attentions.0.proj_attn", Cm, Cm, 1));
    int ch = Cm;
    for (int i = 0; i < NB; ++i) {
        const int cout = boc[NB - 1 - i], g = grp[NB - 1 - i];
        for (int j = 0; j < cfg.layers_per_block + 1; ++j) {
            snprintf(key, sizeof(key), "decoder.up_blocks.%d.resnets.%d", i, j);
            CHECK(set_resnet(dec, key, j == 0 ? ch : cout, cout, g));
        }
        ch = cout;
    }
    CHECK(set_norm(dec, "decoder.conv_norm_out", ch));
    CHECK(set_conv(dec, "decoder.conv_out", T_PAD, ch / cfg.conv_init_groups, 3));
    CHECK(wfd_decoder_finalize_weights(dec));

    size_t ws_bytes = wfd_decoder_workspace_bytes(dec);
    void* ws = NULL;
    CUDA(cudaMalloc(&ws, ws_bytes));
    CHECK(wfd_decoder_bind_workspace(dec, ws, ws_bytes));
    int Ho = 0, Wo = 0;
    CHECK(wfd_decoder_out_hw(dec, &Ho, &Wo));
    printf("decoder: %zu MB workspace, map %dx%d, %d tile channels\n", ws_bytes >> 20, Ho, Wo, T_PAD);

    /* adjacency rules: tile j allowed right of / below tile i with probability 0.1 */
    uint8_t* mx = (uint8_t*)malloc(T * T);
    uint8_t* my = (uint8_t*)malloc(T * T);
    for (int i = 0; i < T * T; ++i) {
        mx[i] = frand() > 0.8f;
        my[i] = frand() > 0.8f;
    }
    wfd_wfc* wfc = NULL;
    CHECK(wfd_wfc_create(mx, my, T, T_PAD, 1, &wfc));
    void* saved = NULL;
    CUDA(cudaMalloc(&saved, wfd_wfc_saved_bytes(wfc, B, Ho, Wo)));

    /* latents (fp32 NCHW), loss and gradient buffers */
    const size_t nz = (size_t)B * 4 * H * W;
    float* hz = (float*)malloc(nz * sizeof(float));
    for (size_t i = 0; i < nz; ++i) hz[i] = frand();
    float *dz_in = NULL, *dloss = NULL, *dgrad = NULL;
    CUDA(cudaMalloc((void**)&dz_in, nz * sizeof(float)));
    CUDA(cudaMalloc((void**)&dloss, B * sizeof(float)));
    CUDA(cudaMalloc((void**)&dgrad, nz * sizeof(float)));
    CUDA(cudaMemcpy(dz_in, hz, nz * sizeof(float), cudaMemcpyHostToDevice));

    float loss5[B], loss10[B];
    float* hgrad = (float*)malloc(nz * sizeof(float));
    cudaStream_t stream;
    CUDA(cudaStreamCreate(&stream));
    CHECK(wfd_guidance_step(dec, wfc, dz_in, /*fp32*/ 0, B, 1.0f / 0.18215f, 5.0f, saved, dloss, dgrad, stream));
    CUDA(cudaStreamSynchronize(stream));
    CUDA(cudaMemcpy(loss5, dloss, sizeof(loss5), cudaMemcpyDeviceToHost));
    CHECK(wfd_guidance_step(dec, wfc, dz_in, 0, B, 1.0f / 0.18215f, 10.0f, saved, dloss, dgrad, stream));
    CUDA(cudaStreamSynchronize(stream));
    CUDA(cudaMemcpy(loss10, dloss, sizeof(loss10), cudaMemcpyDeviceToHost));
    CUDA(cudaMemcpy(hgrad, dgrad, nz * sizeof(float), cudaMemcpyDeviceToHost));

    double gnorm = 0.0;
    int finite = 1;
    for (size_t i = 0; i < nz; ++i) {
        if (!isfinite(hgrad[i])) finite = 0;
        gnorm += (double)hgrad[i] * hgrad[i];
    }
    printf("loss (strength 5): %.6f %.6f   loss (strength 10): %.6f %.6f   |d loss / d z| = %.4e   kernels launched: %lld\n",
           loss5[0], loss5[1], loss10[0], loss10[1], sqrt(gnorm), wfd_launch_count());
    int ok = finite && gnorm > 0.0;
    for (int b = 0; b < B; ++b) ok = ok && loss5[b] > 0.f && fabsf(loss10[b] / loss5[b] - 2.f) < 1e-4f;
    printf(ok ? "OK\n" : "FAILED\n");

    wfd_wfc_destroy(wfc);
    wfd_decoder_destroy(dec);
    cudaFree(ws); cudaFree(saved); cudaFree(dz_in); cudaFree(dloss); cudaFree(dgrad);
    free(mx); free(my); free(hz); free(hgrad);
    return ok ? 0 : 2;
}
