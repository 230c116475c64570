/*
 * wfd_b200.h — C ABI of libwfd_b200.so, the B200 (sm_100a) implementation of WaveFunctionDiffusion's WFC-guidance
 * hot path. Plain pointers and sizes only; no torch types. The reference has no FFI of its own: these entry points are
 * what a ctypes binding placed behind the reference's Python objects calls. Each block cites the
 * reference code it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; wfd_last_error() gives the message (thread local)
 *   - all tensor arguments are device pointers owned by the caller unless stated otherwise; nothing is freed or
 *     retained past the call except inside an explicit handle
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises the device
 *   - internal activations are channels-last [B, H*W, C] 16-bit ("nhwc16"); dtype16: 0 = fp16, 1 = bf16
 *   - handles are not thread safe individually; distinct handles may be used from distinct threads/streams
 */
#ifndef WFD_B200_H
#define WFD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFD_ABI_VERSION 1

#if defined(__GNUC__)
#define WFD_API __attribute__((visibility("default")))
#else
#define WFD_API
#endif

WFD_API int wfd_version(void);
WFD_API const char* wfd_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py reports it as gpu_launches) */
WFD_API long long wfd_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Tile-VAE decoder: AutoencoderTile._decode = post_quant_conv + DecoderTile.forward
 *   wfd/wf_diffusers/wf_vae.py:151-232 (DecoderTile), :592-599 (_decode), :617-627 (decode)
 *   wfd/wf_diffusers/resnet.py:369-499 (ResnetBlock2D), :147-192 (Downsample2D)
 *   wfd/wf_diffusers/attention.py:247-380 (AttentionBlock), wf_unet_2d_blocks.py:388-464, :1771-1883
 *   configs/tilenet/tilenet_sc_space_{64x64,32x32}.json
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct wfd_decoder wfd_decoder;

typedef struct {
    int latent_channels;        /* 4 */
    int out_channels;           /* T_pad = round_up(shrink_count, 128), train_tile_vae.py:141-144 */
    int n_blocks;               /* len(block_out_channels) */
    int block_out_channels[8];  /* config order, e.g. 3072,1536,768,384 (reversed inside, wf_vae.py:185) */
    int conv_groups[8];         /* config order, e.g. 64,16,4,1 (reversed inside, wf_vae.py:206) */
    int block_down[8];          /* decoder order: 1 where up_block_types[i] == "DownDecoderBlock2D" */
    int conv_init_groups;       /* groups of conv_out, 128 */
    int layers_per_block;       /* config value (decoder uses +1, wf_vae.py:195) */
    int norm_num_groups;        /* 8 */
} wfd_decoder_config;

/* h, w: latent plane; B_max: largest batch one call will carry; dtype16: operand/activation type */
WFD_API int wfd_decoder_create(const wfd_decoder_config* cfg, int h, int w, int B_max, int dtype16, wfd_decoder** out);
WFD_API void wfd_decoder_destroy(wfd_decoder* d);
/* output plane (H, W) = (h, w) or halved per DownDecoderBlock2D */
WFD_API int wfd_decoder_out_hw(const wfd_decoder* d, int* H, int* W);

/* Weights by state_dict key (e.g. "decoder.up_blocks.1.resnets.0.conv1.weight"), fp32, HOST memory, `numel` floats.
 * Keys not on the decode path (encoder.*, quant_conv.*) are accepted and ignored (returns 1). */
WFD_API int wfd_decoder_set_weight(wfd_decoder* d, const char* key, const float* host_data, long long numel);
/* Packs weights for the tensor-core kernels (forward and backward-data forms) and uploads them. */
WFD_API int wfd_decoder_finalize_weights(wfd_decoder* d);

/* Activation / scratch memory is caller owned: query, allocate (256-byte aligned), bind. */
WFD_API size_t wfd_decoder_workspace_bytes(const wfd_decoder* d);
WFD_API int wfd_decoder_bind_workspace(wfd_decoder* d, void* ws, size_t bytes);

/* z: NCHW [B, 4, h, w], z_dtype 0 = fp32, 1 = fp16. Computes logits = decode(z * in_scale) into the plan's internal
 * nhwc16 logits buffer and keeps what backward needs. */
WFD_API int wfd_decoder_forward(wfd_decoder* d, const void* z, int z_dtype, int B, float in_scale, void* stream);
/* device pointer to the internal logits [B, H*W, T_pad] nhwc16 */
WFD_API void* wfd_decoder_logits(wfd_decoder* d);
/* copies logits out as NCHW-shaped data: layout 0 = NCHW contiguous, 1 = channels-last (NHWC memory);
 * out_dtype 0 = fp32, 1 = fp16, 2 = bf16 */
WFD_API int wfd_decoder_export_logits(wfd_decoder* d, int B, void* out, int out_dtype, int layout, void* stream);
/* data gradient only (autograd.grad w.r.t. latents, pipeline_wfd.py:613): dlogits nhwc16 [B, H*W, T_pad] ->
 * dz fp32 NCHW [B, 4, h, w], already multiplied by in_scale (and by scale_b[b] if given). dlogits == NULL uses the
 * internal dlogits buffer (wfd_decoder_dlogits). */
WFD_API int wfd_decoder_backward(wfd_decoder* d, const void* dlogits, int B, const float* scale_b, float* dz, void* stream);
WFD_API void* wfd_decoder_dlogits(wfd_decoder* d);
/* imports a gradient given in NCHW-shaped form (layout/dtype as in export) into the internal dlogits buffer */
WFD_API int wfd_decoder_import_dlogits(wfd_decoder* d, int B, const void* in, int in_dtype, int layout, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Tile-VAE encoder: AutoencoderTile.encode = quant_conv(EncoderTile(x))   (forward only)
 *   wfd/wf_diffusers/wf_vae.py:68-148 (EncoderTile), :582-590 (encode); img2img wave input
 *   wfd/wf_diffusers/pipeline_wfd_img2img.py:68-74 (preprocess_wave), :478-486 (prepare_latents)
 * cfg: the same struct as the decoder (block_out_channels / conv_groups in config order; out_channels = the wave's channel
 * count T_pad). Only configurations whose encoder blocks keep the resolution ("EvenEncoderBlock2D": block_down all 0).
 * Weights by state_dict key: encoder.* and quant_conv.* (others are ignored, returns 1).
 * moments: fp32 NCHW [B, 2 * latent_channels, H, W] = the `parameters` of the reference's DiagonalGaussianDistribution.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct wfd_encoder wfd_encoder;
WFD_API int wfd_encoder_create(const wfd_decoder_config* cfg, int H, int W, int B_max, int dtype16, wfd_encoder** out);
WFD_API void wfd_encoder_destroy(wfd_encoder* e);
WFD_API int wfd_encoder_set_weight(wfd_encoder* e, const char* key, const float* host_data, long long numel);
WFD_API int wfd_encoder_finalize_weights(wfd_encoder* e);
WFD_API size_t wfd_encoder_workspace_bytes(const wfd_encoder* e);
WFD_API int wfd_encoder_bind_workspace(wfd_encoder* e, void* ws, size_t bytes);
/* wave: NCHW [B, T_pad, H, W], dtype 0 = fp32, 1 = fp16 (any values: one-hot, soft) */
WFD_API int wfd_encoder_forward_wave(wfd_encoder* e, const void* wave, int dtype, int B, float* moments, void* stream);
/* tiles: int64 [B, H, W] = the one-hot wave given by its indices (preprocess_wave): conv_in becomes a gather */
WFD_API int wfd_encoder_forward_tiles(wfd_encoder* e, const long long* tiles, int B, float* moments, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * WFC transition-rule loss: WFCLossEinsum
 *   wfd/wfdf/losses.py:43-75; buffers built from wfc_mats of wfd/scmap/tile_data/wfc/*.npz (pipeline_wfd.py:180-182)
 *   softmax over all T_pad channels: pipeline_wfd.py:405
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct wfd_wfc wfd_wfc;

/* mat_x[i*T+j] != 0 <=> tile j allowed to the right of i; mat_y[i*T+j] != 0 <=> j allowed below i (HOST memory) */
WFD_API int wfd_wfc_create(const uint8_t* mat_x, const uint8_t* mat_y, int T, int T_pad, int dtype16, wfd_wfc** out);
WFD_API void wfd_wfc_destroy(wfd_wfc* h);
/* bytes of the `saved` block one forward needs for (B, H, W) */
WFD_API size_t wfd_wfc_saved_bytes(const wfd_wfc* h, int B, int H, int W);
/* in: nhwc16 [B, H*W, T_pad]; is_logits != 0: softmax is applied first (decode_latents_tile), else `in` is the wave.
 * loss[b] = strength * WFCLossEinsum(wave)[b]. `in` must stay valid until backward when is_logits == 0. */
WFD_API int wfd_wfc_forward(wfd_wfc* h, const void* in, int is_logits, int B, int H, int W, float strength, void* saved,
                    float* loss, void* stream);
/* dout nhwc16 [B, H*W, T_pad]: gradient of sum_b dloss[b]*loss[b] w.r.t. `in` of the matching forward
 * (dloss == NULL means ones). */
WFD_API int wfd_wfc_backward(wfd_wfc* h, void* saved, const float* dloss, void* dout, void* stream);
/* device pointer to the wave [B, H*W, T_pad] nhwc16 held in `saved` after a forward with is_logits != 0 */
WFD_API void* wfd_wfc_saved_wave(const wfd_wfc* h, void* saved, int B, int H, int W);

/* ------------------------------------------------------------------------------------------------------------------
 * One whole guidance step: decode + softmax + loss + backward to the latents
 *   pipeline_wfd.py:609-613 minus scheduler.step (its scalar chain factor stays with the caller)
 * loss[b] = strength * loss_b ; dz = d(sum_b loss[b]) / d z   (fp32 NCHW, includes in_scale)
 * ---------------------------------------------------------------------------------------------------------------- */
WFD_API int wfd_guidance_step(wfd_decoder* d, wfd_wfc* h, const void* z, int z_dtype, int B, float in_scale, float strength,
                      void* wfc_saved, float* loss, float* dz, void* stream);

/* n_steps guided updates of the latents for a scheduler whose step is affine in the sample, prev = a x + b eps (DDIM
 * eta = 0, PNDM/PLMS): the whole body of the reference's guidance block, pipeline_wfd.py:605-615, and with n_steps = 20
 * its final-stage loop :626-639 (same eps and t every iteration), without leaving the library:
 *     eps  = eps_uncond + cfg_scale * (eps_text - eps_uncond)        (:601-603; eps_text == NULL: eps = eps_uncond)
 *     z    = a * lat + b * eps                                       (:611 scheduler.step(...).prev_sample)
 *     loss = strength * wfc_loss(softmax(decode(z * in_scale)))       (:612)
 *     lat -= a * d(sum_b loss_b)/dz                                   (:613-614; a = d prev / d sample)
 * lat: fp32 NCHW [B,4,h,w], updated in place; eps_*: fp32, same shape; loss_out: [n_steps][B]; z_scratch / dz_scratch:
 * fp32 buffers of lat's size. A fixed launch sequence: capture it once, replay it per call. */
WFD_API int wfd_guidance_loop(wfd_decoder* d, wfd_wfc* h, float* lat, const float* eps_uncond, const float* eps_text,
                              float cfg_scale, float a, float b, int B, float in_scale, float strength, int n_steps,
                              void* wfc_saved, float* loss_out, float* z_scratch, float* dz_scratch, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Row bands: ONE large map (BASELINE config 4, 256x256) split along y over R GPUs, one band of h/R rows per rank.
 * The reference runs such a map on one device (pipeline_wfd.py:283-290 only widens width/height); cutting the plane
 * turns its single-device ops into exchanges between neighbouring bands: a one-row halo for every 3x3 conv
 * (resnet.py:416,432, wf_vae.py:218,232) and for the vertical WFC pairs (losses.py:60-66), an all-reduce of the
 * GroupNorm sums (resnet.py:407,423) and of the loss (losses.py:69-75), an all-gather of the mid-block keys/values and
 * a reduce-scatter of their gradients (attention.py:329-380).
 *
 * wfd_comm: the exchange transport. NCCL (one process per GPU; `libnccl_path` names the libnccl the process already
 * uses, e.g. torch's bundled one; the 128-byte id comes from rank 0 via wfd_comm_unique_id and travels to the other
 * ranks by whatever channel the host has) or an in-process loopback group (R bands on one GPU, one host thread per
 * band; for tests).
 * A band decoder takes the WHOLE latent plane z [1,4,h,w] in wfd_decoder_forward / wfd_guidance_step and produces the
 * logits / dz of its own rows only: logits [h/R * w, T_pad], dz fp32 [1,4,h/R,w]. The loss is the map's on every rank.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct wfd_comm wfd_comm;
typedef struct wfd_loop_group wfd_loop_group;
WFD_API int wfd_comm_unique_id(const char* libnccl_path, void* id128);
WFD_API int wfd_comm_create_nccl(const char* libnccl_path, const void* id128, int rank, int nranks, wfd_comm** out);
WFD_API wfd_loop_group* wfd_loop_group_create(int nranks);
WFD_API void wfd_loop_group_destroy(wfd_loop_group* g);
WFD_API int wfd_comm_create_loopback(wfd_loop_group* g, int rank, wfd_comm** out);
WFD_API void wfd_comm_destroy(wfd_comm* c);
/* Peer mailboxes (NCCL communicators, one process per GPU on one NVLink / NVSwitch node): the small exchanges of a band
 * step - halo rows, GroupNorm sums, the loss - become direct stores into the neighbour's memory plus a flag (one short
 * kernel each, replayable from a CUDA graph) instead of ncclSend / Recv / AllReduce; the attention all-gather and
 * reduce-scatter stay on NCCL. wfd_comm_peer_handle allocates this rank's mailbox (rows of up to max_row_bytes) and
 * returns its 64-byte CUDA IPC handle; the caller gathers the handles of all ranks (rank order) and hands them to
 * wfd_comm_peer_attach on every rank. Exchanges that do not fit the mailbox keep using NCCL. */
WFD_API int wfd_comm_peer_handle(wfd_comm* c, size_t max_row_bytes, void* handle64);
WFD_API int wfd_comm_peer_attach(wfd_comm* c, const void* handles);
/* h is the height of the whole latent plane; batch is 1 */
WFD_API int wfd_decoder_create_band(const wfd_decoder_config* cfg, int h, int w, int rank, int nranks, int dtype16,
                                    wfd_decoder** out);
WFD_API int wfd_decoder_attach_comm(wfd_decoder* d, wfd_comm* c);
/* after this, wfd_wfc_forward takes the band's logits (B = 1, H = H_total / nranks) */
WFD_API int wfd_wfc_attach_comm(wfd_wfc* h, wfd_comm* c, int H_total);

/* ------------------------------------------------------------------------------------------------------------------
 * Wave argmax into tiles: numpy.argmax(wave, axis=-1) (wfd/webui/ui.py:228, README.md:90-91)
 * x: [rows, n] contiguous, dtype 0 = fp32, 1 = fp16, 2 = bf16; out: int64 [rows]. First maximum wins, NaN is maximal.
 * ---------------------------------------------------------------------------------------------------------------- */
WFD_API int wfd_argmax_tiles(const void* x, int dtype, long long rows, int n, long long* out, void* stream);

/* Rule violations of integer tile maps (what the reference's users check after the argmax; it is WFCLossEinsum of the
 * one-hot wave, losses.py:63-75, times the pair count): tiles int64 [B, H, W] on the device, out int64 [B, 2] =
 * {horizontal pairs with mat_x[left, right] == 0, vertical pairs with mat_y[top, bottom] == 0}; ids outside [0, T)
 * (padded channels) always violate. */
/* fraction of the [256 tiles] x [64 neighbour tiles] blocks of the packed adjacency operand that hold an allowed pair and
 * are therefore visited by the contraction (the skipped ones add exact zeros); bench.py reports executed flops with it */
WFD_API double wfd_wfc_kc_kept(const wfd_wfc* h);
WFD_API int wfd_wfc_count_violations(const wfd_wfc* h, const long long* tiles, int B, int H, int W, long long* out,
                                     void* stream);

/* Alternative contraction for the loss: with on != 0 the handle computes a[n, i] through a CUDA-core gather
 * over the allowed pairs of the adjacency matrices (0.2 % dense for the StarCraft tilesets; sliced-ELL lists, four cells
 * per CTA) instead of the dense tensor-core product. Same `saved` layout, same finishing kernels, same gradient; fp32
 * accumulation in a different order, so results agree to rounding, not bitwise. Off by default: north_star asks for the
 * tensor-core contraction, bench.py reports this one separately. Not available on row-band handles. */
WFD_API int wfd_wfc_set_sparse(wfd_wfc* h, int on);

/* ------------------------------------------------------------------------------------------------------------------
 * Wave-seeded constraint propagation: the information propagation of the reference's prior-distribution
 * WFC solver (wfd/wfc/wfc_prior_dist.py:136-174 propagate_information, :176-199 double_check) as bit-set arc
 * consistency over whole maps, on the device.
 *   allowed  uint32 [B, H, W, Wd], Wd = ceil(T / 32): bit (t & 31) of word (t >> 5) = tile t still possible in the cell
 *            (map_allowed_tiles, wfc_prior_dist.py:14); bits >= T must be zero
 *   conn     uint32 [4, T, Wd]: conn[d][s] = tiles allowed at direction d of tile s (connections[s, d, :], :11), d ordered
 *            as the reference's directions (:25): 0 left, 1 up, 2 right, 3 down
 *   boundary uint8 [B, H, W] or NULL: non-zero cells take no part (boundary_grids / is_grid_allowed, :41-43)
 * wfd_propagator_create takes conn from HOST memory and re-encodes it for the device as supporter lists (for direction
 * d and target tile t the source tiles s with conn[d][s][t] set; the rules are 0.2 % dense, so ~7 16-bit ids per list).
 * wfd_propagator_run repeats   allowed'[n] = allowed[n] & AND_d OR_{s in allowed[m_d]} conn[d][s]   over every cell n and its
 * in-map sources m_d = n - direction d (an empty source does not propagate, :152-153), every sweep from the previous
 * sweep's state, until a sweep changes nothing or max_sweeps ran. Order-free and deterministic; while no cell runs empty
 * it ends in the state the reference reaches by repeating its raster-order depth-1 propagation until nothing changes
 * (the update is monotone). The handle owns its device scratch (flags + a second state buffer, grown on demand).
 * *sweeps_done / *converged are host ints (the call synchronises the stream every few sweeps to read the change
 * counters). wfd_propagator_entries: list entries after the re-encoding. wfd_wfc_allowed_count: tiles left per cell
 * (is_fixed / is_zero, :86-90), int32 [B, H, W] on the device.
 * wfd_wave_to_allowed seeds the sets from a wave: allowed[n, t] = first_tile <= t < T and (wave[n, t] >= threshold or
 * (keep_argmax and t == argmax over [first_tile, T), first maximum wins)); wave [rows, ld] with dtype 0 fp32 / 1 fp16 /
 * 2 bf16; first_tile = 1 keeps the null tile out as allow_all_connections_to_null does (:207-212).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct wfd_propagator wfd_propagator;
WFD_API int wfd_propagator_create(const uint32_t* conn_bits, int T, wfd_propagator** out);
WFD_API void wfd_propagator_destroy(wfd_propagator* p);
WFD_API long long wfd_propagator_entries(const wfd_propagator* p);
WFD_API int wfd_propagator_run(wfd_propagator* p, uint32_t* allowed, const uint8_t* boundary, int B, int H, int W,
                               int max_sweeps, int* sweeps_done, int* converged, void* stream);
WFD_API int wfd_wfc_allowed_count(const uint32_t* allowed, int B, int H, int W, int T, int* count, void* stream);
WFD_API int wfd_wave_to_allowed(const void* wave, int dtype, long long rows, int ld, int T, int first_tile,
                                float threshold, int keep_argmax, uint32_t* allowed, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Tile post-processing on the device (what wfd/webui/ui.py:228-243 does on the host after the argmax):
 *   optional mirror symmetry   tiles' = [tiles | symm[tiles[:, ::-1]]]          wfd/scmap/data_processing.py:191-196
 *   id chain                   sc = gen_to_sc[shrink_to_gen[tiles']]            wfd/scmap/map_display.py:55,
 *                                                                                 data_processing.py:206-211
 * tiles int64 [B, H, W] (device); symm (may be NULL), shrink_to_gen (may be NULL = identity), gen_to_sc: int32 tables on
 * the device; tiles_out int64 [B, H, W'] and sc_out uint16 [B, H, W'] (either may be NULL), W' = 2 W with symmetry.
 * *bad (device int64) counts ids outside a table (an argmax over padded channels): they come out as 0xFFFF, where the
 * reference's numpy indexing would raise. So only a (H, W') uint16 map has to leave the GPU, not 52 MB of wave.
 * ---------------------------------------------------------------------------------------------------------------- */
WFD_API int wfd_tiles_finish(const long long* tiles, int B, int H, int W, const int* symm, int n_symm,
                             const int* shrink_to_gen, int n_shrink, const int* gen_to_sc, int n_gen,
                             long long* tiles_out, uint16_t* sc_out, long long* bad, void* stream);

/* layout helpers used by the Python layer: NCHW (fp32/fp16) <-> nhwc16, 16-bit -> fp32 */
WFD_API int wfd_nchw_to_nhwc16(const void* in, int in_dtype, void* out, int B, int C, int HW, int dtype16, void* stream);
WFD_API int wfd_nhwc16_to_nchw(const void* in, void* out, int out_dtype, int B, int C, int HW, int dtype16, void* stream);
WFD_API int wfd_cvt16_to_f32(const void* in, float* out, long long n, int dtype16, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Measurement
 * ---------------------------------------------------------------------------------------------------------------- */
/* per-launch CUDA-event profiling (bench.py roofline): enable(1) clears and starts recording, report() synchronises the
 * recorded events and writes "name launches total_ms" lines into buf (returns the untruncated length) */
WFD_API int wfd_profile_enable(int on);
WFD_API int wfd_profile_report(char* buf, size_t cap);
#ifdef __cplusplus
}
#endif
#endif /* WFD_B200_H */
