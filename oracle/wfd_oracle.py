"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's WFC-guidance path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module, and
only as the checker or the reported CPU baseline. The product path (wavefunctiondiffusion_b200/) never imports it.

Each function restates, with plain CPU tensor arithmetic (torch used as an ndarray library: conv2d, group_norm,
matmul; any float dtype, fp64 for pinning), what the cited reference code computes, working directly from a
state_dict so that it shares no code with the product's packing or kernels.

Pinning: the reference ships no tests or golden vectors. The oracle is pinned against the UNMODIFIED
reference modules executed in the authoring container by oracle/make_golden.py (fixtures under tests/golden/) and
against analytic known-answer values of the loss (uniform wave, one-hot maps); tests/test_oracle.py checks both.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LATENT_SCALE = 0.18215  # reference pipeline_wfd.py:403


# ------------------------------------------------------------------------------------------------ decoder
def _p(sd, key, dtype):
    return sd[key].detach().to("cpu", dtype)


def resnet_block(sd, prefix, x, conv_groups, norm_groups, dtype):
    """ResnetBlock2D.forward with temb=None (reference wfd/wf_diffusers/resnet.py:458-499)."""
    h = F.group_norm(x, norm_groups, _p(sd, prefix + ".norm1.weight", dtype), _p(sd, prefix + ".norm1.bias", dtype), 1e-6)
    h = F.silu(h)
    h = F.conv2d(h, _p(sd, prefix + ".conv1.weight", dtype), _p(sd, prefix + ".conv1.bias", dtype), padding=1,
                 groups=conv_groups)
    h = F.group_norm(h, norm_groups, _p(sd, prefix + ".norm2.weight", dtype), _p(sd, prefix + ".norm2.bias", dtype), 1e-6)
    h = F.silu(h)
    h = F.conv2d(h, _p(sd, prefix + ".conv2.weight", dtype), _p(sd, prefix + ".conv2.bias", dtype), padding=1,
                 groups=conv_groups)
    if prefix + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, _p(sd, prefix + ".conv_shortcut.weight", dtype), _p(sd, prefix + ".conv_shortcut.bias", dtype),
                     groups=conv_groups)
    return x + h


def attention_block(sd, prefix, x, norm_groups, dtype):
    """AttentionBlock.forward, one head (reference wfd/wf_diffusers/attention.py:329-380)."""
    b, c, hh, ww = x.shape
    t = F.group_norm(x, norm_groups, _p(sd, prefix + ".group_norm.weight", dtype), _p(sd, prefix + ".group_norm.bias", dtype), 1e-6)
    t = t.reshape(b, c, hh * ww).transpose(1, 2)
    q = t @ _p(sd, prefix + ".query.weight", dtype).t() + _p(sd, prefix + ".query.bias", dtype)
    k = t @ _p(sd, prefix + ".key.weight", dtype).t() + _p(sd, prefix + ".key.bias", dtype)
    v = t @ _p(sd, prefix + ".value.weight", dtype).t() + _p(sd, prefix + ".value.bias", dtype)
    scores = (q @ k.transpose(1, 2)) * (1.0 / math.sqrt(c))
    probs = torch.softmax(scores.float(), dim=-1).to(scores.dtype)  # the reference softmaxes in fp32 (attention.py:367)
    o = probs @ v
    o = o @ _p(sd, prefix + ".proj_attn.weight", dtype).t() + _p(sd, prefix + ".proj_attn.bias", dtype)
    return o.transpose(1, 2).reshape(b, c, hh, ww) + x


def decode(sd, cfg, z, dtype=torch.float64):
    """AutoencoderTile._decode: post_quant_conv then DecoderTile.forward (reference wf_vae.py:592-599, 216-232).

    cfg: dict with the tilenet config fields (configs/tilenet/*.json + in/out_channels)."""
    n_res = cfg["layers_per_block"] + 1
    G = cfg["norm_num_groups"]
    rev_groups = list(reversed(cfg["conv_groups"]))
    x = z.to(dtype)
    x = F.conv2d(x, _p(sd, "post_quant_conv.weight", dtype), _p(sd, "post_quant_conv.bias", dtype))
    x = F.conv2d(x, _p(sd, "decoder.conv_in.weight", dtype), _p(sd, "decoder.conv_in.bias", dtype), padding=1)
    x = resnet_block(sd, "decoder.mid_block.resnets.0", x, 1, G, dtype)
    x = attention_block(sd, "decoder.mid_block.attentions.0", x, G, dtype)
    x = resnet_block(sd, "decoder.mid_block.resnets.1", x, 1, G, dtype)
    for i, kind in enumerate(cfg["up_block_types"]):
        for j in range(n_res):
            x = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", x, rev_groups[i], G, dtype)
        if kind == "DownDecoderBlock2D":  # Downsample2D, padding 0: pad (0,1,0,1) then stride-2 conv (resnet.py:185-190)
            key = f"decoder.up_blocks.{i}.downsamplers.0.conv"
            x = F.conv2d(F.pad(x, (0, 1, 0, 1)), _p(sd, key + ".weight", dtype), _p(sd, key + ".bias", dtype), stride=2,
                         groups=rev_groups[i])
    x = F.group_norm(x, G, _p(sd, "decoder.conv_norm_out.weight", dtype), _p(sd, "decoder.conv_norm_out.bias", dtype), 1e-6)
    x = F.silu(x)
    return F.conv2d(x, _p(sd, "decoder.conv_out.weight", dtype), _p(sd, "decoder.conv_out.bias", dtype), padding=1,
                    groups=cfg["conv_init_groups"])


def encode(sd, cfg, x, dtype=torch.float64):
    """AutoencoderTile.encode moments: quant_conv(EncoderTile.forward(x)) (reference wf_vae.py:582-584, 131-148) for the
    configurations whose encoder blocks keep the resolution. x: (B, T_pad, H, W) wave (one-hot or soft)."""
    G = cfg["norm_num_groups"]
    x = x.to(dtype)
    x = F.conv2d(x, _p(sd, "encoder.conv_in.weight", dtype), _p(sd, "encoder.conv_in.bias", dtype), padding=1,
                 groups=cfg["conv_init_groups"])
    for i, kind in enumerate(cfg["down_block_types"]):
        assert kind == "EvenEncoderBlock2D", kind
        for j in range(cfg["layers_per_block"]):
            x = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", x, cfg["conv_groups"][i], G, dtype)
    x = resnet_block(sd, "encoder.mid_block.resnets.0", x, 1, G, dtype)
    x = attention_block(sd, "encoder.mid_block.attentions.0", x, G, dtype)
    x = resnet_block(sd, "encoder.mid_block.resnets.1", x, 1, G, dtype)
    x = F.group_norm(x, G, _p(sd, "encoder.conv_norm_out.weight", dtype), _p(sd, "encoder.conv_norm_out.bias", dtype), 1e-6)
    x = F.conv2d(F.silu(x), _p(sd, "encoder.conv_out.weight", dtype), _p(sd, "encoder.conv_out.bias", dtype), padding=1)
    return F.conv2d(x, _p(sd, "quant_conv.weight", dtype), _p(sd, "quant_conv.bias", dtype))


# ------------------------------------------------------------------------------------------------ WFC loss
def wfc_loss(wave, mat_x, mat_y):
    """WFCLossEinsum.forward (reference wfd/wfdf/losses.py:63-75): wave (B, C>=T, H, W) -> (B,).

    Written as per-pair bilinear forms through one matmul per direction instead of the reference's three-operand einsum."""
    T = mat_x.shape[0]
    cx = 1.0 - torch.as_tensor(np.asarray(mat_x), dtype=wave.dtype)
    cy = 1.0 - torch.as_tensor(np.asarray(mat_y), dtype=wave.dtype)
    p = wave[:, :T].permute(0, 2, 3, 1)  # (B, H, W, T)
    B, H, W, _ = p.shape
    left, right = p[:, :, :-1], p[:, :, 1:]
    top, bottom = p[:, :-1], p[:, 1:]
    sx = ((left.reshape(-1, T) @ cx) * right.reshape(-1, T)).sum(-1).reshape(B, -1).sum(-1)
    sy = ((top.reshape(-1, T) @ cy) * bottom.reshape(-1, T)).sum(-1).reshape(B, -1).sum(-1)
    return (sx + sy) / (H * (W - 1) + (H - 1) * W)


def wfc_loss_sparse(wave, mat_x, mat_y):
    """Same value through the complement identity p^T (1-A) q = sum(p) sum(q) - p^T A q with A sparse:
    cheap enough for the largest maps; fp64 in, fp64 out."""
    import scipy.sparse as sp

    T = mat_x.shape[0]
    ax = sp.csr_matrix(np.asarray(mat_x, dtype=np.float64))
    ay = sp.csr_matrix(np.asarray(mat_y, dtype=np.float64))
    p = np.asarray(wave[:, :T].permute(0, 2, 3, 1).to(torch.float64))
    B, H, W, _ = p.shape
    S = p.sum(-1)
    out = np.zeros(B)
    for b in range(B):
        l, r = p[b, :, :-1].reshape(-1, T), p[b, :, 1:].reshape(-1, T)
        t, d = p[b, :-1].reshape(-1, T), p[b, 1:].reshape(-1, T)
        sx = (S[b, :, :-1] * S[b, :, 1:]).sum() - ((ax.T @ l.T).T * r).sum()
        sy = (S[b, :-1] * S[b, 1:]).sum() - ((ay.T @ t.T).T * d).sum()
        out[b] = (sx + sy) / (H * (W - 1) + (H - 1) * W)
    return torch.as_tensor(out)


def wfc_grad_closed_form(p, mat_x, mat_y, strength=1.0):
    """d(strength * loss_b)/d p in closed form: p (B, C>=T, H, W) -> same shape, zero beyond T."""
    T = mat_x.shape[0]
    cx = 1.0 - torch.as_tensor(np.asarray(mat_x), dtype=p.dtype)
    cy = 1.0 - torch.as_tensor(np.asarray(mat_y), dtype=p.dtype)
    q = p[:, :T].permute(0, 2, 3, 1)
    B, H, W, _ = q.shape
    g = torch.zeros_like(q)
    g[:, :, :-1] += q[:, :, 1:] @ cx.t()   # right neighbour: sum_j cx[i,j] p_right[j]
    g[:, :, 1:] += q[:, :, :-1] @ cx       # left neighbour:  sum_j cx[j,i] p_left[j]
    g[:, :-1] += q[:, 1:] @ cy.t()         # below
    g[:, 1:] += q[:, :-1] @ cy             # above
    out = torch.zeros_like(p)
    out[:, :T] = (g * (strength / (H * (W - 1) + (H - 1) * W))).permute(0, 3, 1, 2)
    return out


def uniform_wave_loss(mat_x, mat_y, t_pad):
    """Analytic loss of the all-equal wave over t_pad channels (SURVEY §4.2 KAT)."""
    T = mat_x.shape[0]
    return (2.0 * T * T - float(np.count_nonzero(mat_x)) - float(np.count_nonzero(mat_y))) / (2.0 * t_pad * t_pad)


def forbidden_pair_fraction(tiles, mat_x, mat_y):
    """Loss of the one-hot wave of an integer tile map (H, W): forbidden adjacent pairs / all adjacent pairs."""
    tiles = np.asarray(tiles)
    H, W = tiles.shape
    bad = np.count_nonzero(np.asarray(mat_x)[tiles[:, :-1], tiles[:, 1:]] == 0)
    bad += np.count_nonzero(np.asarray(mat_y)[tiles[:-1], tiles[1:]] == 0)
    return bad / (H * (W - 1) + (H - 1) * W)


# ------------------------------------------------------------------------------------------------ whole step
def guidance_step(sd, cfg, mat_x, mat_y, latents, strength, dtype=torch.float64, in_scale=1.0 / LATENT_SCALE):
    """The arithmetic of the guidance block without scheduler.step (reference pipeline_wfd.py:609-613, :402-408):
    logits = decode(latents / 0.18215); wave = softmax over all channels; loss = strength * WFCLossEinsum(wave);
    grad = d(sum_b loss_b)/d latents (sum: the reference's own autograd.grad only accepts batch 1)."""
    lat = latents.detach().to("cpu", dtype).requires_grad_(True)
    with torch.enable_grad():
        logits = decode(sd, cfg, lat * in_scale, dtype)
        wave = torch.softmax(logits, dim=1)
        loss = wfc_loss(wave, mat_x, mat_y) * strength
        (grad,) = torch.autograd.grad(loss.sum(), lat)
    return {"logits": logits.detach(), "wave": wave.detach(), "loss": loss.detach(), "grad": grad.detach()}


def argmax_tiles(wave_nhwc):
    """Caller-side argmax of the (B, H, W, T_pad) wave (reference wfd/webui/ui.py:228): numpy semantics."""
    return np.argmax(np.asarray(wave_nhwc), axis=-1)


def add_symmetry(tile_result, mapping):
    """Mirror symmetry of a tile map (reference wfd/scmap/data_processing.py:191-196): (H, W) -> (H, 2 W)."""
    tile_result = np.asarray(tile_result)
    return np.concatenate([tile_result, np.asarray(mapping)[tile_result[:, ::-1]]], axis=1)


def tiles_to_sc(tile_result, shrink_to_gen, gen_to_sc):
    """Shrink ids -> StarCraft tile ids (reference wfd/scmap/map_display.py:55, data_processing.py:206-211)."""
    return np.asarray(gen_to_sc)[np.asarray(shrink_to_gen)[np.asarray(tile_result)]]


def ddim_coefficients(num_inference_steps=50, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """a_i, b_i of prev = a_i * x + b_i * eps for DDIM(eta=0) on SD's scaled-linear betas: the affine scheduler.step
    stand-in of the synthetic 50-step loop. d prev / d x = a_i is the chain factor the guidance block
    applies on top of our gradient."""
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
    acp = np.cumprod(1.0 - betas)
    step = num_train_timesteps // num_inference_steps
    ts = (np.arange(0, num_inference_steps) * step)[::-1]
    a, b = [], []
    for t in ts:
        prev_t = t - step
        at = acp[t]
        ap = acp[prev_t] if prev_t >= 0 else 1.0
        a.append(math.sqrt(ap / at))
        b.append(math.sqrt(1 - ap) - math.sqrt(ap * (1 - at) / at))
    return np.array(a), np.array(b)
