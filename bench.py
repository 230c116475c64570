#!/usr/bin/env python
"""Benchmark of the WFC-guidance hot path (decode + softmax + WFC loss + backward to the latents).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (CPU arm)

One "step" is one guidance step for the whole per-GPU batch of synthetic latents (BASELINE config 2: ice tileset,
64x64 map, batch 8, random-init AutoencoderTile from configs/tilenet/tilenet_sc_space_64x64.json, strength 5).
`value` is map-steps per second over all GPUs with inputs resident in HBM; `e2e` is the same through the pipeline's
guidance block with host (pinned) latents and noise prediction copied in and the new latents copied out every step.
Rank 0 prints exactly one JSON line.
"""
import argparse
import ctypes as C
import json
import re
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STRENGTH = 5.0
LATENT_SCALE = 0.18215


# --------------------------------------------------------------------------------------------------- flop model
def decoder_layers(cfg, h, w):
    """(name, macs) per tensor-core layer of one decode, from the architecture (SURVEY §8d closed form). The launch class
    of a layer (the tag the library stamps on its contractions) follows from its name: see layer_class()."""
    rev = list(reversed(cfg["block_out_channels"]))
    rev_g = list(reversed(cfg["conv_groups"]))
    n_res = cfg["layers_per_block"] + 1
    cm = rev[0]
    hw = h * w
    out = [("conv_in", hw * 4 * cm * 9)]

    def resnet(name, cin, cout, g, px):
        out.append((name + ".conv1", px * cout * (cin // g) * 9))
        out.append((name + ".conv2", px * cout * (cout // g) * 9))
        if cin != cout:
            out.append((name + ".shortcut", px * cout * (cin // g)))

    resnet("mid.0", cm, cm, 1, hw)
    out.append(("attn.qkv_proj", hw * cm * cm * 4))
    out.append(("attn.bmm", 2 * hw * hw * cm))
    resnet("mid.1", cm, cm, 1, hw)
    ch, px = cm, hw
    for i, kind in enumerate(cfg["up_block_types"]):
        for j in range(n_res):
            resnet("up%d.%d" % (i, j), ch if j == 0 else rev[i], rev[i], rev_g[i], px)
        ch = rev[i]
        if kind == "DownDecoderBlock2D":
            px //= 4
            out.append(("up%d.down" % i, px * ch * (ch // rev_g[i]) * 9))
    out.append(("conv_out", px * cfg["out_channels"] * (ch // cfg["conv_init_groups"]) * 9))
    return out, px


def layer_class(cfg, name):
    """dense / g4 / g16 / g64 (resnet convs by conv group count), conv_out, attn, down - the tags of csrc/decoder.cu."""
    rev_g = list(reversed(cfg["conv_groups"]))
    if name.startswith("mid."):
        return "dense"
    if name.startswith("attn."):
        return "attn"
    if name == "conv_out":
        return "conv_out"
    if name.endswith(".down"):
        return "down"
    if name.startswith("up"):
        g = rev_g[int(name[2:name.index(".")])]
        return "dense" if g == 1 else "g%d" % g
    return "other"


def class_flops(cfg, h, w, T, t_pad, kc_kept):
    """Tensor-core flops of ONE map-step per launch class (2 flops per MAC): `algorithmic` as SURVEY §8d counts them
    (forward + backward-data of every decoder contraction, one extra attention product pair in backward, the dense
    4 T^2 P loss), and `executed` where the kernel provably does less or more: the WFC contraction visits only the K
    chunks of the packed adjacency operand that hold an allowed pair (kc_kept of them, over T_pad-padded operands)."""
    layers, px = decoder_layers(cfg, h, w)
    alg = {}
    for name, macs in layers:
        if name == "conv_in":
            continue  # 4-channel latent conv: a CUDA-core kernel (latent_in / latent_out), 0.1 GF
        c = layer_class(cfg, name)
        alg[c] = alg.get(c, 0.0) + 2 * 2.0 * macs           # forward + backward-data
    cm = list(reversed(cfg["block_out_channels"]))[0]
    alg["attn"] = alg.get("attn", 0.0) + 4.0 * (h * w) ** 2 * cm   # attention backward needs 4 products, forward 2
    side = int(round(px ** 0.5))
    P = 2 * side * (side - 1)
    alg["wfc"] = 4.0 * T * T * P
    exe = dict(alg)
    n_tiles = (t_pad + 255) // 256
    exe["wfc"] = 2.0 * px * (n_tiles * 256) * (4 * t_pad) * kc_kept
    if os.environ.get("WFD_NO_FUSE_ATTN_FWD", "0") != "1":
        # the fp32 row softmax runs inside TWO passes of the score product (row statistics, then probabilities):
        # one more N x N x C product is executed than the algorithm needs
        exe["attn"] += 2.0 * (h * w) ** 2 * cm
    return alg, exe


def step_flops(cfg, h, w, T):
    layers, px = decoder_layers(cfg, h, w)
    f_dec = 2.0 * sum(m for _, m in layers)
    f_attn = 4.0 * (h * w) ** 2 * list(reversed(cfg["block_out_channels"]))[0]
    side = int(round(px ** 0.5))
    H = W = side
    P = H * (W - 1) + (H - 1) * W
    f_loss = 4.0 * T * T * P
    return {"F_dec": f_dec, "F_attn": f_attn, "F_loss": f_loss, "F_step": 2 * f_dec + f_attn + f_loss}


# --------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------- workload
def load_workload(args):
    from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config

    mx, my = load_wfc_mats(args.tileset)
    T = mx.shape[0]
    t_pad = round_up_channels(T)
    cfg = tilenet_config(args.arch, t_pad)
    return mx, my, T, t_pad, cfg


def reference_cpu_step_fn(args, cfg, mx, my):
    """One guidance step of ONE map on the host cores, fp32: (callable, kind). kind "reference" = the UNMODIFIED
    reference modules (imported through oracle/reference_loader.py from /root/reference or the offline install under
    baseline/_ref) executing pipeline_wfd.py:609-613; kind "port" = the oracle restatement when neither is present."""
    import torch

    from oracle import wfd_oracle as orc
    from oracle.reference_loader import load_reference, reference_root

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    lat = torch.randn(1, 4, args.latent, args.latent, generator=g)
    if reference_root() is not None:
        vae, losses = load_reference()
        torch.manual_seed(0)
        net = vae.AutoencoderTile(**cfg).eval()
        lossm = losses.WFCLossEinsum(mx, my)

        def step():
            with torch.enable_grad():
                _latents = lat.detach().requires_grad_()
                wave = net.decode(1 / LATENT_SCALE * _latents).sample.softmax(axis=1)
                _loss = lossm(wave) * STRENGTH
                (grad,) = torch.autograd.grad(_loss.sum(), _latents)
            return _loss.detach(), grad

        return step, "reference"
    from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile

    torch.manual_seed(0)
    sd = AutoencoderTile(**cfg).state_dict()

    def step():
        r = orc.guidance_step(sd, cfg, mx, my, lat, STRENGTH, torch.float32)
        return r["loss"], r["grad"]

    return step, "port"


def cpu_oracle_maps_per_s(args, cfg, mx, my, repeats, maps=1):
    """The reference's CPU path on the host cores: best of `repeats` after one warm-up."""
    import torch

    step, kind = reference_cpu_step_fn(args, cfg, mx, my)
    step()
    best = float("inf")
    for _ in range(repeats):
        t0 = time.time()
        step()
        best = min(best, time.time() - t0)
    return 1.0 / best, torch.get_num_threads(), best, kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, all threads, fp32
    (set_precision is a no-op on CPU, pipeline_wfd.py:205-206). The timed code is the UNMODIFIED reference
    (baseline/_ref install or /root/reference, through the stub-diffusers loader) when present, else the oracle port.
    One step = ONE map of the batch (maps are independent and the metric is per map): the bounded sample that keeps
    `--steps K` within minutes on a host whose step costs ~1 s per map."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    mx, my, T, t_pad, cfg = load_workload(args)
    step, kind = reference_cpu_step_fn(args, cfg, mx, my)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = args.steps
    t0 = time.time()
    for _ in range(steps):
        loss, _ = step()
    dt = time.time() - t0
    v = steps / dt
    sample = ("each step = 1 map of the batch of %d (fp32, %s), %d steps, %.2f s per map-step"
              % (args.batch, "unmodified reference modules" if kind == "reference" else "oracle port", steps, dt / steps))
    print(json.dumps({
        "impl": "reference", "metric": "wfc_guidance_map_steps_per_s", "value": v, "unit": "maps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, T, t_pad, per_gpu_batch=args.batch),
        "cpu_baseline": {"value": v, "unit": "maps/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss_check": [float(x) for x in loss.reshape(-1)[:1]],
    }))


def gpu_eager_baseline(args, cfg, mx, my, B, dev, ours_loss, ours_grad, lat):
    """The reference's own eager PyTorch path on THIS B200 (what a user of the reference runs today): the unmodified
    modules in fp32 and after `.half()` (set_precision("half"), pipeline_wfd.py:207-212), same weights and latents as the
    CUDA arm. Reports ms per step, maps/s, and the deviations SURVEY 7.4 asks for: ours vs the reference's fp32, and the
    reference's own fp16 vs its fp32 (its half-precision backward underflows at random init)."""
    import torch

    from oracle.reference_loader import load_reference, reference_root

    if reference_root() is None:
        return None
    vae, losses = load_reference()
    torch.manual_seed(0)
    net = vae.AutoencoderTile(**cfg).eval().to(dev)
    lossm = losses.WFCLossEinsum(mx, my).to(dev)
    out = {"code": "unmodified reference modules (stub diffusers), lines pipeline_wfd.py:609-613, batch %d" % B}

    def step(x):
        with torch.enable_grad():
            _latents = x.detach().requires_grad_()
            wave = net.decode(1 / LATENT_SCALE * _latents).sample.softmax(axis=1)
            _loss = lossm(wave) * STRENGTH
            (grad,) = torch.autograd.grad(_loss.sum(), _latents)
        return _loss.detach().float(), grad.float()

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()

    ref32 = None
    for tag in ("fp32", "fp16"):
        if tag == "fp16":
            net.half()
            lossm.half()
        x = lat if tag == "fp32" else lat.half()
        try:
            step(x)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            n = 3
            for _ in range(n):
                l, g = step(x)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n
        except Exception as e:  # e.g. out of memory on a smaller part
            out[tag] = {"error": str(e)[:200]}
            continue
        rec = {"ms_per_step": ms, "maps_per_s": B / (ms / 1e3)}
        if tag == "fp32":
            ref32 = (l, g)
            rec["ours_vs_this"] = {"loss_rel_max": ((ours_loss - l).abs() / l.abs()).max().item(), "grad_rel_l2": rel(ours_grad, g)}
        elif ref32 is not None:
            rec["this_vs_reference_fp32"] = {"loss_rel_max": ((l - ref32[0]).abs() / ref32[0].abs()).max().item(),
                                             "grad_rel_l2": rel(g, ref32[1]),
                                             "grad_zero_fraction": (g == 0).float().mean().item()}
        out[tag] = rec
    del net, lossm
    torch.cuda.empty_cache()
    return out


def workload_config(args, T, t_pad, per_gpu_batch):
    return {"workload": "%s tileset %dx%d map, batch %d per GPU, tilenet_sc_space_%s random init (seed 0), one "
                        "WFC-guidance step = decode + softmax + WFC loss + grad to latents, strength %g"
                        % (args.tileset, args.latent, args.latent, per_gpu_batch, args.arch, STRENGTH),
            "tileset": args.tileset, "T": T, "T_pad": t_pad, "latent_hw": args.latent, "per_gpu_batch": per_gpu_batch,
            "l2": "working set per step (activations + operands, several GB) exceeds the 126 MB L2; no flush needed"}


def _timed_ms(fn, steps, warmup, dev, world):
    """Device time of `steps` calls after `warmup`, barrier + synchronize on both sides, max over ranks; ms per call."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item() / steps


def _guidance_runner(arch, tileset, B, latent, dev, seed_off=0):
    """(callable running one guidance step of B maps through the library, F_step per map, loss getter) for a config."""
    import torch

    from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config
    from wavefunctiondiffusion_b200.wf_diffusers.native_ops import GuidanceLossFunction
    from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile
    from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum

    mx, my = load_wfc_mats(tileset)
    T = mx.shape[0]
    t_pad = round_up_channels(T)
    cfg = tilenet_config(arch, t_pad)
    torch.manual_seed(0)
    net = AutoencoderTile(**cfg).eval().to(dev)
    handle = WFCLossEinsum(mx, my).to(dev).native_handle(t_pad, dev)
    lat = torch.randn(B, 4, latent, latent, generator=torch.Generator().manual_seed(1234 + seed_off)).to(dev)
    box = {}

    def step():
        box["loss"] = GuidanceLossFunction.apply(lat, net._native, handle, 1.0 / LATENT_SCALE, STRENGTH)

    return step, step_flops(cfg, latent, latent, T)["F_step"], (lambda: box["loss"]), net


def extra_configs(args, world, rank, dev, peak):
    """The other BASELINE configs under the driver's own command, as sub-objects of the one JSON line (the headline
    value stays config 2). One GPU: config 1 (batch 1), config 4's map on a single GPU, config 5 (32x32 architecture,
    batch 256 in micro-batches). N GPUs: config 3 (8 tilesets x 8 maps, balanced over the ranks by F_step) and config 4
    in row bands (one 256x256 map over all ranks, strong scaling). Every rank must call this."""
    import gc

    import torch
    import torch.distributed as dist

    from wavefunctiondiffusion_b200.utils.sharding import balance_by_cost

    out = {}

    def free():
        gc.collect()
        torch.cuda.empty_cache()

    def single(name, arch, tileset, B, latent, steps):
        step, f_map, get_loss, net = _guidance_runner(arch, tileset, B, latent, dev)
        ms = _timed_ms(step, steps, 2, dev, 1)
        out[name] = {"workload": "%s, %s architecture, %dx%d latents, batch %d" % (tileset, arch, latent, latent, B),
                     "ms_per_step": ms, "maps_per_s": B / (ms / 1e3), "loss_check": float(get_loss()[0]),
                     "frac_of_ideal": (f_map * B / (peak * 1e12)) / (ms / 1e3)}
        del step, get_loss, net
        free()

    def sparse_loss(B=8, latent=64):
        """SURVEY 8f-4, reported on its own: the loss contraction as a CUDA-core gather over the allowed pairs
        (csrc/wfc_sparse.cu) next to the dense tensor-core launch it replaces, on the bench workload's wave. Not part of
        `value` / `roofline`: the headline path keeps the tensor-core contraction north_star names."""
        import numpy as np

        from wavefunctiondiffusion_b200 import _native as nat
        from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels
        from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum

        mx, my = load_wfc_mats("ice_64x64")
        T = mx.shape[0]
        t_pad = round_up_channels(T)
        nnz4 = 2 * (int(np.count_nonzero(mx)) + int(np.count_nonzero(my)))  # every allowed pair is seen from both cells
        logits = 3.0 * torch.randn(B, t_pad, latent, latent, generator=torch.Generator(device=dev).manual_seed(7), device=dev)
        lib = nat.lib()
        res = {}
        for name in ("dense", "sparse"):
            mod = WFCLossEinsum(mx, my).to(dev).set_sparse(name == "sparse")
            x = logits.clone().requires_grad_()
            for _ in range(3):
                loss = mod(x, softmax=True)
            (g,) = torch.autograd.grad(loss.sum(), x)
            lib.wfd_profile_enable(1)
            reps = 10
            for _ in range(reps):
                mod(x, softmax=True)
            torch.cuda.synchronize()
            buf = C.create_string_buffer(1 << 16)
            lib.wfd_profile_report(buf, len(buf))
            lib.wfd_profile_enable(0)
            ms = 0.0
            for line in buf.value.decode().splitlines():
                n, cnt, tot = line.rsplit(" ", 2)
                if n.startswith("tapgemm") or n.startswith("wfc_sparse"):
                    ms += float(tot) / reps
            res[name] = (ms, loss.detach().double().cpu(), g.double().cpu())
            del mod, x
        cells = B * latent * latent
        d_ms, s_ms = res["dense"][0], res["sparse"][0]
        out["sparse_loss"] = {
            "workload": "ice 64x64 batch %d wave: a[n,i] of the WFC loss, CUDA-core gather over allowed pairs vs the dense tensor-core launch" % B,
            "sparse_ms_per_launch": s_ms, "dense_ms_per_launch": d_ms,
            "allowed_pairs_per_cell": nnz4, "gathered_pairs_per_s": cells * nnz4 / (s_ms / 1e3) if s_ms > 0 else None,
            "bound": "shared-memory gathers (two 32-bit gathers per pair and four cells)",
            "hbm_bytes_algorithmic": 2 * cells * t_pad * 2, "hbm_GBps": (2 * cells * t_pad * 2 / 1e9) / (s_ms / 1e3) if s_ms > 0 else None,
            "loss_rel_vs_dense": float(((res["sparse"][1] - res["dense"][1]).abs() / res["dense"][1].abs()).max()),
            "grad_rel_l2_vs_dense": float((res["sparse"][2] - res["dense"][2]).norm() / res["dense"][2].norm())}
        free()

    def wave_prior_propagation(B=8, latent=64):
        """the wave as the prior of the reference's WFC solver - allowed-tile bit sets seeded from a wave on
        the device and propagated to the fixed point (csrc/propagate.cu), at the bench shape. Integer work; reported per
        sweep with the state bytes a sweep touches at most (read own + four neighbours, write own)."""
        from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels
        from wavefunctiondiffusion_b200.wfc import WFCGeneratorPriorDist

        mx, my = load_wfc_mats("ice_64x64")
        T = mx.shape[0]
        t_pad = round_up_channels(T)
        g = torch.Generator(device=dev).manual_seed(5)
        cases = {}
        for name, sharp, thr in (("peaked_wave", 6.0, 2e-3), ("flat_wave", 1.0, 2e-4), ("all_tiles_allowed", 0.0, 1e-4)):
            wave = torch.softmax(sharp * torch.randn(B, latent, latent, t_pad, generator=g, device=dev), dim=-1).to(torch.bfloat16)
            gen = WFCGeneratorPriorDist.from_wave(wave, mx, my, threshold=thr)
            seed = gen.allowed_bits.clone()
            n0 = gen.count_allowed().float().mean().item()
            gen.propagate_all(2048)  # warm-up (builds the rule lists)
            gen.set_allowed_bits(seed.clone())
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ok = gen.propagate_all(2048)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            cases[name] = {"threshold": thr, "allowed_per_cell_seeded": n0,
                           "allowed_per_cell_fixed_point": gen.count_allowed().float().mean().item(),
                           "satisfiable": bool(ok), "sweeps": gen.last_sweeps, "ms_total": ms, "ms_per_sweep": ms / max(1, gen.last_sweeps)}
            del gen, wave, seed
        Wd = (T + 31) // 32
        out["wave_prior_propagation"] = {
            "workload": "ice 64x64 batch %d: allowed-tile bit sets (%d words per cell) seeded from a bf16 wave, swept to the fixed point" % (B, Wd),
            "state_bytes_per_full_sweep": B * latent * latent * Wd * 4 * 6, "cases": cases,
            "bound": "L1/L2-resident supporter lists + launch / barrier latency of one CTA per cell; HBM traffic is the state words"}
        free()

    if world == 1:
        for fn, key in ((sparse_loss, "sparse_loss"), (wave_prior_propagation, "wave_prior_propagation")):
            try:
                fn()
            except Exception as e:  # reported, never fatal for the headline line
                out[key] = {"error": repr(e)[:300]}
        single("config1_ice_64x64_batch1", "64x64", "ice_64x64", 1, 64, 10)
        single("config4_ice_256x256_one_gpu", "64x64", "ice_64x64", 1, 256, 3)
        single("config5_platform_32x32_batch256", "32x32", "platform_32x32", 256, 64, 2)
        return out
    # ---- config 3: 8 tilesets x 8 maps over the ranks, balanced by cost (maps are independent: no data-path collective)
    names = ["ashworld_64x64", "badlands_64x64", "desert_64x64", "ice_64x64", "install_64x64", "jungle_64x64",
             "platform_64x64", "twilight_64x64"]
    from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config

    costs = []
    for n in names:
        T = load_wfc_mats(n)[0].shape[0]
        f = step_flops(tilenet_config("64x64", round_up_channels(T)), 64, 64, T)["F_step"]
        costs += [f] * 8
    assign = balance_by_cost(costs, world)
    mine = {}
    for i in assign[rank]:
        mine[names[i // 8]] = mine.get(names[i // 8], 0) + 1
    runners = [_guidance_runner("64x64", n, b, 64, dev, seed_off=rank)[0] for n, b in sorted(mine.items())]

    def step3():
        for r in runners:
            r()

    ms3 = _timed_ms(step3, 5, 2, dev, world)
    loads = [sum(costs[i] for i in a) for a in assign]
    out["config3_all_tilesets_64_maps"] = {
        "workload": "8 tilesets x 8 maps (64x64), maps assigned to %d ranks by F_step (greedy longest-processing-time)" % world,
        "ms_per_step": ms3, "maps_per_s": 64 / (ms3 / 1e3), "maps_per_rank": [len(a) for a in assign],
        "f_step_balance": max(loads) / (sum(loads) / world),
        "frac_of_ideal": (sum(costs) / world / (peak * 1e12)) / (ms3 / 1e3)}
    del runners
    free()
    # ---- config 4: ONE 256x256 map in row bands over all ranks (strong scaling); rank 0 first times the single plan
    t1 = torch.zeros(1, device=dev)
    if rank == 0:
        step, _, _, net = _guidance_runner("64x64", "ice_64x64", 1, 256, dev)
        t1[0] = _timed_ms(step, 3, 2, dev, 1)
        del step, net
        free()
    dist.broadcast(t1, 0)
    rb = _rowband_core(args, 256, 5, 3, world, rank, dev)
    out["config4_rowband_256"] = dict(rb, one_gpu_ms_per_step=t1.item(), speedup=t1.item() / rb["ms_per_step"],
                                      efficiency=t1.item() / rb["ms_per_step"] / world)
    return out


def _rowband_core(args, L, steps, warmup, world, rank, dev):
    """One L x L ice map split into `world` row bands over NCCL; returns ms per step (graph replay) and the loss."""
    import torch

    from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config
    from wavefunctiondiffusion_b200.wf_diffusers.rowband import BandComm, BandGuidance
    from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile

    mx, my = load_wfc_mats("ice_64x64")
    t_pad = round_up_channels(mx.shape[0])
    torch.manual_seed(0)
    net = AutoencoderTile(**tilenet_config("64x64", t_pad)).eval().to(dev)
    comm = BandComm.from_torch_distributed(dev)
    lat = torch.randn(1, 4, L, L, generator=torch.Generator().manual_seed(1234)).to(dev)
    bg = BandGuidance(net, mx, my, L, L, comm, dev)
    box = {}

    def step():
        box["loss"], _ = bg.step(lat, 1.0 / LATENT_SCALE, STRENGTH, graph=True)

    ms = _timed_ms(step, steps, max(warmup, 3), dev, world)
    loss = float(box["loss"][0])
    bg.close()
    comm.close()
    return {"workload": "ONE ice 256x256 map in %d row bands of %d rows over NCCL (halo rows, GroupNorm sums, keys/values, "
                        "loss), CUDA-graph replay" % (world, L // world), "ms_per_step": ms, "maps_per_s": 1e3 / ms,
            "loss_check": loss}


def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    ge.build()
    from wavefunctiondiffusion_b200 import _native as nat
    from wavefunctiondiffusion_b200.utils.harness import AffineDDIMScheduler, ConstantTextEncoder, RandomNoiseUNet, TinyImageVAE, WhitespaceTokenizer
    from wavefunctiondiffusion_b200.wf_diffusers.pipeline_wfd import WaveFunctionDiffusionPipeline
    from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mx, my, T, t_pad, cfg = load_workload(args)
    B = args.batch
    torch.manual_seed(0)
    tile_vae = AutoencoderTile(**cfg).eval().to(dev)
    npz = "/tmp/wfd_bench_%s_%d.npz" % (args.tileset, rank)
    np.savez(npz, wfc_mats=np.stack([mx, my]))
    sched = AffineDDIMScheduler()
    sched.set_timesteps(50, device=dev)
    pipe = WaveFunctionDiffusionPipeline(TinyImageVAE(), ConstantTextEncoder(), WhitespaceTokenizer(), tile_vae,
                                         RandomNoiseUNet(4, args.latent), sched, None, None, wfc_data_path=npz)
    pipe.to(dev)
    g = torch.Generator().manual_seed(1234 + rank)
    lat_host = torch.randn(B, 4, args.latent, args.latent, generator=g).pin_memory()
    eps_host = torch.randn(B, 4, args.latent, args.latent, generator=g).pin_memory()
    out_host = torch.empty_like(lat_host).pin_memory()
    lat = lat_host.to(dev)
    lib = nat.lib()
    plan = tile_vae._native.plan(B, args.latent, args.latent, dev)
    handle = pipe.wfc_loss.native_handle(t_pad, dev)
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    dz = torch.empty(B, 4, args.latent, args.latent, dtype=torch.float32, device=dev)
    saved = handle.saved_ptr(B, plan.H, plan.W)

    def step_resident():
        nat.check(lib.wfd_guidance_step(plan.handle, handle.handle, C.c_void_p(lat.data_ptr()), 0, B, 1.0 / LATENT_SCALE,
                                        STRENGTH, saved, C.c_void_p(loss.data_ptr()), C.c_void_p(dz.data_ptr()),
                                        nat.stream_ptr()), "wfd_guidance_step")

    # the resident step is a fixed launch sequence: replay it from a CUDA graph (as the pipeline path does)
    launches_per_step = None
    if not args.no_graph:
        n0 = lib.wfd_launch_count()
        step_resident()
        launches_per_step = lib.wfd_launch_count() - n0
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_resident()
        step_plain = step_resident

        def step_resident():  # noqa: F811
            graph.replay()

    t_step = sched.timesteps[20]

    def step_e2e():
        l = lat_host.to(dev, non_blocking=True)
        e = eps_host.to(dev, non_blocking=True)
        new = pipe.wfc_guidance(l, e, t_step, STRENGTH)
        out_host.copy_(new, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.wfd_launch_count()
    ms = timed(step_resident, args.steps, max(args.warmup, 3))
    launches = (lib.wfd_launch_count() - n0) // (args.steps + max(args.warmup, 3)) * args.steps
    if launches_per_step is not None:
        launches = launches_per_step * args.steps  # kernels inside the replayed graph
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
    extras = None
    if not args.no_extra_configs:
        try:
            peak_tf = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0)
        except Exception:
            peak_tf = 1400.0
        extras = extra_configs(args, world, rank, dev, peak_tf)
    # per-kernel CUDA-event profile of a few extra steps (separate from the timed region) for the roofline numbers
    prof = {}
    if rank == 0:
        lib.wfd_profile_enable(1)
        psteps = 3
        for _ in range(psteps):
            (step_plain if not args.no_graph else step_resident)()
        torch.cuda.synchronize()
        buf = C.create_string_buffer(1 << 16)
        lib.wfd_profile_report(buf, len(buf))
        lib.wfd_profile_enable(0)
        for line in buf.value.decode().splitlines():
            name, cnt, tot = line.rsplit(" ", 2)
            prof[name] = (int(cnt) / psteps, float(tot) / psteps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_maps = world * B
    value = total_maps * args.steps / (ms / 1e3)
    e2e = total_maps * args.steps / (ms_e2e / 1e3)
    fl = step_flops(cfg, args.latent, args.latent, T)
    tg_ms = sum(t for n, (c, t) in prof.items() if n.startswith("tapgemm"))
    tg_launches = sum(c for n, (c, t) in prof.items() if n.startswith("tapgemm"))
    all_ms = sum(t for c, t in prof.values())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    # flops the tensor-core launches EXECUTE (the block-sparse WFC walk skips provably-zero K chunks), per launch class
    kc_kept = float(lib.wfd_wfc_kc_kept(handle.handle))
    alg, exe = class_flops(cfg, args.latent, args.latent, T, t_pad, kc_kept)
    cls_ms = {}
    for n, (c, t) in prof.items():
        if n.startswith("tapgemm["):
            k = n[len("tapgemm["):n.index("]")]
            cls_ms[k] = cls_ms.get(k, 0.0) + t
    classes = {}
    for k in sorted(set(cls_ms) | set(exe)):
        ms_k, f_k = cls_ms.get(k, 0.0), exe.get(k, 0.0) * B
        classes[k] = {"ms_per_step": ms_k, "executed_tflop_per_step": f_k / 1e12,
                      "tflops": f_k / (ms_k / 1e3) / 1e12 if ms_k > 0 else None,
                      "frac": f_k / (ms_k / 1e3) / 1e12 / peak if ms_k > 0 else None}
    f_exec = sum(exe.values()) * B
    achieved = f_exec / (tg_ms / 1e3) / 1e12 if tg_ms > 0 else None
    dense_eq = fl["F_step"] * B / (tg_ms / 1e3) / 1e12 if tg_ms > 0 else None
    weakest = min((k for k in classes if classes[k]["frac"] is not None and classes[k]["ms_per_step"] > 0.2),
                  key=lambda k: classes[k]["frac"], default=None)
    roofline = {"bound": "tensor", "kernel": "tapgemm_kernel (all tensor-core contractions of the step)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if achieved else None,
                "flops": "EXECUTED flops: algorithmic decoder flops (fwd + bwd-data, SURVEY 8d) + the WFC contraction over the "
                         "%.1f%% of its K chunks that hold an allowed pair (the rest are exact zeros and are skipped)" % (100 * kc_kept),
                "frac_dense_equivalent": dense_eq / peak if dense_eq else None,
                "wfc_chunks_kept": kc_kept,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)",
                "executed_flops_per_step": f_exec, "algorithmic_flops_per_step": fl["F_step"] * B,
                "launches_per_step": tg_launches,
                "avg_launch_ms": tg_ms / tg_launches if tg_launches else None, "kernel_ms_per_step": tg_ms,
                "share_of_step": tg_ms / all_ms if all_ms else None, "classes": classes, "weakest_class": weakest,
                "whole_step_frac_of_ideal": (fl["F_step"] * B / (peak * 1e12)) / (ms / args.steps / 1e3),
                "traffic": None}
    hbm_peak = peaks.get("hbm_gbs", 6551.4)
    # standalone GroupNorm passes: the launches that actually ran carry their shape in the profile name (layers whose
    # normalisation is applied inside the consuming conv have no gn_apply launch); 16-bit streams per launch: apply reads x
    # and writes y, the backward reduction reads dy and x, the backward apply reads dy, x (+ the residual gradient), writes dx
    gn = {}
    for n, (c, t) in prof.items():
        m = re.match(r"(gn_apply|gn_bwd_reduce|gn_bwd_apply):C=(\d+),HW=(\d+),B=(\d+)(?:,add=(\d))?", n)
        if not m:
            continue
        streams = {"gn_apply": 2, "gn_bwd_reduce": 2, "gn_bwd_apply": 3 + int(m.group(5) or 0)}[m.group(1)]
        e = gn.setdefault(m.group(1), {"ms": 0.0, "bytes": 0.0, "launches": 0.0})
        e["ms"] += t
        e["launches"] += c
        e["bytes"] += c * streams * 2.0 * int(m.group(2)) * int(m.group(3)) * int(m.group(4))
    roofline["hbm_kernels"] = {
        k: {"ms_per_step": e["ms"], "launches_per_step": e["launches"], "algorithmic_gb_per_step": e["bytes"] / 1e9,
            "gbps": e["bytes"] / (e["ms"] / 1e3) / 1e9 if e["ms"] > 0 else None,
            "frac": e["bytes"] / (e["ms"] / 1e3) / 1e9 / hbm_peak if e["ms"] > 0 else None}
        for k, e in gn.items()}
    # DRAM bytes per launch come from the committed ncu capture of the same workload (they cannot be measured live)
    if args.tileset == "ice_64x64" and args.latent == 64 and args.arch == "64x64" and B == 8:
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_tapgemm_dram.json")))
            roofline["traffic"] = tr["traffic_bytes_per_launch"]
            roofline["traffic_unit"] = "bytes per launch (dram read + write, mean over the %d launches of one step)" % tr["launches"]
            roofline["traffic_source"] = "profiles/r02_tapgemm_dram.json: " + tr["source"]
        except Exception:
            pass
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:args.top]
    if args.verbose:
        for n, (c, t) in top:
            print("  %-60s %6.1f launches %9.3f ms/step" % (n, c, t), file=sys.stderr)
        print("  profiled total %.3f ms/step, tapgemm %.3f ms" % (all_ms, tg_ms), file=sys.stderr)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sec, kind = cpu_oracle_maps_per_s(args, cfg, mx, my, repeats=3)
        cpu = {"value": v, "unit": "maps/s", "cores": cores, "kind": kind,
               "sample": "1 map of the batch (fp32, %s), 1 warm-up + best of 3, %.2f s per map-step"
                         % ("unmodified reference modules" if kind == "reference" else "oracle port", sec)}
    eager = None
    if world == 1 and not args.no_eager_baseline:
        eager = gpu_eager_baseline(args, cfg, mx, my, B, dev, loss.clone(), dz.clone(), lat)
    out = {
        "metric": "wfc_guidance_map_steps_per_s", "value": value, "unit": "maps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, T, t_pad, B),
        "steps_per_s": args.steps / (ms / 1e3), "cuda_graph": not args.no_graph,
        "e2e": {"value": e2e, "unit": "maps/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": 2 * lat_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                "api": "WaveFunctionDiffusionPipeline.wfc_guidance (pipeline_wfd.py:606-615 block) with pinned host tensors"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
        "other_configs": extras,
        "loss_check": [float(x) for x in loss.cpu()[:2]],
        "profile_top": [{"kernel": n, "launches_per_step": c, "ms_per_step": t} for n, (c, t) in top],
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_rowband(args):
    """ONE map split into row bands over the ranks (BASELINE config 4, strong scaling; wf_diffusers/rowband.py)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    ge.build()
    from wavefunctiondiffusion_b200 import _native as nat
    from wavefunctiondiffusion_b200.wf_diffusers.rowband import BandComm, BandGuidance, LoopbackGroup, gather_rows
    from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = BandComm.from_torch_distributed(dev)
    else:
        comm = LoopbackGroup(1).comm(0)
    mx, my, T, t_pad, cfg = load_workload(args)
    torch.manual_seed(0)
    tile_vae = AutoencoderTile(**cfg).eval().to(dev)
    L = args.latent
    parity = None
    gold = os.path.join(ROOT, "tests", "golden", "ice_c1_step.npz")
    if L == 64 and args.tileset == "ice_64x64" and args.arch == "64x64":
        z = np.load(gold)
        lat_host = torch.from_numpy(z["latents"]).pin_memory()
    else:
        z = None
        lat_host = torch.randn(1, 4, L, L, generator=torch.Generator().manual_seed(1234)).pin_memory()
    bg = BandGuidance(tile_vae, mx, my, L, L, comm, dev)
    lat = lat_host.to(dev)
    dz_host = torch.empty(1, 4, bg.h_band, L).pin_memory()
    lib = nat.lib()

    use_graph = not args.no_graph

    def step_resident():
        return bg.step(lat, 1.0 / LATENT_SCALE, STRENGTH, graph=use_graph)

    def step_e2e():
        l = lat_host.to(dev, non_blocking=True)
        loss, dz = bg.step(l, 1.0 / LATENT_SCALE, STRENGTH, graph=use_graph)
        dz_host.copy_(dz, non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    loss, dz = step_resident()
    full = gather_rows(dz, 2) if world > 1 else dz
    if z is not None and rank == 0:
        ref_l = float(np.asarray(z["loss_f64"]).reshape(-1)[0])
        g = torch.from_numpy(z["grad_f64"])
        parity = {"against": "tests/golden/ice_c1_step.npz (unmodified reference, fp64)",
                  "loss": loss.item(), "loss_ref": ref_l,
                  "grad_rel_l2": ((full.cpu().double() - g).norm() / g.norm()).item()}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    W = max(args.warmup, 3)
    n0 = lib.wfd_launch_count()
    bg.step(lat, 1.0 / LATENT_SCALE, STRENGTH)
    launches = (lib.wfd_launch_count() - n0) * args.steps  # the library's own kernels of one step (replayed from the graph)
    ms = timed(step_resident, args.steps, W)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps, W)
    prof = {}
    if rank == 0:
        lib.wfd_profile_enable(1)
    psteps = 2
    for _ in range(psteps):  # every rank must take part in the exchanges; plain launches (the per-kernel events are host side)
        bg.step(lat, 1.0 / LATENT_SCALE, STRENGTH)
    torch.cuda.synchronize()
    if rank == 0:
        buf = C.create_string_buffer(1 << 16)
        lib.wfd_profile_report(buf, len(buf))
        lib.wfd_profile_enable(0)
        for line in buf.value.decode().splitlines():
            name, cnt, tot = line.rsplit(" ", 2)
            prof[name] = (int(cnt) / psteps, float(tot) / psteps)
    # teardown in dependency order: graphs (they hold the NCCL exchanges), communicator, process group
    bg.close()
    comm.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    fl = step_flops(cfg, L, L, T)
    tg_ms = sum(t for n, (c, t) in prof.items() if n.startswith("tapgemm"))
    tg_launches = sum(c for n, (c, t) in prof.items() if n.startswith("tapgemm"))
    all_ms = sum(t for c, t in prof.values())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = fl["F_step"] / world / (tg_ms / 1e3) / 1e12 if tg_ms > 0 else None
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:args.top]
    if args.verbose:
        for n, (c, t) in top:
            print("  %-60s %6.1f launches %9.3f ms/step" % (n, c, t), file=sys.stderr)
        print("  profiled total %.3f ms/step, tapgemm %.3f ms" % (all_ms, tg_ms), file=sys.stderr)
    out = {
        "metric": "wfc_guidance_map_steps_per_s", "value": args.steps / (ms / 1e3), "unit": "maps/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "cuda_graph": use_graph,
        "config": {"workload": "ONE %s tileset %dx%d map (batch 1) split into %d row bands of %d rows, one per GPU; "
                               "tilenet_sc_space_%s random init (seed 0), one WFC-guidance step, strength %g"
                               % (args.tileset, L, L, world, bg.h_band, args.arch, STRENGTH),
                   "tileset": args.tileset, "T": T, "T_pad": t_pad, "latent_hw": L, "row_bands": world,
                   "exchange": "NCCL send/recv halos, all-reduce of GroupNorm sums and loss, all-gather of keys/values, "
                               "reduce-scatter of their gradients" if world > 1 else "none (one band)",
                   "l2": "working set per step exceeds the 126 MB L2; no flush needed"},
        "e2e": {"value": args.steps / (ms_e2e / 1e3), "unit": "maps/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": lat_host.numel() * 4, "d2h_bytes_per_step": dz_host.numel() * 4,
                "api": "rowband.BandGuidance.step with pinned host latents; this rank's dz rows read back"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "tapgemm_kernel (rank 0's share of the step's contractions)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if achieved else None,
                     "algorithmic_flops_per_step": fl["F_step"] / world, "launches_per_step": tg_launches,
                     "kernel_ms_per_step": tg_ms, "share_of_step": tg_ms / all_ms if all_ms else None, "traffic": None},
        "cpu_baseline": None, "parity": parity, "loss_check": [loss.item()],
        "profile_top": [{"kernel": n, "launches_per_step": c, "ms_per_step": t} for n, (c, t) in top],
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="maps per GPU")
    ap.add_argument("--tileset", default="ice_64x64")
    ap.add_argument("--arch", default="64x64", choices=["64x64", "32x32"])
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the other BASELINE configs (other_configs: batch 1, 256x256, 32x32 batch 256; N > 1: "
                         "all tilesets, row bands)")
    ap.add_argument("--no-eager-baseline", action="store_true",
                    help="skip timing the reference's eager PyTorch path on the GPU (gpu_eager_baseline)")
    ap.add_argument("--no-graph", action="store_true", help="plain launches instead of CUDA-graph replay")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--top", type=int, default=12, help="kernels listed in profile_top")
    ap.add_argument("--rowband", action="store_true",
                    help="strong scaling: ONE --latent x --latent map split into row bands over the GPUs (config 4)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.rowband:
        run_rowband(args)
        return
    run_native(args)


if __name__ == "__main__":
    main()
