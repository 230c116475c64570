"""CPU: libwfd_b200.so loads and exports every symbol include/wfd_b200.h declares; host-side mirrors keep the
reference's interface (names, state_dict keys, error behaviour). No compute calls."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "wfd_b200.h")).read() + open(os.path.join(ROOT, "include", "wfd_b200_test.h")).read()
    return sorted(set(re.findall(r"WFD_API[^;(]*?\b(wfd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from wavefunctiondiffusion_b200 import _native as nat

    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(nat.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(nat.SIGNATURES) == names
    assert nat.lib().wfd_version() == 1
    # the product header carries no test hook
    product = open(os.path.join(ROOT, "include", "wfd_b200.h")).read()
    assert "wfd_test_tapgemm" not in product and "wfd_debug_set_ref_gemm" not in product


def test_hooks_need_the_environment_switch():
    """include/wfd_b200_test.h: a process started without WFD_ENABLE_TEST_HOOKS=1 cannot reroute its contractions."""
    import subprocess
    import sys

    code = ("import os; os.environ.pop('WFD_ENABLE_TEST_HOOKS', None); "
            "from wavefunctiondiffusion_b200 import _native as nat; l = nat.lib(); "
            "assert l.wfd_debug_set_ref_gemm(1) == -13 and b'WFD_ENABLE_TEST_HOOKS' in l.wfd_last_error(); print('ok')")
    env = {k: v for k, v in os.environ.items() if k != "WFD_ENABLE_TEST_HOOKS"}
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_state_dict_keys_and_config():
    from wavefunctiondiffusion_b200.utils.tilesets import tilenet_config
    from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile

    cfg = tilenet_config("32x32", 256)
    cfg.update(block_out_channels=(256, 128, 64, 64), conv_groups=(8, 4, 2, 1), conv_init_groups=16)
    net = AutoencoderTile(**cfg)
    keys = set(net.state_dict())
    for k in ["post_quant_conv.weight", "decoder.conv_in.bias", "decoder.mid_block.attentions.0.proj_attn.weight",
              "decoder.mid_block.resnets.1.conv2.weight", "decoder.up_blocks.2.resnets.0.conv_shortcut.weight",
              "decoder.up_blocks.2.downsamplers.0.conv.weight", "decoder.conv_norm_out.weight", "decoder.conv_out.weight",
              "encoder.down_blocks.1.upsamplers.0.conv.weight", "quant_conv.bias"]:
        assert k in keys, k
    assert net.config.conv_init_groups == 16 and net.in_channels == 256
    assert net.decoder.conv_out.weight.shape == (256, 256 // 16, 3, 3)
    net.enable_slicing()
    assert net.use_slicing
    net.disable_slicing()
    assert not net.use_slicing


def test_cpu_inputs_fail_loudly():
    from wavefunctiondiffusion_b200._native import WfdError
    from wavefunctiondiffusion_b200.utils.tilesets import tilenet_config
    from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile
    from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum

    cfg = tilenet_config("64x64", 128)
    cfg.update(block_out_channels=(128, 64, 64, 64), conv_groups=(2, 1, 1, 1), conv_init_groups=16)
    net = AutoencoderTile(**cfg)
    with pytest.raises(WfdError):
        net.decode(torch.zeros(1, 4, 16, 16))
    loss = WFCLossEinsum(torch.ones(5, 5), torch.ones(5, 5))
    assert loss.count == 5 and loss.cx.shape == (1, 5, 5, 1, 1)
    with pytest.raises(WfdError):
        loss(torch.zeros(1, 64, 4, 4))
    with pytest.raises(AssertionError):
        WFCLossEinsum(torch.ones(5, 5), torch.ones(4, 4))
