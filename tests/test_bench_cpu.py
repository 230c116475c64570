"""The CPU legs of bench.py: the --impl reference arm prints one JSON line with the contract's keys (rank 0 only under
torchrun), and the CUDA arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=env, timeout=900, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "wfc_guidance_map_steps_per_s" and d["unit"] == "maps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f32"
    from oracle.reference_loader import reference_root

    # the unmodified reference modules when they are present (/root/reference or the baseline/_ref install), else the port
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_root() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] == 1 and "1 map of the batch" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same `config` object as the CUDA arm prints for the same arguments (the driver compares them)
    assert d["config"]["per_gpu_batch"] == 8 and set(d["config"]) == {"workload", "tileset", "T", "T_pad", "latent_hw",
                                                                      "per_gpu_batch", "l2"}
    assert abs(d["loss_check"][0] - 4.6) < 0.01


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return  # on a GPU box the arm runs; the GPU tier covers it
    r = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
