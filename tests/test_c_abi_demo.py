"""The C ABI is usable from plain C: examples/c_abi_demo.c compiles with gcc -std=c99 against include/wfd_b200.h and
libwfd_b200.so (no Python, no torch in the process), and on a GPU runs two whole guidance steps."""
import os
import shutil
import subprocess

import pytest

from wavefunctiondiffusion_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _compile(out):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    nat.lib()  # the library must be built
    cmd = ["gcc", "-std=c99", "-O2", os.path.join(ROOT, "examples", "c_abi_demo.c"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(CUDA, "include"), "-L" + os.path.dirname(nat.LIB_PATH), "-lwfd_b200",
           "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-lm", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_c_demo_compiles_as_c99(tmp_path):
    _compile(str(tmp_path / "c_abi_demo"))


@pytest.mark.gpu
def test_c_demo_runs_two_guidance_steps(tmp_path):
    exe = _compile(str(tmp_path / "c_abi_demo"))
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.dirname(nat.LIB_PATH), os.path.join(CUDA, "lib64"),
                                              env.get("LD_LIBRARY_PATH", "")])
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
