#!/bin/bash
# memcheck of the small-config paths (single plan, row bands on the loopback transport)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_rowband_gpu.py tests/test_parity_gpu.py -x -q -k "tiny_arch or fused_guidance_step_small or argmax" --timeout 450 > gpurun_out/sanitize.log 2>&1
echo "rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize.log | tail -12
