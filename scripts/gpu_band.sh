#!/bin/bash
# single-GPU row-band parity (loopback transport)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowband_gpu.py -x -q 2>&1 | tail -40 > gpurun_out/band.log
cat gpurun_out/band.log
