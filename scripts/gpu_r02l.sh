#!/bin/bash
# round 2, call L: state check after the encoder commit (full -m gpu suite) + the issue-pattern rows of the MMA micro-benchmark
O=gpurun_out/r02l; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I wavefunctiondiffusion_b200/csrc scripts/ubench/mma_rate.cu -o /tmp/mma_rate && timeout 120 /tmp/mma_rate > $O/mma_rate.txt 2>&1; echo "ubench rc=$?"
tail -12 $O/mma_rate.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest.log
