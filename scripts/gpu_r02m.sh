#!/bin/bash
# round 2, call M: which part of the patch schedule's issue pattern costs 85 cycles per narrow MMA (alignment of the A start / SBO)
O=gpurun_out/r02m; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I wavefunctiondiffusion_b200/csrc scripts/ubench/mma_rate.cu -o /tmp/mma_rate && timeout 60 /tmp/mma_rate > $O/mma_rate.txt 2>&1; echo "ubench rc=$?"
tail -16 $O/mma_rate.txt
