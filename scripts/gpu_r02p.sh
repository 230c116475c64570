#!/bin/bash
# round 2, call P: special-function unit rates (which sigmoid formulation the fused GroupNorm+SiLU transform should use)
O=gpurun_out/r02p; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/ubench/mufu_rate.cu -o /tmp/mufu_rate && timeout 60 /tmp/mufu_rate > $O/mufu_rate.txt 2>&1; echo "rc=$?"
cat $O/mufu_rate.txt
