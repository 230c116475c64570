#!/bin/bash
# round 2, call A: suite on the uniform-warp default build, the suite with fp16 operands, bench profile, MMA micro-benchmark
O=gpurun_out/r02a; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > $O/pytest_bf16.log 2>&1; echo "pytest bf16 rc=$?"; tail -3 $O/pytest_bf16.log
WFD_DTYPE16=fp16 timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_fp16.log 2>&1; echo "pytest fp16 rc=$?"; tail -25 $O/pytest_fp16.log
python bench.py --steps 20 --warmup 3 --verbose --top 100 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -60 $O/bench.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I wavefunctiondiffusion_b200/csrc scripts/ubench/mma_rate.cu -o /tmp/mma_rate && timeout 120 /tmp/mma_rate > $O/mma_rate.txt 2>&1; echo "ubench rc=$?"
cat $O/mma_rate.txt
