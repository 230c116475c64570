#!/bin/bash
O=gpurun_out/r02y; mkdir -p $O
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 100 -s -k "two_kernel" > $O/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -i "rel-L2\|passed\|failed" $O/pytest_quick.log | tail -12
