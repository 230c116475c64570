#!/bin/bash
# GPU parity run: tapgemm unit tests, small-config parity, then the full-size cases with per-launch fault localisation.
export WFD_DEBUG_SYNC=${WFD_DEBUG_SYNC:-0}
timeout 1500 python -m pytest tests/test_tapgemm_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 300 -k "not full_size and not ice_config1" > gpurun_out/parity_a.log 2>&1
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -s -k "full_size or ice_config1" > gpurun_out/parity_b.log 2>&1
grep -E "passed|failed|wfd debug|wfd:" gpurun_out/parity_a.log | tail -20; grep -E "^(FAILED|ERROR)" gpurun_out/parity_a.log | head -30
grep -E "passed|failed|wfd debug|wfd:|ice C1" gpurun_out/parity_b.log | tail -20
