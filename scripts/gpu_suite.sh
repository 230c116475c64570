#!/bin/bash
# the whole -m gpu suite (per-test timeout so that a hung kernel cannot eat the box's time limit)
O=gpurun_out/suite; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $O/pytest.log
