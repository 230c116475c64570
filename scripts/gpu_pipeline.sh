#!/bin/bash
timeout 1700 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --timeout 900 -s 2>&1 | tail -30
python __graft_entry__.py --smoke 2>&1 | tail -3
