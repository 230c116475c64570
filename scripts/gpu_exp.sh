#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 64 96 128 0; do
  WFD_EPI_SPLIT_MAX=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('epi_split_max $m: ms/step %.3f loss %s'%(d['ms_per_step'], d['loss_check']))"
done
