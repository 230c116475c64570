#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for u in 2 4; do
  WFD_GN_REDUCE_U=$u python bench.py --steps 10 --warmup 3 --no-cpu-baseline --verbose --top 8 2> gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gn_reduce U=$u: ms/step %.3f'%d['ms_per_step'])"
  grep "gn_bwd_reduce" gpurun_out/exp.err
done
