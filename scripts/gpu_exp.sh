#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 768 1536 3072; do
  WFD_FUSE_BWD_MAX_C=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse_bwd_max_c $c: ms/step %.3f'%d['ms_per_step'])"
done
# DRAM bytes of every tensor-core launch of one step (for roofline.traffic)
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tapgemm_kernel -s 75 -c 75 --csv --log-file gpurun_out/tapgemm_dram_r1c.csv python scripts/ncu_step.py 8 > gpurun_out/ncu_dram.log 2>&1
tail -1 gpurun_out/ncu_dram.log; wc -l gpurun_out/tapgemm_dram_r1c.csv
