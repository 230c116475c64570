#!/bin/bash
# patch height of the row-reuse contractions: bench at 4 and 8, then the whole GPU suite at 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for hb in 4 8; do
  WFD_REUSE_HBOX=$hb python bench.py --steps 10 --warmup 3 --no-cpu-baseline --verbose --top 40 2> gpurun_out/hbox_$hb.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reuse_hbox $hb: ms/step %.3f loss %s'%(d['ms_per_step'], d['loss_check']))"
  grep -E "reuse=1" gpurun_out/hbox_$hb.err | head -6
done
WFD_REUSE_HBOX=8 timeout 400 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3
