#!/bin/bash
# round-1 final profile: launch list of one whole step + full captures of the WFC contraction (CTA pair) and a groups=64 conv
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 161 -c 161 --csv --log-file gpurun_out/launches_r1b.csv python scripts/ncu_step.py 8 > gpurun_out/ncu_list.log 2>&1
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 28 -c 9 -o gpurun_out/prof_tapgemm_r1b python scripts/ncu_step.py 8 > gpurun_out/ncu_full.log 2>&1
python bench.py --batch 1 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=1 ms/step %.3f value %.1f e2e %.1f launches %d'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches']))"
ls -la gpurun_out | tail -5
