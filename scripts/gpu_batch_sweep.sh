#!/bin/bash
for b in 2 4 8 16; do
python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$b ms/step %.3f  ms/map %.3f value %.1f'%(d['ms_per_step'],d['ms_per_step']/$b,d['value']))" || tail -3 gpurun_out/bench.err
done
