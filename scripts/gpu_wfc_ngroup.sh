#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0 1 2; do
  WFD_WFC_DENSE=$d python bench.py --steps 5 --warmup 3 --verbose --no-cpu-baseline > gpurun_out/ng.json 2> gpurun_out/ng.err
  echo "dense $d: $(grep 'taps=4' gpurun_out/ng.err) | $(python -c "import json;d=json.loads(open('gpurun_out/ng.json').read().strip().splitlines()[-1]);print('ms/step %.3f'%d['ms_per_step'])")"
done
