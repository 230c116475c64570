#!/bin/bash
# other BASELINE configs, for the record (config 2 is the bench default)
run() { python bench.py "$@" --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('%-18s latent %3d batch %3d: ms/step %9.3f  maps/s %8.1f  e2e %8.1f  tensor frac %.3f'%(c['tileset'],c['latent_hw'],c['per_gpu_batch'],d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'] or 0))" || tail -3 gpurun_out/bench.err; }
run --tileset ice_64x64 --batch 1
for t in ashworld badlands desert install jungle platform twilight; do run --tileset ${t}_64x64 --batch 8; done
run --tileset platform_32x32 --arch 32x32 --batch 64
run --tileset platform_32x32 --arch 32x32 --batch 256
run --tileset ice_64x64 --latent 128 --batch 2
run --tileset ice_64x64 --latent 256 --batch 1
