#!/bin/bash
# round 2, call AF: the fused GroupNorm-backward drain and the plain conv drains as their own kernel instantiations
# (WFD_NO_EPI_SPLIT=1: plain conv launches through the common kernel): class times, quick parity
O=gpurun_out/r02af; mkdir -p $O
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --verbose --top 80 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print("%-10s ms/step %.3f frac %.3f mhz %s loss %s"%(sys.argv[2], d['ms_per_step'], r['frac'], d['clocks']['sm_mhz'], d['loss_check'][:1]), {k:round(v['ms_per_step'],3) for k,v in r['classes'].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run split1 A=1
run common WFD_NO_EPI_SPLIT=1
run split1b A=1
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 100 -x -k "small or ice_config" 2>&1 | tail -2
