#!/bin/bash
# round 2, call B: MMA micro-benchmark (fixed), A/B of the grouped-conv knobs on the uniform-warp build, role counters
O=gpurun_out/r02b; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I wavefunctiondiffusion_b200/csrc scripts/ubench/mma_rate.cu -o /tmp/mma_rate && timeout 120 /tmp/mma_rate > $O/mma_rate.txt 2>&1; echo "ubench rc=$?"
cat $O/mma_rate.txt
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --verbose --top 14 --no-cpu-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("%-22s ms/step %.3f  loss %s"%(sys.argv[2], d['ms_per_step'], d['loss_check']))
PY
  grep "tapgemm:N=3072,BN=48,taps=9\|tapgemm:N=1536,BN=96,taps=9\|N=3200,BN=64" $O/bench_$name.err | head -8
}
run base A=1
run no2cta WFD_REUSE_2CTA=0
run hbox4 WFD_REUSE_HBOX=4
run hbox8 WFD_REUSE_HBOX=8
run no2cta_hbox8 WFD_REUSE_2CTA=0 WFD_REUSE_HBOX=8
run stages3 WFD_MAX_STAGES=3
WFD_TG_DBG=1 python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/roles.txt
python scripts/tgdbg_report.py $O/roles.txt | tail -40
