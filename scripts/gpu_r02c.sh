#!/bin/bash
# round 2, call C: which engine bounds the row-reuse (grouped 3x3) contractions? WFD_TG_SKIP isolates TMA / MMA / epilogue
O=gpurun_out/r02c; mkdir -p $O
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --verbose --top 40 --no-cpu-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  echo "== $name"
  grep "reuse=1" $O/bench_$name.err | sort -k2 | head -12
}
run base A=1
run noTMA WFD_TG_SKIP=3
run noMMA WFD_TG_SKIP=4
run noEPI WFD_TG_SKIP=8
run onlyMMA WFD_TG_SKIP=11
run onlyTMA WFD_TG_SKIP=12
run onlyEPI WFD_TG_SKIP=7
run noA WFD_TG_SKIP=1
run noB WFD_TG_SKIP=2
