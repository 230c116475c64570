#!/bin/bash
# round 2, call S: does the transform loop slow down because of the other roles? (role counters with the epilogue / MMA / TMA switched off)
O=gpurun_out/r02s; mkdir -p $O
for skip in 0 8 4 12 15; do
  WFD_TG_SKIP=$skip WFD_TG_DBG=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-extra-configs 2> $O/tgdbg_$skip.err >/dev/null
  echo "== skip=$skip"
  grep tgdbg $O/tgdbg_$skip.err | grep -v "XF 0 wait" | tail -24 | grep "N=3200\|N=3072 BN=48 taps=9 cpt=1 aC=3072 reuse=2 cg=2 side=0\|N=384 BN=192 taps=9 cpt=6 aC=384 reuse=2 cg=2 side=0" | tail -3 | sed -e 's/.*| MMA/MMA/' 
done
