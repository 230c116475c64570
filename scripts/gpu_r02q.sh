#!/bin/bash
# round 2, call Q: per-role counters of the launches with the operand transform
O=gpurun_out/r02q; mkdir -p $O
WFD_TG_DBG=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-extra-configs 2> $O/tgdbg.err >/dev/null
grep tgdbg $O/tgdbg.err | grep -v "XF 0 wait" | tail -24 > $O/roles_xf.txt
cat $O/roles_xf.txt | cut -c1-400
