#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0 1 2; do
  WFD_WFC_DBG=1 WFD_WFC_DENSE=$d python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline 2>&1 >/dev/null | grep "wfc dbg" | head -1 | sed "s/^/dense $d: /"
done
