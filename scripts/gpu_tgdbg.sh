#!/bin/bash
# per-role wait counters (CTA 0) of every tensor-core launch of one guidance step
cd "$(dirname "$0")/.."
WFD_TG_DBG=1 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline 2> gpurun_out/tgdbg.err >/dev/null
grep tgdbg gpurun_out/tgdbg.err | tail -108 > gpurun_out/tgdbg.txt
wc -l gpurun_out/tgdbg.txt
