#!/bin/bash
# first bench line + ncu launch list of the same command
python bench.py --steps 10 --warmup 3 --verbose > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -20 gpurun_out/bench.err
