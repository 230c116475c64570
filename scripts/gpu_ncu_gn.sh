#!/bin/bash
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gn_apply_k -s 24 -c 6 -o gpurun_out/prof_gn_apply python scripts/ncu_step.py 8 > gpurun_out/ncu_gn1.log 2>&1
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gn_bwd -s 0 -c 4 -o gpurun_out/prof_gn_bwd python scripts/ncu_step.py 8 > gpurun_out/ncu_gn2.log 2>&1
ls -la gpurun_out/*.ncu-rep
