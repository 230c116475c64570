#!/bin/bash
# launch list of one whole step (second step: launches 179..357) and a full capture of the top kernels
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 179 -c 179 --csv --log-file gpurun_out/launches_r1.csv python scripts/ncu_step.py 8 > gpurun_out/ncu_list.log 2>&1
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 28 -c 3 -o gpurun_out/prof_tapgemm_g64 python scripts/ncu_step.py 8 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | tail; tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
