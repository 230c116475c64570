#!/bin/bash
# round 2, call G: ncu --set full with source on the groups=64 launches of the second (warm) step
O=gpurun_out/r02g; mkdir -p $O
python scripts/ncu_step.py 8 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 103 -c 5 -o $O/prof_g64 python scripts/ncu_step.py 8 > $O/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/ncu.log; ls -la $O
