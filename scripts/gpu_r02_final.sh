#!/bin/bash
# round 2 evidence in one call: the bench line (driver command), ncu launch list of the step, DRAM bytes of every
# tensor-core launch, full ncu sets of selected launches, per-role counters. Tests run separately (scripts/gpu_suite.sh).
cd "$(dirname "$0")/.."
O=gpurun_out/r02final2; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 3 --verbose --top 100 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02final2/bench.json').read().strip().splitlines()[-1])
print("bench: ms/step %.3f value %.1f e2e %.1f roofline %.3f launches %d clocks %s"%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['gpu_launches'],d['clocks']))
PY
timeout 300 python scripts/ncu_step.py 8 > $O/ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $O/launches_all.csv python scripts/ncu_step.py 8 > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log
if [ "$WFD_FINAL_FULL" = "1" ]; then
# full ncu sets: exported to CSV on the box (the .ncu-rep files together exceed what gpurun copies back)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 110 -c 24 -o $O/prof_tapgemm python scripts/ncu_step.py 8 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ncu -i $O/prof_tapgemm.ncu-rep --page raw --csv > $O/prof_tapgemm_raw.csv 2>/dev/null; rm -f $O/prof_tapgemm.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:gn_ -s 44 -c 6 -o $O/prof_gn python scripts/ncu_step.py 8 > $O/ncu_gn.log 2>&1
tail -2 $O/ncu_gn.log
ncu -i $O/prof_gn.ncu-rep --page raw --csv > $O/prof_gn_raw.csv 2>/dev/null; rm -f $O/prof_gn.ncu-rep
fi
# role counters need the diagnosis build (-DWFD_TG_DBG_BUILD); the product build is restored afterwards
WFD_NVCC_EXTRA=-DWFD_TG_DBG_BUILD=1 python -m wavefunctiondiffusion_b200.build --force > /dev/null 2>&1
WFD_TG_DBG=1 timeout 120 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-extra-configs 2> $O/tgdbg.err >/dev/null
grep tgdbg $O/tgdbg.err | tail -120 > $O/roles.txt; wc -l $O/roles.txt
python -m wavefunctiondiffusion_b200.build --force > /dev/null 2>&1
ls -la $O
