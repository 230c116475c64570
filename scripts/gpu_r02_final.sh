#!/bin/bash
# round 2 evidence in one call: the bench line (driver command), ncu launch list of the step, DRAM bytes of every
# tensor-core launch, full ncu sets of selected launches, per-role counters. Tests run separately (scripts/gpu_suite.sh).
cd "$(dirname "$0")/.."
O=gpurun_out/r02final; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 3 --verbose --top 100 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02final/bench.json').read().strip().splitlines()[-1])
print("bench: ms/step %.3f value %.1f e2e %.1f roofline %.3f launches %d clocks %s"%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['gpu_launches'],d['clocks']))
PY
timeout 300 python scripts/ncu_step.py 8 > $O/ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $O/launches_all.csv python scripts/ncu_step.py 8 > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 110 -c 24 -o $O/prof_tapgemm python scripts/ncu_step.py 8 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
timeout 600 ncu --set full --clock-control none -k regex:gn_ -s 44 -c 6 -o $O/prof_gn python scripts/ncu_step.py 8 > $O/ncu_gn.log 2>&1
tail -2 $O/ncu_gn.log
WFD_TG_DBG=1 timeout 120 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-extra-configs 2> $O/tgdbg.err >/dev/null
grep tgdbg $O/tgdbg.err | tail -120 > $O/roles.txt; wc -l $O/roles.txt
ls -la $O
