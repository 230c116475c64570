#!/bin/bash
# round 2, call AD: L2 evict_last for the adjacency operand of the WFC contraction x band width of its tile order:
# DRAM bytes and time of that launch (ncu launch list of one step), then the bench line of the chosen default
O=gpurun_out/r02ad; mkdir -p $O
for CASE in "4 1" "4 0" "13 0" "13 1"; do
  set -- $CASE; NG=$1; NOEL=$2
  WFD_WFC_NGROUP=$NG WFD_WFC_NO_EVICT_LAST=$NOEL timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches_${NG}_${NOEL}.csv python scripts/ncu_step.py 8 > $O/ncu_${NG}_${NOEL}.log 2>&1
  python - $O/launches_${NG}_${NOEL}.csv $NG $NOEL <<'PY'
import sys, csv
rows = {}
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.setdefault(r['ID'], {'name': r['Kernel Name']})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
tg = [v for v in rows.values() if 'tapgemm' in v['name']]
best = max(tg, key=lambda v: v.get('dram__bytes_read.sum', 0))
tot = sum(v['gpu__time_duration.sum'] for v in rows.values()) / 2e6
print("n_group %2s evict_last %s: wfc launch %.3f ms, dram read %.0f MB write %.0f MB | all launches of a step %.3f ms" % (
    sys.argv[2], 'off' if sys.argv[3] == '1' else 'on', best['gpu__time_duration.sum'] / 1e6, best['dram__bytes_read.sum'] / 1e6, best['dram__bytes_write.sum'] / 1e6, tot))
PY
done
tail -2 $O/ncu_13_0.log
