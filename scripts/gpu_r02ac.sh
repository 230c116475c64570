#!/bin/bash
# round 2, call AC: band width of the WFC contraction's tile order (N tiles swept per wave of CTAs): time and DRAM bytes
O=gpurun_out/r02ac; mkdir -p $O
for NG in 4 13 7 2; do
  WFD_WFC_NGROUP=$NG timeout 120 python bench.py --steps 10 --warmup 3 --verbose --top 80 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$NG.json 2> $O/bench_$NG.err
  WFD_WFC_NGROUP=$NG timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches_$NG.csv python scripts/ncu_step.py 8 > $O/ncu_$NG.log 2>&1
  python - $O $NG <<'PY'
import json, sys, csv
O, NG = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open('%s/bench_%s.json' % (O, NG)).read().strip().splitlines()[-1])
    c = d['roofline']['classes']['wfc']
    print("n_group %2s: step %.3f ms, wfc class %.3f ms" % (NG, d['ms_per_step'], c['ms_per_step']), end='')
except Exception as e:
    print("n_group", NG, "bench failed", e, end='')
rows = {}
try:
    lines = [l for l in open('%s/launches_%s.csv' % (O, NG)) if l.startswith('"')]
    for r in csv.DictReader(lines):
        k = r['ID']
        rows.setdefault(k, {'name': r['Kernel Name']})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    best = max((v for v in rows.values() if 'tapgemm' in v['name']), key=lambda v: v.get('dram__bytes_read.sum', 0))
    print(" | ncu: %.3f ms, dram read %.0f MB write %.0f MB" % (best['gpu__time_duration.sum'] / 1e6, best['dram__bytes_read.sum'] / 1e6, best['dram__bytes_write.sum'] / 1e6))
except Exception as e:
    print(" | ncu parse failed", e)
PY
done
