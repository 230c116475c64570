#!/bin/bash
# end-of-round evidence in one call: full GPU test suite, the bench line, the ncu launch list of one step, a full ncu
# capture of nine consecutive tensor-core launches (decoder block 3 forward ... WFC contraction), the per-role wait
# counters, and the single-GPU numbers of configs 1 and 4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log
python bench.py --steps 20 --warmup 3 --verbose --top 40 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print("bench: ms/step %.3f value %.1f e2e %.1f roofline %.3f launches %d cpu %s clocks %s"%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['gpu_launches'],d['cpu_baseline'],d['clocks']))
PY
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 161 -c 161 --csv --log-file gpurun_out/launches_r1c.csv python scripts/ncu_step.py 8 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
python scripts/ncu_step.py 8 > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 28 -c 9 -o gpurun_out/prof_tapgemm_r1c python scripts/ncu_step.py 8 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
bash scripts/gpu_tgdbg.sh
python bench.py --batch 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final_b1.json 2>/dev/null
python bench.py --batch 1 --latent 256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_256.json 2>/dev/null
python - <<'PY'
import json
for f in ("final_b1","final_256"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print("%s: ms/step %.3f value %.1f e2e %.1f frac %.3f"%(f,d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']))
    except Exception as e:
        print(f,"failed",e)
PY
ls -la gpurun_out | grep -E "ncu-rep|launches_r1c"
