#!/bin/bash
# round 2, call Z: role counters of the groups=64 / groups=16 launches with the specialised epilogue (diagnosis build)
O=gpurun_out/r02z; mkdir -p $O
timeout 60 python scripts/diag_g64_layout.py 2>&1 | tail -2
WFD_NVCC_EXTRA=-DWFD_TG_DBG_BUILD=1 python -m wavefunctiondiffusion_b200.build --force > /dev/null 2>&1
WFD_TG_DBG=1 timeout 120 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-extra-configs 2> $O/tgdbg.err >/dev/null
grep tgdbg $O/tgdbg.err | tail -120 > $O/roles.txt
grep "N=3072 BN=48 taps=9\|N=3200 BN=64\|N=1536 BN=96 taps=9 cpt=2" $O/roles.txt | sort | uniq -c | sort -rn | head -3 > /dev/null
python - <<'PY'
import re
seen={}
for l in open('gpurun_out/r02z/roles.txt'):
    m=re.search(r'N=(\d+) BN=(\d+) taps=(\d+) cpt=(\d+) aC=(\d+) reuse=(\d) cg=(\d) side=(\d) tiles=(\d+).*\| MMA (\d+) wte (\d+) wf (\d+) \| EPI (\d+) wtf (\d+) busy (\d+)', l)
    if not m: continue
    k=m.group(0).split(' tiles')[0]
    if k in seen: continue
    seen[k]=1
    N,BN,taps,cpt,aC,reuse,cg,side,tiles,mma,wte,wf,epi,wtf,eb=map(int,m.groups())
    if taps!=9 or BN>96: continue
    per=(tiles+73)//74
    print("%-60s tiles/pair %3d | MMA %7d (wte %5.1f%% wf %5.1f%%) %5d cyc/tile | EPI busy %5d cyc/tile(half-set)"%(k,per,mma,100*wte/mma,100*wf/mma,mma//per,eb*2//per))
PY
python -m wavefunctiondiffusion_b200.build --force > /dev/null 2>&1
