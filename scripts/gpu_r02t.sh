#!/bin/bash
# round 2, call T: what slows the narrow-N MMAs inside the kernel? MMA-thread busy cycles per MMA with engines switched off
O=gpurun_out/r02t; mkdir -p $O
for skip in 0 8 1 2 3 11; do
  WFD_FUSE_APPLY=0 WFD_TG_SKIP=$skip WFD_TG_DBG=1 timeout 120 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-extra-configs 2> $O/tgdbg_$skip.err >/dev/null
  echo "== skip=$skip"
  grep tgdbg $O/tgdbg_$skip.err | tail -108 | grep "N=3072 BN=48 taps=9 cpt=1 aC=3072 reuse=2 cg=2 side=0\|N=1536 BN=96 taps=9 cpt=2 aC=1536 reuse=2 cg=2 side=0\|N=3200 BN=64" | tail -4 | python -c "
import sys,re
for l in sys.stdin:
    m=re.search(r'N=(\d+) BN=(\d+).*cpt=(\d+).*tiles=(\d+).*MMA (\d+) wte (\d+) wf (\d+) \| EPI (\d+) wtf (\d+) busy (\d+)', l)
    N,BN,cpt,tiles,mma,wte,wf,epi,wtf,eb=map(int,m.groups())
    per_tile={48:27,96:54,64:9*(4+2)}[BN]
    n=(tiles+73)//74*per_tile
    print('N=%d BN=%d: MMA total %d busy %d -> %.1f cyc/MMA (wte %d wf %d) | EPI busy %d'%(N,BN,mma,mma-wte-wf,(mma-wte-wf)/n,wte,wf,eb))
"
done
