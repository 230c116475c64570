"""Summarise gpurun_out/tgdbg.txt (WFD_TG_DBG=1 per-role wait counters of CTA 0, see scripts/gpu_tgdbg.sh)."""
import collections
import re
import sys

rows = collections.OrderedDict()
for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tgdbg.txt"):
    m = re.match(r"tgdbg (.*?) \| A (\d+) we (\d+) is (\d+) \| B (\d+) we (\d+) is (\d+) \| MMA (\d+) wte (\d+) wf (\d+) \| EPI (\d+) wtf (\d+) busy (\d+)", l)
    if not m:
        continue
    sig = m.group(1)
    v = list(map(int, m.groups()[1:]))
    r = rows.setdefault(sig, [0] * 13)
    for i, x in enumerate(v):
        r[i] += x
    r[12] += 1
print("%-98s %3s %7s | %5s %5s | %5s %6s | %6s %7s" % ("launch", "n", "Mcyc", "A_we", "A_is", "MMAwf", "MMAwte", "EPIwtf", "EPIbusy"))
for sig, r in sorted(rows.items(), key=lambda kv: -kv[1][6]):
    A, Awe, Ais, B, Bwe, Bis, M, Mwte, Mwf, E, Ewtf, Eb, n = r
    print("%-98s %3d %7.2f | %5.2f %5.2f | %5.2f %6.2f | %6.2f %7.2f" % (sig, n, max(A, M, E, 1) / 1e6, Awe / max(A, 1), Ais / max(A, 1),
                                                                     Mwf / max(M, 1), Mwte / max(M, 1), Ewtf / max(E, 1), Eb / max(E, 1)))
