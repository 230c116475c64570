"""Reads bench.py's JSON line from stdin and prints the step time and the per-class kernel times (A/B scripts)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d["roofline"]
print("ms/step %.3f frac %.3f loss %s" % (d["ms_per_step"], r["frac"], d["loss_check"][:1]),
      {k: round(v["ms_per_step"], 3) for k, v in r["classes"].items()},
      {k: round(v["ms_per_step"], 3) for k, v in r["hbm_kernels"].items()})
