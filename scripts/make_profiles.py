"""Regenerates the tracked evidence under profiles/ from one run directory of scripts/gpu_r02_final.sh
(gpurun_out/<dir>: bench.json, bench.err, launches_all.csv, prof_tapgemm_raw.csv, prof_gn_raw.csv, roles.txt) and from
the built library (SASS mnemonic counts, per-kernel registers / stack).

    python scripts/make_profiles.py gpurun_out/r02final3 [prefix=r02]
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
pre = sys.argv[2] if len(sys.argv) > 2 else "r02"
P = os.path.join(ROOT, "profiles")


def out(name):
    return os.path.join(P, "%s_%s" % (pre, name))


def head_commit():
    try:
        return subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except Exception:
        return "?"


# ---- the bench line and its per-launch CUDA-event profile
line = open(os.path.join(src, "bench.json")).read().strip().splitlines()[-1]
bench = json.loads(line)
open(out("bench.json"), "w").write(json.dumps(bench, indent=1) + "\n")
open(out("bench_profile.txt"), "w").write(open(os.path.join(src, "bench.err")).read())

# ---- ncu launch list (long format: one row per launch and metric) -> one row per launch of the LAST step
rows = {}
order = []
for r in csv.DictReader(l for l in open(os.path.join(src, "launches_all.csv")) if l.startswith('"')):
    k = int(r["ID"])
    if k not in rows:
        rows[k] = {"kernel": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]}
        order.append(k)
    rows[k][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
# scripts/ncu_step.py executes the step several times (warm-up, graph replays); every execution starts with latent_in_k
starts = [i for i, k in enumerate(order) if "latent_in_k" in rows[k]["kernel"]]
last = order[starts[-1]:] if starts else order


def short(name):
    m = re.match(r"(?:void )?(?:wfd::)?([A-Za-z0-9_]+(?:<[^>]*>)?)", name)
    return m.group(1) if m else name


with open(out("launches_step_b8.csv"), "w") as f:
    f.write("id,kernel,grid,block,gpu_time_ns,dram_read_bytes,dram_write_bytes\n")
    for i, k in enumerate(last):
        r = rows[k]
        f.write('%d,"%s","%s","%s",%d,%d,%d\n' % (i, short(r["kernel"]), r["grid"], r["block"], r["gpu__time_duration.sum"],
                                                r["dram__bytes_read.sum"], r["dram__bytes_write.sum"]))
tg = [rows[k] for k in last if "tapgemm" in rows[k]["kernel"]]
dram = {
    "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 python "
              "scripts/ncu_step.py 8 (last execution of the step; ice 64x64 batch 8); profiles/%s_launches_step_b8.csv; build %s" % (pre, head_commit()),
    "launches": len(tg),
    "dram_read_bytes": sum(r["dram__bytes_read.sum"] for r in tg),
    "dram_write_bytes": sum(r["dram__bytes_write.sum"] for r in tg),
    "ncu_ms_total": sum(r["gpu__time_duration.sum"] for r in tg) / 1e6,
}
dram["traffic_bytes_per_launch"] = (dram["dram_read_bytes"] + dram["dram_write_bytes"]) / max(1, len(tg))
open(out("tapgemm_dram.json"), "w").write(json.dumps(dram, indent=1) + "\n")
by = {}
for k in last:
    r = rows[k]
    b = by.setdefault(short(r["kernel"]), [0, 0.0, 0.0, 0.0])
    b[0] += 1
    b[1] += r["gpu__time_duration.sum"] / 1e6
    b[2] += r["dram__bytes_read.sum"] / 1e6
    b[3] += r["dram__bytes_write.sum"] / 1e6
tot = sum(b[1] for b in by.values())
with open(out("launches_by_kernel.txt"), "w") as f:
    f.write("launches %d, total %.3f ms under ncu (cold caches, serialised; build %s)\n" % (len(last), tot, head_commit()))
    for name, b in sorted(by.items(), key=lambda kv: -kv[1][1]):
        f.write("%-46s %4d  %8.3f ms  %5.1f%%  dram rd %9.1f MB  wr %9.1f MB\n" % (name, b[0], b[1], 100 * b[1] / tot, b[2], b[3]))


# ---- ncu --set full exports (wide format, second row = units): a handful of columns per launch
def full(raw, dst, header, cols):
    path = os.path.join(src, raw)
    if not os.path.exists(path):
        return
    rd = list(csv.reader(l for l in open(path) if l.startswith('"')))
    names, units, data = rd[0], rd[1], rd[2:]

    def find(key):
        return names.index(key) if key in names else None

    idx = [(c, find(c)) for c in cols]
    with open(dst, "w") as f:
        f.write("# %s (build %s)\n" % (header, head_commit()))
        f.write("id,kernel,grid,block," + ",".join("%s [%s]" % (c, units[i]) if i is not None else c for c, i in idx) + "\n")
        for d in data:
            f.write('%s,"%s",%s,%s,' % (d[0], short(d[names.index("Kernel Name")])[:40], d[names.index("Grid Size")].strip("()").split(",")[0],
                                        d[names.index("Block Size")].strip("()").split(",")[0]))
            f.write(",".join(d[i].replace(",", "") if i is not None else "" for c, i in idx) + "\n")


full("prof_tapgemm_raw.csv", out("ncu_full_tapgemm.csv"),
     "ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 110 -c 24 python scripts/ncu_step.py 8: "
     "tensor-core launches 33..56 of the second step (ice 64x64 batch 8)",
     ["launch__registers_per_thread", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
      "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
      "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
      "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
      "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
      "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
      "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"])
full("prof_gn_raw.csv", out("ncu_full_gn.csv"),
     "ncu --set full --clock-control none -k regex:gn_ -s 44 -c 6 python scripts/ncu_step.py 8 (GroupNorm launches of the second step)",
     ["launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
      "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
      "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
      "sm__warps_active.avg.pct_of_peak_sustained_active"])

# ---- per-role counters of the diagnosis build
if os.path.exists(os.path.join(src, "roles.txt")):
    open(out("roles_step_b8.txt"), "w").write(open(os.path.join(src, "roles.txt")).read())

# ---- SASS of the built library
lib = os.path.join(ROOT, "wavefunctiondiffusion_b200", "libwfd_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
with open(out("sass.txt"), "w") as f:
    f.write("# cuobjdump -sass wavefunctiondiffusion_b200/libwfd_b200.so | grep -c <mnemonic>  (build of commit %s, nvcc 12.9, "
            "-gencode arch=compute_100a,code=sm_100a)\n" % head_commit())
    lines = sass.splitlines()
    for mn in ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.4D", "UTMALDG.3D", "UTMALDG.4D.2CTA", "UTMAPF", "LDTM", "LDTM.x16", "UTCBAR",
               "UTCBAR.2CTA.MULTICAST", "SYNCS", "R2UR", "MUFU.TANH", "ATOMS", "STL", "LDL"]:
        f.write("%-28s %d\n" % (mn, sum(1 for l in lines if mn in l)))
    f.write("# cuobjdump -res-usage: registers / stack of every kernel (STACK 8 = no spill)\n")
    cur = None
    for l in res.splitlines():
        l = l.strip()
        if l.startswith("Function "):
            cur = l.split()[1].rstrip(":")
        elif cur and l.startswith("REG:"):
            f.write("Function %s\t%s\n" % (cur, "\t".join(l.split()[:3])))
            cur = None
print("profiles written with prefix", pre, "from", src)
