#!/usr/bin/env python
"""Turn the raw gpurun_out/ artefacts of tools/gpu_round.sh into the small tracked summaries under profiles/.
usage: python tools/summarise_profiles.py r01_b      (tag = round + visit)"""
import csv, json, os, re, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

# 1. launch list (ncu --metrics gpu__time_duration.sum): per-kernel totals and shares over the captured launches
rows = [r for r in csv.reader(l for l in open(os.path.join(G, "launches.csv")) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[ix["Metric Value"]]) / 1e3
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"launches_{tag}.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 500 python bench.py --steps 8 --warmup 3` (YDST_GRAPH=0, default micro-batch)\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / tot:.3f} |\n")
    f.write(f"\n{len(data)} launches, {tot:.1f} us total\n")
with open(os.path.join(P, f"launches_{tag}.csv"), "w") as f:
    f.write("id,kernel,grid,block,ns\n")
    for r in data:
        f.write(f'{r[ix["ID"]]},{re.sub(r"\(.*", "", r[ix["Kernel Name"]])},"{r[ix["Grid Size"]]}","{r[ix["Block Size"]]}",{r[ix["Metric Value"]]}\n')

# 2. ncu --set full of the representative convolutions
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
full = []
for name, what in (("76", ["1x1 256->128 @76x76", "3x3 128->256 @76x76"]), ("19", ["1x1 1024->512 @19x19", "3x3 512->1024 @19x19"]),
                   ("mb", ["3x3 128->256 @76x76, default micro-batch"])):
    path = os.path.join(G, f"prof_full_{name}.csv")
    if not os.path.exists(path):
        continue
    rr = list(csv.reader(open(path)))
    h, units = rr[0], rr[1]
    ii = {k: i for i, k in enumerate(h)}
    for j, r in enumerate(rr[2:]):
        d = OrderedDict(layer=what[j] if j < len(what) else "?", kernel=re.sub(r"\(.*", "", r[ii["Kernel Name"]]), grid=r[ii["Grid Size"]])
        for k in KEYS:
            if k in ii:
                d[k] = f"{r[ii[k]]} {units[ii[k]]}".strip()
        stalls = sorted(((float(r[i].replace(",", "")), k) for k, i in ii.items() if "issue_stalled" in k and "per_issue_active" in k and r[i] not in ("", "n/a")), reverse=True)[:5]
        d["top_stalls_per_issue"] = [(round(v, 2), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for v, k in stalls]
        full.append(d)
if full:
    with open(os.path.join(P, f"conv_full_{tag}.json"), "w") as f:
        json.dump(full, f, indent=1)
    # dram traffic per launch of the dominant kernel class (3x3 128->256 @76, the largest FLOP share of yolov3-608)
    def num(s):
        v, u = s.split()[0], (s.split() + [""])[1]
        return float(v.replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
    for d in sorted(full, key=lambda d: "micro-batch" in d["layer"]):        # the micro-batch capture wins when present
        if d["layer"].startswith("3x3 128->256"):
            t = num(d["dram__bytes_read.sum"]) + num(d["dram__bytes_write.sum"])
            json.dump({"layer": d["layer"], "dram_bytes_per_launch": t, "source": f"profiles/conv_full_{tag}.json"}, open(os.path.join(P, "conv_tc_traffic.json"), "w"))

# 3. bench lines and test log
for src, dst in (("bench.json", f"bench_{tag}.json"), ("bench_reference.json", f"bench_reference_{tag}.json"), ("gpu_tests.log", f"gpu_tests_{tag}.log"),
                 ("ops.csv", f"ops_{tag}.csv"), ("plan.txt", f"conv_plan_{tag}.txt"), ("timeline_tail.txt", f"conv_timeline_{tag}.txt")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        open(os.path.join(P, dst), "w").write(open(p).read())
print(open(os.path.join(P, f"launches_{tag}.md")).read())
print(json.dumps(full, indent=1)[:3000])
