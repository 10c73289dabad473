#!/usr/bin/env python
"""Turns the raw outputs of tools/profile_round.sh (gpurun_out/final_*) into the tracked summaries under profiles/.
Needs the `ncu` CLI (reads .ncu-rep files; no GPU)."""
import csv, io, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

def raw(rep):
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    caps = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        d = {k: {"unit": units[hdr.index(k)], "value": r[hdr.index(k)]} for k in KEEP if k in hdr}
        caps.setdefault(name.split("(")[0], d)
    return caps

# 1. launch list
src = os.path.join(G, "final_launches.csv")
lines = [l for l in open(src) if not l.startswith("==")]
open(os.path.join(P, f"launches_{TAG}.csv"), "w").writelines(lines)
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    k = r["Kernel Name"].split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"launches_{TAG}_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e\n")
    f.write("(cold-cache, serialised launches: compare shares, not absolutes)\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:60s} launches {n:4d}  total {ms:9.3f} ms  share {100*ms/tot:5.1f}%\n")

# 2. full captures
caps = {}
for rep, label in (("final_k_merge.ncu-rep", "k_merge_full_size"), ("final_reads.ncu-rep", "reads_full_size"), ("final_scan.ncu-rep", "reads_full_size")):
    pth = os.path.join(G, rep)
    if os.path.exists(pth):
        for name, d in raw(pth).items(): caps[f"{label}:{name}"] = d
json.dump({"note": "ncu --set full --clock-control none captures at the full bench size (4 x 2.5M records; reads: 262144 gets / 16384 "
                   "prefix scans); times under ncu replay are not bench values", "captures": caps},
          open(os.path.join(P, f"ncu_full_size_{TAG}.json"), "w"), indent=1)
km = next((d for k, d in caps.items() if "k_merge" in k), None)
if km:
    rd = float(km["dram__bytes_read.sum"]["value"]); wr = float(km["dram__bytes_write.sum"]["value"])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    rd *= scale[km["dram__bytes_read.sum"]["unit"]]; wr *= scale[km["dram__bytes_write.sum"]["unit"]]
    json.dump({"kernel": "k_merge", "dram_bytes_per_launch": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
               "source": f"profiles/ncu_full_size_{TAG}.json (ncu --set full, bench config 4 x 2.5M records)"},
              open(os.path.join(P, "k_merge_traffic.json"), "w"), indent=1)

# 3. bench line + phase shares
for a, b in (("final_bench.json", f"bench_{TAG}.json"), ("final_phases.txt", f"k_merge_phases_{TAG}.txt")):
    if os.path.exists(os.path.join(G, a)):
        open(os.path.join(P, b), "w").write(open(os.path.join(G, a)).read())
print(open(os.path.join(P, f"launches_{TAG}_summary.txt")).read())
print(json.dumps({k: {m: v["value"] for m, v in d.items() if m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active")} for k, d in caps.items()}, indent=1))
