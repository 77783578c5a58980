#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep + a launch list CSV into the markdown summary committed under profiles/."""
import collections, csv, subprocess, sys

tag = sys.argv[1]          # e.g. r01e
launch_csv = sys.argv[2]
reps = sys.argv[3:]

def launch_table(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| {k} | {n} | {t:.1f} | {t/n:.1f} | {100*t/tot:.1f}% |")
    return "\n".join(out), tot

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    h, units, v = r[0], r[1], r[2]
    name = v[h.index("Kernel Name")] if "Kernel Name" in h else rep
    rows = [f"| {w} | {v[h.index(w)]} | {units[h.index(w)]} |" for w in WANT if w in h]
    return name, "\n".join(["| metric | value | unit |", "|---|---|---|"] + rows)

tbl, tot = launch_table(launch_csv)
print(f"# {tag} — ncu evidence\n")
print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES)\n")
print(tbl)
print(f"\ntotal {tot:.0f} us\n")
for rep in reps:
    name, m = raw_metrics(rep)
    print(f"## {name.split('(')[0]} (`ncu --set full --clock-control none --import-source on`, one launch)\n")
    print(m)
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/_sass.csv", "w").write(sass)
    top = subprocess.run([sys.executable, "scripts/ncu_top.py", "/tmp/_sass.csv", "8"], capture_output=True, text=True).stdout
    print("\n```\n" + top + "```\n")
