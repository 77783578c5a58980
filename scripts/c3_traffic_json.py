#!/usr/bin/env python
"""profiles/<tag>_c3_traffic.json from the two config-3 launch lists of scripts/profile_round.sh (time + DRAM bytes per launch,
cold and warm): per leaf size, the kernels of ONE voxel-filter call (the second call at that leaf) with their times and DRAM
bytes, and the sums `bench.py --workload c3` reports as roofline.traffic.

    python scripts/c3_traffic_json.py r02"""
import csv, json, sys
tag = sys.argv[1]


def calls(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, rows = r, rows[i + 1:]
            break
    ki, mi, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    launches = {}
    for r in rows:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1)
        launches.setdefault(int(r[idi]), {"kernel": r[ki].split("(")[0].replace("void ", "")})[r[mi]] = float(r[vi].replace(",", "")) * scale
    order = [launches[i] for i in sorted(launches)]
    out, cur = [], []
    for l in order:  # a voxel-filter call starts with the bbox kernel
        if l["kernel"].startswith("bbox_kernel") and cur:
            out.append(cur); cur = []
        cur.append(l)
    out.append(cur)
    return out


res = {}
for name, path in (("cold", f"gpurun_out/c3_launches_{tag}.csv"), ("warm", f"gpurun_out/c3_launches_warm_{tag}.csv")):
    cs = calls(path)
    for leaf, idx in (("0.05", 1), ("0.1", 3), ("0.2", 5)):
        c = cs[idx]
        res.setdefault(f"leaf_{leaf}", {})[name] = {
            "launches": [{"kernel": l["kernel"], "us": round(l["gpu__time_duration.sum"], 2),
                          "dram_read_mb": round(l["dram__bytes_read.sum"] / 1e6, 2), "dram_write_mb": round(l["dram__bytes_write.sum"] / 1e6, 2)} for l in c],
            "sum_us": round(sum(l["gpu__time_duration.sum"] for l in c), 1),
            "dram_bytes": sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in c)}
res["capture"] = "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none [--cache-control none] python scripts/c3_voxel_only.py"
json.dump(res, open(f"profiles/{tag}_c3_traffic.json", "w"), indent=1)
for k, v in res.items():
    if k.startswith("leaf"):
        print(k, {n: (x["sum_us"], round(x["dram_bytes"] / 1e6, 1)) for n, x in v.items()})
