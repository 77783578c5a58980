#!/usr/bin/env python
"""profiles/<tag>_traffic.json from `ncu --set full` captures: DRAM bytes (read + write) per launch of the dominant kernels,
divided by the work units of that launch, so that bench.py can scale `roofline.traffic` to its own launch size.

    python scripts/traffic_json.py r02 gpurun_out/knn_cov_kernel_r02.ncu-rep:knn_cov_kernel:query gpurun_out/gicp_search_kernel_first_r02.ncu-rep:gicp_search_kernel:point-pass

The work units of a captured launch are read from the launch itself: grid size x block size is an upper bound, so the capture
command's workload (pairs, filtered points) is passed on the command line as --units <kernel>=<n>."""
import csv, json, subprocess, sys

tag = sys.argv[1]
units = {}
specs = []
for a in sys.argv[2:]:
    if a.startswith("--units="):
        k, v = a[len("--units="):].split("=")
        units[k] = float(v)
    else:
        specs.append(a.split(":"))
out = {}
for rep, kernel, unit in specs:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    h, u, v = r[0], r[1], r[2]
    def val(name):
        x = float(v[h.index(name)].replace(",", ""))
        un = u[h.index(name)]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(un, 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    n = units.get(kernel)
    out[kernel] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "units_in_launch": n, "unit": unit,
                   "dram_bytes_per_unit": (rd + wr) / n if n else None,
                   "lts_hit_rate_pct": float(v[h.index("lts__t_sector_hit_rate.pct")]) if "lts__t_sector_hit_rate.pct" in h else None,
                   "l1_hit_rate_pct": float(v[h.index("l1tex__t_sector_hit_rate.pct")]) if "l1tex__t_sector_hit_rate.pct" in h else None,
                   "capture": rep.split("/")[-1] + ", ncu --set full --clock-control none (cache flushed before the launch)"}
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
