#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep (needs -lineinfo + --import-source on): warp instructions executed, lane use and stall
samples per CUDA source line.  usage: ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = None; hdr = None; rows = []
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0]: continue  # SASS rows have an empty line number
    try:
        rows.append((fname, int(r[0]), r[1].strip(), int(r[hdr.index("# Samples")] or 0), int(r[hdr.index("Instructions Executed")] or 0),
                     int(r[hdr.index("Thread Instructions Executed")] or 0)))
    except ValueError:
        pass
ti = sum(r[4] for r in rows); ts = sum(r[3] for r in rows)
print(f"total warp instructions {ti:,}  samples {ts:,}")
print("  exe%   smp%  lanes  file:line  source")
for f, ln, src, s, n, tn in sorted(rows, key=lambda r: -r[4])[:top]:
    print(f"{100*n/max(ti,1):6.2f} {100*s/max(ts,1):6.2f}  {tn/max(n,1):5.1f}  {f}:{ln}  {src[:110]}")
