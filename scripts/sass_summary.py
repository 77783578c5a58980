#!/usr/bin/env python
"""profiles/<tag>_sass_summary.md from `cuobjdump -sass` of the in-tree library (runs on the CPU box): per kernel the instruction
count and the opcodes that carry the design (global / generic / shared accesses, atomics, fences, sleeps, barriers, FP64, MATCH),
plus the synchronisation instructions of the kernels that hand data between CTAs.

    python scripts/sass_summary.py r02"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "slam3d_b200", "libs3d_b200.so")], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"^void ", "", name).split("(")[0].replace("s3d::", "")
        cur = kernels.setdefault(name, [])  # static kernels of sort.cuh appear once per translation unit: identical copies
        if cur:
            cur = None  # a second copy: skip
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur is not None:
        cur.append(m.group(1).strip())


def count(ins, pred):
    return sum(1 for i in ins if pred(re.sub(r"^@!?U?P\d+\s+", "", i)))


cols = [("LDG", lambda i: i.startswith("LDG")), ("LD (generic)", lambda i: re.match(r"LD(\.|\s)", i) is not None), ("STG", lambda i: i.startswith("STG")),
        ("ST (generic)", lambda i: re.match(r"ST(\.|\s)", i) is not None), ("LDS", lambda i: i.startswith("LDS")), ("STS", lambda i: i.startswith("STS")),
        ("ATOMS", lambda i: i.startswith("ATOMS")), ("ATOMG/RED", lambda i: i.startswith("ATOMG") or i.startswith("REDG") or i.startswith("RED.")),
        ("MATCH", lambda i: i.startswith("MATCH")), ("MEMBAR", lambda i: i.startswith("MEMBAR")), ("CCTL.IVALL", lambda i: i.startswith("CCTL.IVALL")),
        ("NANOSLEEP", lambda i: i.startswith("NANOSLEEP")), ("BAR", lambda i: i.startswith("BAR")),
        ("DFMA+DMUL+DADD", lambda i: i.startswith("DFMA") or i.startswith("DMUL") or i.startswith("DADD"))]
out = [f"# {tag} — SASS summary of libs3d_b200.so (cuobjdump -sass, sm_100a; `scripts/sass_summary.py`): instruction counts and the opcodes that carry the design", "",
       "No tensor-core or TMA opcodes (UTC*MMA, LDTM, UTMALDG) are expected: nothing on this path is a dense contraction or a tiled copy (DESIGN.md 4).",
       "What is Blackwell-era here is the control structure: the persistent loop kernel hands data between CTAs with release atomics",
       "(`MEMBAR.ALL.GPU` + `ATOMG ... .STRONG.GPU`, no `CCTL.IVALL`), reads it with `LDG.E...STRONG.GPU` / `.CG`-class loads, sleeps with `NANOSLEEP`,",
       "and takes its arguments as a `__grid_constant__` block (`LDC`/`LDCU` from constant bank 0, plain `LDG`/`STG` through them); the radix",
       "sort and the centroid kernel chain their tiles through decoupled look-back words (`LDG.E.64.STRONG.GPU` / `STG.E.64.STRONG.GPU`, `NANOSLEEP`",
       "in the bounded spin) and rank keys with `MATCH.ANY`; `bbox_kernel`'s last CTA fences with `MEMBAR` (+ `CCTL.IVALL`: it re-reads the slot table).", "",
       "| kernel | SASS instructions | " + " | ".join(c for c, _ in cols) + " |", "|---|---|" + "---|" * len(cols)]
tensor = 0
for name, ins in kernels.items():
    if not ins:
        continue
    out.append(f"| {name} | {len(ins)} | " + " | ".join(str(count(ins, p)) for _, p in cols) + " |")
    tensor += count(ins, lambda i: i.startswith("UTC") or i.startswith("LDTM") or i.startswith("UTMA") or i.startswith("HMMA") or i.startswith("IMMA"))
out += ["", f"Tensor-core / TMEM / TMA opcodes in the library: {tensor}.", ""]
for k in ("gicp_loop_kernel", "sort_pass_kernel", "voxel_centroid_kernel"):
    ins = kernels.get(k, [])
    sync = [i for i in ins if re.search(r"STRONG|MEMBAR|NANOSLEEP|CCTL|ATOMG|REDG|MATCH", i)]
    out += [f"## {k}: the synchronisation instructions (all of them)", "", "```"] + sync + ["```", ""]
open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md"), "w").write("\n".join(out))
print("\n".join(out[:60]))
