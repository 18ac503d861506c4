#!/usr/bin/env python
"""Per-CUDA-source-line hot spots of an .ncu-rep (needs -lineinfo + --import-source on).

    python profiles/ncu_source_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, out = "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        try:
            out.append((int(r[si] or 0), int(r[ei] or 0), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot_s = sum(o[0] for o in out) or 1
tot_e = sum(o[1] for o in out) or 1
print(f"# {rep}: {tot_s} samples, {tot_e} warp-instructions over {len(out)} source lines")
print("# %samples  %instr   file:line  source")
for s, e, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * s / tot_s:6.2f}  {100 * e / tot_e:6.2f}  {f}:{ln:<5d} {src}")
