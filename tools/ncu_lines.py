"""Attribute samples / executed instructions of an .ncu-rep to CUDA source lines (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [n_top=50]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= iE or r[0] == "":
        continue
    try:
        lines.append((r[0], r[1], int(r[iS] or 0), int(r[iE] or 0)))
    except ValueError:
        pass
ts, te = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
print("total samples", ts, "warp instructions", te)
for l in sorted(lines, key=lambda x: -x[3])[:ntop]:
    print(f"{l[0]:>5s} smp {100 * l[2] / ts:5.1f}% ins {100 * l[3] / te:5.1f}%  {l[1].strip()[:105]}")
