"""DRAM traffic of a profiler range from an `ncu ... --page raw --csv` export -> profiles/r2_traffic.json (read by bench.py
for `roofline.traffic`). Each entry records the bytes, the commit the capture was taken on and the command.
    python tools/range_traffic.py <key> <raw.csv> "<command that produced the capture>" ["<what it covers>"]
Capture recipe (on the GPU box; the `profile_range` option brackets a solve / a batched call with cudaProfilerStart/Stop):
    PTP_PROFILE_RANGE=1 ncu --replay-mode app-range --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,\\
        lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,... \\
        -o /tmp/rep python tools/run_batched.py 447 1024 1          # or tools/run_single.py 1000 2
    ncu -i /tmp/rep.ncu-rep --page raw --csv > gpurun_out/<name>_raw.csv"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
key, path, command = sys.argv[1], sys.argv[2], sys.argv[3]
what = sys.argv[4] if len(sys.argv) > 4 else ""
rows = list(csv.reader(open(path)))
hdr, units, first = rows[0], rows[1], rows[2]
tot = sum(float(first[hdr.index(k)].replace(",", "")) * UNIT[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
out_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out[key] = {"bytes": tot, "commit": commit, "command": command, "what": what}
json.dump(out, open(out_path, "w"), indent=1)
print(key, tot)
