"""Extract DRAM traffic per launch from .ncu-rep captures into profiles/r1_traffic.json (read by bench.py).
    python tools/ncu_traffic.py key=path.ncu-rep [key=path ...]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "r1_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        tot += float(vals[i].replace(",", "")) * UNIT[units[i]]
    out[key] = tot
    print(key, tot)
json.dump(out, open(out_path, "w"), indent=1)
