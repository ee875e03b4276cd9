"""Static SASS statistics of one kernel of a built library (a cheap proxy before spending GPU time):
    python tools/sass_stats.py <lib.so> <substring of the mangled kernel name> [--dump out.sass]
Prints instruction count, opcode histogram head, local-memory traffic instructions and MUFU / CALL counts."""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    n = 0
    for line in b.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2).split(".")[0]] += 1
            n += 1
    print(name[:110])
    print("  instructions", n, " LDL", ops["LDL"], "STL", ops["STL"], "MUFU", ops["MUFU"], "CALL", ops["CALL"], "FCHK", ops["FCHK"],
          "BSSY", ops["BSSY"], "BRA", ops["BRA"], "MOV", ops["MOV"] + ops["IMAD"])
    print("  ", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
    if "--dump" in sys.argv:
        open(sys.argv[sys.argv.index("--dump") + 1], "w").write(b)
