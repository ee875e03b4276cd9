"""Grid barrier + one dependent load of team-written data: acquire poll + plain load vs relaxed poll + ld.cg."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import _lib
L = _lib.lib()
for ctas in (1, 112, 148):
    for block in (32, 768):
        r = [L.ptp_debug_barrier_ns(ctas, mode * 100000 + block, 20000) for mode in (0, 1, 2)]
        print(f"ctas {ctas:4d} block {block:4d}: barrier only {r[0]:7.1f} ns | acquire + plain load {r[1]:7.1f} ns | relaxed poll + ld.cg {r[2]:7.1f} ns", flush=True)
