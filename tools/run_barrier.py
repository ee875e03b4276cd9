import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import _lib
L = _lib.lib()
for ctas in (1, 8, 32, 74, 148, 296):
    for block in (32, 256, 512, 1024):
        if ctas == 296 and block == 1024: continue
        print(f"ctas {ctas:4d} block {block:5d}: {L.ptp_debug_barrier_ns(ctas, block, 20000):8.1f} ns / barrier", flush=True)
for cs in (1, 2, 4, 8, 16):
    for block in (32, 256, 1024):
        print(f"cluster of {cs:2d} block {block:5d}: {L.ptp_debug_barrier_ns(-cs, block, 20000):8.1f} ns / cluster barrier", flush=True)
