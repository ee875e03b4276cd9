"""Measurement tool behind the barrier numbers quoted in DESIGN.md §4 / profiles/README.md (ptp_debug_barrier_ns):
  1. ns per fused grid barrier (arrive + reduce + poll) for several team sizes and CTA widths,
  2. ns per hardware cluster barrier for cluster sizes 1..16 (the BFS team of the single solve),
  3. grid barrier + one dependent load of team-written data: acquire poll + plain load vs relaxed poll + ld.cg.
    python tools/run_barrier.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import _lib  # noqa: E402

L = _lib.lib()
for ctas in (1, 8, 32, 74, 148, 296):
    for block in (32, 256, 512, 1024):
        if ctas == 296 and block == 1024:
            continue
        print(f"ctas {ctas:4d} block {block:5d}: {L.ptp_debug_barrier_ns(ctas, block, 20000):8.1f} ns / barrier", flush=True)
for cs in (1, 2, 4, 8, 16):
    for block in (32, 256, 1024):
        print(f"cluster of {cs:2d} block {block:5d}: {L.ptp_debug_barrier_ns(-cs, block, 20000):8.1f} ns / cluster barrier", flush=True)
for ctas in (1, 112, 148):
    for block in (32, 768):
        r = [L.ptp_debug_barrier_ns(ctas, mode * 100000 + block, 20000) for mode in (0, 1, 2)]
        print(f"ctas {ctas:4d} block {block:4d}: barrier only {r[0]:7.1f} ns | acquire + plain load {r[1]:7.1f} ns | "
              f"relaxed poll + ld.cg {r[2]:7.1f} ns", flush=True)
