"""ANALYSIS (test infrastructure): lane- and warp-level potential of a two-sided causal skip; see two_sided_skip.c.
    python tests/analysis/two_sided_skip.py [f=100] [noise=0]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gproshan_b200 import meshgen as mg  # noqa: E402
import oracle_lib as ol  # noqa: E402

so = os.path.join(HERE, "_two_sided_skip.so")
subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "two_sided_skip.c"), "-lm"], check=True)
L = C.CDLL(so)
f = int(sys.argv[1]) if len(sys.argv) > 1 else 100
noise = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
mesh = mg.icosphere(f, noise_sigma=noise * mg.mean_edge_icosphere(f), seed=7, dtype=np.float32) if noise else mg.icosphere(f, dtype=np.float32)
orc = ol.Oracle()
src = np.array([12345 % mesh.n_vertices], dtype=np.uint32)
tl, srt, lim = orc.compute_toplesets(mesh, src)
inv = np.full(mesh.n_vertices, 0xFFFFFFFF, dtype=np.uint32)
n = int(lim[-1])
inv[srt[:n][::-1]] = np.arange(n - 1, -1, -1, dtype=np.uint32)
out = (C.c_uint64 * 8)()
p = lambda a, t=C.c_uint32: a.ctypes.data_as(C.POINTER(t))
GT = np.ascontiguousarray(mesh.GT, dtype=np.float32)
L.analyze2_f32(mesh.n_vertices, p(GT, C.c_float), p(mesh.VT), p(mesh.OT), p(mesh.EVT), p(src), 1, p(lim), lim.size, p(srt), p(inv), out)
o = list(out)
print("V", mesh.n_vertices, "relaxations", o[0], "triangle slots", o[1])
print("lane level: evaluated with the causal skip %.1f %%, with the two-sided skip as well %.1f %%" % (100 * o[2] / o[1], 100 * o[3] / o[1]))
print("warp level: evaluated with the causal skip %.1f %%, with the two-sided skip as well %.1f %%" % (100 * o[5] / o[4], 100 * o[6] / o[4]))
print("violations (rule fired, p < cur):", o[7])
ex = (C.c_uint64 * 4)()
L.analyze2_extra(ex)
print("of the %d triangles still evaluated: %.1f %% return p < cur, %.1f %% have both corners below thr" % (ex[0], 100 * ex[1] / max(ex[0], 1), 100 * ex[2] / max(ex[0], 1)))
