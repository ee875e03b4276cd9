"""ANALYSIS (test infrastructure): potential of evaluating only the triangles with a changed corner in the change-driven
sweep; see changed_corners.c.   python tests/analysis/changed_corners.py [f=100]"""
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

so = os.path.join(HERE, "_changed_corners.so")
subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "changed_corners.c"), "-lm"], check=True)
L = C.CDLL(so)
f = int(sys.argv[1]) if len(sys.argv) > 1 else 100
mesh = mg.icosphere(f, dtype=np.float32)
orc = ol.Oracle()
src = np.array([12345 % mesh.n_vertices], dtype=np.uint32)
tl, srt, lim = orc.compute_toplesets(mesh, src)
u32p, fp = C.POINTER(C.c_uint32), C.POINTER(C.c_float)
out = (C.c_uint64 * 10)()
p = lambda a, t=C.c_uint32: a.ctypes.data_as(C.POINTER(t))
GT = np.ascontiguousarray(mesh.GT, dtype=np.float32)
L.analyze_f32(mesh.n_vertices, p(GT, C.c_float), p(mesh.VT), p(mesh.OT), p(mesh.EVT), p(src), 1, p(lim), lim.size, p(srt), out)
o = list(out)
print("V", mesh.n_vertices, "vertex-updates", o[3], "relaxations", o[0], "(%.1f %%)" % (100 * o[0] / o[3]))
print("triangles", o[1], "with a changed corner", o[2], "(%.1f %%)" % (100 * o[2] / o[1]), "| causal-pass", o[6], "(%.1f %%)" % (100 * o[6] / o[1]),
      "| both", o[5], "(%.1f %%)" % (100 * o[5] / o[1]))
print("restricted-minimum mismatches", o[4], "skip mismatches", o[7])
print("two-sided bound: additionally skippable among causal-pass", o[8], "(%.1f %% of all triangles)" % (100 * o[8] / o[1]), "violations", o[9])
