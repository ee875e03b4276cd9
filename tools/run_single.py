"""Profiling driver: a few single-source solves on the C3 mesh (or a smaller icosphere), nothing else.
    python tools/run_single.py [f=1000] [n=3] [dtype=f64]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import api, meshgen as mg  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dt = np.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else np.float64
mesh = mg.icosphere(f, noise_sigma=0.2 * mg.mean_edge_icosphere(f), seed=12345, dtype=dt)
with api.DeviceMesh(mesh, 0) as dm:
    for _ in range(n):
        d, _, _ = dm.geodesics([0])
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in dm.last_stats.items()}, flush=True)
    print("checksum", int(d.view(np.int64 if dt == np.float64 else np.int32).astype(np.int64).sum()), flush=True)  # A/B builds must agree
