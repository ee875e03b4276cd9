"""Farthest-point sampling on the resident mesh: n samples = n-1 multi-source solves + arg-max.
    python tools/run_fps.py [f=447] [n=64] [dtype=f32]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import api, meshgen as mg
f = int(sys.argv[1]) if len(sys.argv) > 1 else 447
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dt = np.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else np.float32
mesh = mg.icosphere(f, dtype=dt)
with api.DeviceMesh(mesh, 0) as dm:
    for rep in range(2):
        samples = [0]
        t = time.perf_counter()
        md, secs = api.farthest_point_sampling_ptp_gpu(dm, samples, n)
        wall = time.perf_counter() - t
        print(f"V={mesh.n_vertices} {dt.__name__}: {n} samples in {secs*1e3:.1f} ms device ({wall*1e3:.1f} ms wall), "
              f"{secs*1e3/(n-1):.2f} ms per sample, max_dist={md:.6f}, first={samples[:6]}", flush=True)
