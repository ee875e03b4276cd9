"""Farthest-point sampling on the resident mesh: n samples = n-1 multi-source solves + arg-max.
    python tools/run_fps.py [f=447] [n=64] [dtype=f32] [--ref]
--ref also times the reference's own farthest_point_sampling_ptp_gpu (oracle/_ref/libgproshan_ref_cuda_*.so, in a process
of its own: it calls cudaDeviceReset) on the same mesh and start sample, and compares the sample lists.
BENCH INFRASTRUCTURE (the --ref leg loads oracle/_ref through tests/ref_gpu_run.py)."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gproshan_b200 import api, meshgen as mg  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
f = int(args[0]) if len(args) > 0 else 447
n = int(args[1]) if len(args) > 1 else 64
dt = np.float64 if (len(args) > 2 and args[2] == "f64") else np.float32
mesh = mg.icosphere(f, dtype=dt)
ours = None
if "--ref" in sys.argv:
    api.set_option("newest", 1)  # the buffer the reference's arg-max reads (src/cuda/geodesics_ptp.cu:139-141)
with api.DeviceMesh(mesh, 0) as dm:
    for rep in range(2):
        samples = [0]
        t = time.perf_counter()
        md, secs = api.farthest_point_sampling_ptp_gpu(dm, samples, n)
        wall = time.perf_counter() - t
        ours = (samples, secs, wall)
        print(f"V={mesh.n_vertices} {dt.__name__}: {n} samples in {secs*1e3:.1f} ms device ({wall*1e3:.1f} ms wall), "
              f"{secs*1e3/(n-1):.2f} ms per sample, max_dist={md:.6f}, first={samples[:6]}", flush=True)
if "--ref" in sys.argv:
    out = os.path.join(tempfile.mkdtemp(), "fps.npz")
    env = dict(os.environ, REF_FPS_F=str(f))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_gpu_run.py"), "fps", "f64" if dt == np.float64 else "f32", str(n), "0", out],
                       env=env, capture_output=True, text=True, timeout=3000)
    if r.returncode != 0:
        print("reference FPS failed:", r.stderr[-1500:])
    else:
        d = np.load(out)
        same = int((np.array(ours[0]) == d["samples"]).sum()) if len(ours[0]) == d["samples"].size else -1
        print(f"reference farthest_point_sampling_ptp_gpu (its own CUDA code, sm_100a): {n} samples in {float(d['seconds'])*1e3:.1f} ms by its own "
              f"timer ({float(d['wall'])*1e3:.1f} ms wall incl. its host-side toplesets per sample); ours {ours[1]*1e3:.1f} ms -> "
              f"{float(d['seconds'])/ours[1]:.1f}x (timer) / {float(d['wall'])/ours[2]:.1f}x (wall); identical samples: {same} of {n} "
              f"(reference first={d['samples'][:6].tolist()}; on the symmetric icosphere the arg-max is a tie among many vertices, which its "
              f"FMA-contracted kernels break differently; the seeded noisy mesh of tests/test_gpu_vs_reference_cuda.py gives identical lists)", flush=True)
