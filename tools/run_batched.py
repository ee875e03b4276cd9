"""Profiling driver: batched single-source solves on an icosphere.
    python tools/run_batched.py [f=447] [n_sources=128] [reps=2] [dtype=f32]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import api, meshgen as mg  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 447
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dt = np.float64 if (len(sys.argv) > 4 and sys.argv[4] == "f64") else np.float32
mesh = mg.icosphere(f, dtype=dt)
src = mg.random_sources(1024, 1024, mesh.n_vertices, unique=True)[:nsrc]
import torch  # noqa: E402  (device buffer for the rows)
rows = torch.empty((nsrc, mesh.n_vertices), dtype=torch.float32 if dt == np.float32 else torch.float64, device="cuda")
with api.DeviceMesh(mesh, 0) as dm:
    for _ in range(reps):
        dm.solve_batched(src, rows_device_ptr=rows.data_ptr())
        st = dm.last_stats
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
              "sources/s", round(nsrc / (st["ms_total"] / 1e3), 1), flush=True)
    # bit-level checksum of the whole matrix: A/B builds must print the same number
    iv = rows.view(torch.int32 if dt == np.float32 else torch.int64)
    print("checksum", int(iv.to(torch.int64).sum().item()), int((iv.to(torch.int64) * (torch.arange(iv.shape[1], device="cuda") % 8191 + 1)).sum().item()), flush=True)
