"""The C5 distance-matrix job through the NATIVE multi-GPU entry (ptp_solve_batched_multi_f32: one process, one host thread
per device, rows gathered on device 0 with NCCL) on every visible GPU. Prints one JSON line.
    python tools/run_multi.py [f=447] [n_sources=1024] [reps=2] [host|device]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import api, meshgen as mg  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 447
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
where = sys.argv[4] if len(sys.argv) > 4 else "device"
import torch  # noqa: E402

G = api.device_count()
mesh = mg.icosphere(f, dtype=np.float32)
src = mg.random_sources(1024, 1024, mesh.n_vertices, unique=True)[:nsrc]
meshes = [api.DeviceMesh(mesh, d) for d in range(G)]
rows_dev = torch.empty((nsrc, mesh.n_vertices), dtype=torch.float32, device="cuda:0") if where == "device" else None
rows_host = torch.empty((nsrc, mesh.n_vertices), dtype=torch.float32, pin_memory=True).numpy() if where == "host" else None
best = None
for _ in range(reps + 1):
    t = time.perf_counter()
    if where == "device":
        api.solve_batched_multi(meshes, src, rows_device_ptr=rows_dev.data_ptr())
        torch.cuda.synchronize()
    else:
        api.solve_batched_multi(meshes, src, rows=rows_host)
    dt = time.perf_counter() - t
    best = dt if best is None else min(best, dt)
st = meshes[0].last_stats
ref = meshes[0].solve_batched(src[:3])
got = rows_dev[:3].cpu().numpy() if where == "device" else rows_host[:3]
print(json.dumps({"entry": "ptp_solve_batched_multi_f32", "gpus": G, "rows": where, "sources": nsrc, "V": mesh.n_vertices,
                  "seconds": best, "sources_per_s": nsrc / best, "slowest_device_kernel_ms": st["ms_solve"], "wall_ms_in_call": st["ms_total"],
                  "first_rows_equal_single_device": bool(np.array_equal(ref, got))}))
for m in meshes:
    m.close()
