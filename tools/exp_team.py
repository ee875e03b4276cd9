"""Experiment driver: batched C5-style solves for several team sizes (CTAs per solve) of the batched path.
    python tools/exp_team.py [f=447] [n_sources=296] [teams=1,2,4,...] [maps=1]
Prints sources/s per configuration and checks that every configuration returns the same bits as the first one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import api, meshgen as mg  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 447
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 296
teams = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,4,8,16,37").split(",")]
maps = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "1").split(",")]
dt = np.float64 if (len(sys.argv) > 5 and sys.argv[5] == "f64") else np.float32
import torch  # noqa: E402

mesh = mg.icosphere(f, dtype=dt)
src = mg.random_sources(1024, 1024, mesh.n_vertices, unique=True)[:nsrc]
tdt = torch.float32 if dt == np.float32 else torch.float64
rows = torch.empty((nsrc, mesh.n_vertices), dtype=tdt, device="cuda")
ref = None
with api.DeviceMesh(mesh, 0) as dm:
    for mp in maps:
        for t in teams:
            api.set_option("team", t)
            rows.zero_()
            best = None
            for rep in range(2):
                dm.solve_batched(src, rows_device_ptr=rows.data_ptr())
                st = dm.last_stats
                best = st["ms_total"] if best is None else min(best, st["ms_total"])
            same = None
            if ref is None:
                ref = rows.clone()
            else:
                same = bool(torch.equal(ref, rows))
            print(f"team={t:3d} map={mp} kernel={dm.last_kernel:28s} ms={best:9.2f} sources/s={nsrc / (best / 1e3):8.1f} "
                  f"bfs+layout={st['ms_toplesets']:.2f} sweep={st['ms_solve']:.2f} relax={st['relaxations']} same_bits={same} "
                  f"dev_GB={dm.device_bytes / 1e9:.1f}", flush=True)
