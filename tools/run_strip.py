"""Latency floor: a long thin strip (tiny windows, thousands of levels) -> time per BFS level / PTP iteration."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import api, meshgen as mg
for nx, ny, dt in ((3, 3000, np.float64), (3, 3000, np.float32), (64, 3000, np.float64)):
    mesh = mg.grid(nx, ny, dtype=dt)
    with api.DeviceMesh(mesh, 0) as dm:
        for _ in range(3):
            dm.geodesics([0])
        st = dm.last_stats
        print(nx, ny, dt.__name__, "levels", st["n_levels"], "iters", st["iterations"],
              "bfs us/level %.2f" % (1e3 * st["ms_toplesets"] / st["n_levels"]),
              "sweep us/iter %.2f" % (1e3 * st["ms_solve"] / st["iterations"]), flush=True)
