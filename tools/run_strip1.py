import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import api, meshgen as mg
mesh = mg.grid(64, 3000, dtype=np.float64)
with api.DeviceMesh(mesh, 0) as dm:
    for _ in range(3):
        dm.geodesics([0]); print(dm.last_stats["ms_solve"], dm.last_stats["iterations"])
