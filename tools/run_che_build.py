"""CHE construction (OT/EVT from faces): device build vs the reference's che constructor (update_evt_ot_et).
    python tools/run_che_build.py [f=1000] [--ref]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gproshan_b200 import api, meshgen as mg
f = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1000
t = time.perf_counter(); m = mg.icosphere(f); t_host = time.perf_counter() - t
print(f"icosphere f={f}: V={m.n_vertices} H={m.n_half_edges}; host generator + OpenMP che build {t_host:.2f}s", flush=True)
for _ in range(3):
    t = time.perf_counter(); OT, EVT, mf, ms = api.che_build(m.VT, m.n_vertices); wall = time.perf_counter() - t
    print(f"device che_build: {ms:.2f} ms device, {wall*1e3:.1f} ms incl. H2D/D2H; manifold={mf}; equal to host tables: {np.array_equal(OT, m.OT) and np.array_equal(EVT, m.EVT)}", flush=True)
if "--ref" in sys.argv:
    from oracle_lib import Reference
    ref = Reference(np.float64)
    t = time.perf_counter(); rc = ref.che(m.GT, m.VT); t_ref = time.perf_counter() - t
    _, _, OTr, EVTr = rc.tables()
    print(f"reference che(vertices, faces) constructor (update_evt_ot_et + eht + bt): {t_ref:.2f} s; tables equal: {np.array_equal(OT, OTr) and np.array_equal(EVT, EVTr)}", flush=True)
