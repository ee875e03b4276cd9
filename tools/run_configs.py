#!/usr/bin/env python
"""All five workloads of BASELINE.json at full size through the CUDA path, one JSON object on stdout
(profiles/r1_configs.json). bench.py's line carries C5 (batched) and C3 (single source); this adds C1, C2 and C4 and
times the reference's CPU PTP (oracle/_ref, all host threads) beside each one, with the bit-equality check where the CPU
run is affordable. BENCH INFRASTRUCTURE (uses oracle/_ref like bench.py's cpu_baseline leg).
    python tools/run_configs.py [--quick] [--no-cpu]"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gproshan_b200 import api, meshgen as mg  # noqa: E402


def cpu_solver(dtype):
    import oracle_lib as ol
    if ol.ref_available(dtype):
        ref = ol.Reference(dtype)
        return "reference", lambda mesh: ref.che_raw(mesh)
    return None, None


def run_one(name, mesh, src, clusters, reps, cpu):
    out = {"config": name, "V": mesh.n_vertices, "dtype": "f32" if mesh.GT.dtype == np.float32 else "f64", "sources": int(len(src)),
           "clusters": clusters}
    with api.DeviceMesh(mesh, 0) as dm:
        ms, wall = [], []
        for _ in range(reps + 1):
            t = time.perf_counter()
            dist, cl, _ = dm.geodesics(src, clusters=clusters)
            wall.append((time.perf_counter() - t) * 1e3)
            ms.append(dm.last_stats["ms_total"])
        st = dict(dm.last_stats)
        out.update(kernel=dm.last_kernel, ms_per_solve=statistics.median(ms[1:]), e2e_ms=statistics.median(wall[1:]),
                   levels=st["n_levels"], iterations=st["iterations"], vertex_updates=st["vertex_updates"],
                   relaxations=st["relaxations"], max_window=st["max_window"],
                   vertex_updates_per_s=st["vertex_updates"] / (statistics.median(ms[1:]) / 1e3))
    if clusters:
        out["clusters_labelled"] = bool(cl.min() >= 1 and cl.max() <= len(src))
    if cpu:
        kind, make = cpu_solver(mesh.GT.dtype)
        if kind:
            rc = make(mesh)
            s = np.ascontiguousarray(src, dtype=np.uint32)
            t = time.perf_counter()
            top, srt, lim = rc.compute_toplesets(s)
            t_top = time.perf_counter() - t
            t = time.perf_counter()
            want = rc.ptp_cpu(s, lim, srt)
            t_ptp = time.perf_counter() - t
            out["cpu_baseline"] = {"kind": kind, "cores": os.cpu_count(), "toplesets_ms": t_top * 1e3, "ptp_ms": t_ptp * 1e3,
                                   "ms_per_solve": (t_top + t_ptp) * 1e3}
            out["bit_equal_to_cpu"] = bool(np.array_equal(dist, want))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    q = a.quick
    res = []
    n = 80 if q else 317
    for dt in (np.float64, np.float32):
        g = mg.grid(n, dtype=dt)
        res.append(run_one(f"C1 grid {n}x{n}", g, [(n // 2) * n + n // 2], False, 5, not a.no_cpu))
    f = 60 if q else 316
    res.append(run_one(f"C2 icosphere f={f}", mg.icosphere(f, dtype=np.float32), [0], False, 5, not a.no_cpu))
    nu, nv = (378, 132) if q else (3780, 1323)
    for dt in (np.float64, np.float32):
        t = mg.torus(nu, nv, 1.0, 0.35, dtype=dt)
        src = mg.random_sources(7, 64, t.n_vertices)
        res.append(run_one(f"C4 torus {nu}x{nv}, 64 sources, Voronoi clusters", t, src, True, 2, not a.no_cpu and dt == np.float64))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
