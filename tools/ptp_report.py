#!/usr/bin/env python
"""gproshan's `test_geodesics` output for the PTP GPU arm (see gproshan_b200/report.py).
    python tools/ptp_report.py out_dir [--off DIR --exact DIR name ...] [--synthetic] [--n-test 10] [--f32]
--synthetic runs the synthetic meshes of BASELINE.json's small configs (grid 317x317, icosphere f=316) with analytic
exact distances; otherwise every <name> is read from DIR/<name>.off with DIR2/<name>.exact like the reference."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gproshan_b200 import meshgen as mg, off_io, report  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out_dir")
    ap.add_argument("names", nargs="*")
    ap.add_argument("--off", default=".")
    ap.add_argument("--exact", default=".")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--n-test", type=int, default=10)
    ap.add_argument("--f32", action="store_true")
    a = ap.parse_args()
    dt = np.float32 if a.f32 else np.float64
    meshes = []
    if a.synthetic:
        g = mg.grid(317, dtype=dt)
        meshes.append(("grid317", g, report.analytic_exact("plane", g, 0)))
        s = mg.icosphere(316, dtype=dt)
        meshes.append(("icosphere316", s, report.analytic_exact("sphere", s, 0)))
    for name in a.names:
        xyz, faces = off_io.read_off(os.path.join(a.off, name + ".off"), dtype=dt)
        m = mg.che_from_faces(xyz, faces)
        meshes.append((name, m, report.load_exact_geodesics(os.path.join(a.exact, name + ".exact"), m.n_vertices)))
    for r in report.run(meshes, a.out_dir, n_test=a.n_test):
        print(r)


if __name__ == "__main__":
    main()
