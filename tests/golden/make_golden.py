"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref, the unmodified gproshan CPU sources
compiled by oracle/Makefile). Run in the build container where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Each file holds the inputs (GT float64, faces, sources) and the reference's outputs: OT / EVT from
che::update_evt_ot_et, toplesets / sorted / limits from che::compute_toplesets, and the distances of
parallel_toplesets_propagation_cpu in both precisions (float meshes are the float64 coordinates rounded)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cases import fan_mesh  # noqa: E402
from gproshan_b200 import meshgen as mg  # noqa: E402
from oracle_lib import NIL, Reference  # noqa: E402

GOLDEN = [
    ("grid21_center", mg.grid(21), [10 * 21 + 10], NIL),
    ("ico6_noise", mg.icosphere(6, 1e-2, seed=12345), [0], NIL),
    ("torus24x10_multi_dup", mg.torus(24, 10), [5, 100, 233, 5], NIL),
    ("grid19_hole_multi", mg.punch_hole(mg.grid(19), 9 * 19 + 9, 2), [3, 340], NIL),
    ("fan14_rim", fan_mesh(14, True), [20], NIL),
    ("fan11_open", fan_mesh(11, False), [3], NIL),
    ("ico5_cap3", mg.icosphere(5), [7], 3),
]


def main():
    for name, mesh, src, k in GOLDEN:
        out = dict(GT=mesh.GT.astype(np.float64), faces=mesh.VT.copy(), sources=np.array(src, dtype=np.uint32),
                   k=np.uint32(k))
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            ref = Reference(dt)
            rc = ref.che(mesh.GT.astype(dt), mesh.VT)
            _, _, OT, EVT = rc.tables()
            top, srt, lim = rc.compute_toplesets(src, k)
            dist = rc.ptp_cpu(src, lim, srt)
            # the coalescence variant overruns its vertex buffer when duplicates make limits.back() > V
            # (src/geodesics_ptp.cpp:19,27-31) and drops faces under a level cap: only cross-check it otherwise
            dist_c = rc.ptp_cpu(src, lim, srt, coalescence=True) if (k == NIL and lim[-1] <= mesh.n_vertices) else dist
            assert np.array_equal(dist, dist_c, equal_nan=True), "reference plain vs coalescence CPU disagree"
            if tag == "f64":
                out.update(OT=OT, EVT=EVT, toplesets=top, sorted=srt[:lim[-1]].copy(), limits=lim)
            out["dist_" + tag] = dist
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, mesh.n_vertices, "vertices ->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
