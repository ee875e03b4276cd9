"""Helper of tests/test_gpu_variants.py: the kernel selection of libptp_b200 is read from the environment once per
process (PTP_FUSED, PTP_GEO, PTP_CLUSTER, PTP_STAGE, PTP_ELASTIC), so every variant is checked in a process of its
own. Runs single-source, multi-source + clusters and batched solves on small meshes and compares them bit for bit
with the CPU oracle. Prints "OK <kernel of the last single solve>" or raises."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from cases import assert_dist_parity  # noqa: E402
from gproshan_b200 import api, meshgen as mg  # noqa: E402
from oracle_lib import Oracle  # noqa: E402


def main():
    orc = Oracle()
    kernel = ""
    for dtype in (np.float32, np.float64):
        for name, m, src in (("icosphere", mg.icosphere(40, noise_sigma=0.02, dtype=dtype), [7]),
                             ("torus-wide-windows", mg.torus(240, 120).astype(dtype), [17, 5000]),
                             ("holes", mg.punch_hole(mg.grid(60, dtype=dtype), 1830, 3), [0, 3599, 0])):
            t0, s0, l0 = orc.compute_toplesets(m, src)
            want, want_cl, st = orc.ptp_cpu(m, src, l0, s0, clusters=True)
            with api.DeviceMesh(m, 0) as dm:
                got, _, srt = dm.geodesics(src, want_sorted=True)
                kernel = dm.last_kernel
                stats = dict(dm.last_stats)
                got_c, cl, _ = dm.geodesics(src, clusters=True)
                rows = dm.solve_batched(np.array(src[:2], dtype=np.uint32))
            assert np.array_equal(srt, s0[:l0[-1]]), name
            assert_dist_parity(got, want, dtype, name)
            assert_dist_parity(got_c, want, dtype, name + " (clusters)")
            assert np.array_equal(cl, want_cl), name
            assert stats["iterations"] == st["iterations"] and stats["vertex_updates"] == st["vertex_updates"], name
            for b, s in enumerate(src[:2]):
                tb, sb, lb = orc.compute_toplesets(m, [s])
                assert_dist_parity(rows[b], orc.ptp_cpu(m, [s], lb, sb)[0], dtype, f"{name} row {b}")
    print("OK", kernel)


if __name__ == "__main__":
    main()
