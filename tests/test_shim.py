"""The C++ drop-in (gproshan_b200/shim): reference signatures, reference `che`, reference CPU PTP as the judge."""
import subprocess

import numpy as np
import pytest

from cases import assert_dist_parity, small_cases
from shim_lib import Shim, shim_available, shim_path

needs_shim = pytest.mark.skipif(not shim_available(), reason="shim not built (needs /root/reference at build time)")

# symbols a replacement library must define (f64 mangling, f32 swaps the trailing d for f)
EXPECTED = [
    "_ZN8gproshan34parallel_toplesets_propagation_gpuERKNS_9ptp_out_tEPNS_3cheERKSt6vectorIjSaIjEERKNS_11toplesets_tE",
    "_ZN8gproshan46parallel_toplesets_propagation_coalescence_gpuERKNS_9ptp_out_tEPNS_3cheERKSt6vectorIjSaIjEERKNS_11toplesets_tERKb",
    "_ZN8gproshan31farthest_point_sampling_ptp_gpuEPNS_3cheERSt6vectorIjSaIjEERdm",
]


@needs_shim
@pytest.mark.parametrize("dtype,last", [(np.float64, "d"), (np.float32, "f")])
def test_shim_exports_reference_symbols(dtype, last):
    out = subprocess.run(["nm", "-D", "--defined-only", shim_path(dtype)], capture_output=True, text=True, check=True).stdout
    names = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for e in EXPECTED[:2]:
        assert e in names, e
    assert EXPECTED[2] + last in names


CASES = [c for c in small_cases() if c[0] not in ("single_triangle_all_sources",)]


@needs_shim
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_shim_gpu_equals_reference_cpu(case, dtype):
    """parallel_toplesets_propagation_gpu (this repo, through the shim) vs parallel_toplesets_propagation_cpu
    (the reference, same process, same che, same toplesets)."""
    name, mesh, src = case
    sh = Shim(dtype)
    h, n_v = sh.che(mesh.GT, mesh.VT)
    try:
        for coal in (False, True):
            secs, dg, dc, cl = sh.gpu_vs_cpu(h, n_v, src, coalescence=coal, clusters=True)
            assert secs > 0
            assert_dist_parity(dg, dc, dtype, f"{name} coalescence={coal}")
            reached = np.isfinite(dg)
            assert (cl[reached] >= 1).all() and (cl[reached] <= len(src)).all()
        secs, d, srt = sh.geodesics(h, n_v, src)
        assert_dist_parity(d, dc, dtype, name + " geodesics_ptp_b200")
        assert srt[0] == src[0]
    finally:
        sh.destroy(h)
