"""Every selectable kernel variant of the single-solve and batched paths against the oracle (bit-equal), each in a
process of its own because the selection is read from the environment once; plus the second-baseline tool."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

VARIANTS = [
    ("default", {}, "k_sweep_streamed"),
    ("one-cluster-launch", {"PTP_FUSED": "4"}, "k_geodesics_cluster"),
    ("cluster-16", {"PTP_CLUSTER": "16"}, None),           # non-portable cluster size: may fall back, must stay correct
    ("cluster-staged", {"PTP_FUSED": "4", "PTP_STAGE": "1"}, None),
    ("two-team", {"PTP_FUSED": "1"}, "k_geodesics_fused"),
    ("two-team-staged", {"PTP_FUSED": "1", "PTP_STAGE": "2"}, "k_geodesics_fused"),
    ("three-launches", {"PTP_FUSED": "0"}, "k_solve_grid"),
    ("three-launches-unstaged", {"PTP_FUSED": "0", "PTP_STAGE": "0"}, "k_solve_grid"),
    ("geometry-table", {"PTP_GEO": "1", "PTP_GEO_SINGLE": "1"}, None),
    ("no-elastic", {"PTP_ELASTIC": "0"}, None),
    ("no-causal-skip", {"PTP_CAUSAL": "0"}, None),
    ("no-short-sign-test", {"PTP_SIGN_SHORT": "0"}, None),
    ("no-two-sided-skip", {"PTP_TWO_SIDED": "0"}, None),
    ("batched-teams-of-4", {"PTP_TEAM": "4"}, None),
    ("batched-teams-of-37-no-causal", {"PTP_TEAM": "37", "PTP_CAUSAL": "0"}, None),
    ("newest-buffer-off-explicit", {"PTP_NEWEST": "0"}, None),
]


@pytest.mark.parametrize("name,env,kernel", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_variant_bit_equal(name, env, kernel):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(HERE, "variant_check.py")], env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.startswith("OK"), last
    if kernel:
        assert kernel in last, last


def test_reference_gpu_tool_runs():
    """tools/ref_gpu_bench.py (the reference's own CUDA PTP, built under oracle/_ref) on the small workload: finite
    times, and its distances within 1e-2 of its own CPU path (it is compiled with FMA contraction: not bit-equal)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgproshan_ref_cuda_f64.so")):
        pytest.skip("reference CUDA build not present (make -C oracle refgpu needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_bench.py"), "--workload", "c3", "--quick"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    s = line["solves"][0]
    assert s["gpu_ms"] > 0 and s["max_rel_err_vs_reference_cpu"] < 1e-2


def test_report_harness(tmp_path):
    """gproshan's test_geodesics output for the PTP GPU arm (gproshan_b200/report.py) on two small synthetic meshes"""
    import numpy as np
    sys.path.insert(0, ROOT)
    from gproshan_b200 import meshgen as mg, report
    s = mg.icosphere(30, dtype=np.float32)
    g = mg.grid(40)
    res = report.run([("sphere30", s, report.analytic_exact("sphere", s, 0)), ("grid40", g, report.analytic_exact("plane", g, 0))],
                     str(tmp_path), n_test=2, fps_counts=(2, 4))
    assert 0 < res[0]["error_pct"] < 5 and 0 < res[1]["error_pct"] < 5 and res[0]["seconds"] > 0
    for f in ("ptp_results.tex", "ptp_results_double.tex", "sphere30.deg", "sphere30_toplesets.dist",
              "sphere30_toplesets_sorted.dist", "sphere30.fps", "grid40.deg", "sphere30_error.iter", "grid40_error_double.iter"):
        assert (tmp_path / f).stat().st_size > 0, f
    assert sum(int(l.split()[1]) for l in open(tmp_path / "sphere30_toplesets.dist")) == s.n_vertices
