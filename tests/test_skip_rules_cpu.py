"""CPU checks of the two exact work-skipping rules of the batched sweep, on the oracle's update_step (plain C mirrors of
the rules in tests/analysis/*.c, a whole PTP run each):
  * a triangle whose two neighbour values are bit-identical to those of the vertex's previous relaxation cannot lower it
    (the basis of the change-driven relaxation: restricted minimum == full minimum, skipped vertices already hold their result),
  * whenever the two-sided causal skip fires, update_step does not return less than the vertex's value."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from gproshan_b200 import meshgen as mg

HERE = os.path.dirname(os.path.abspath(__file__))


def _build(name):
    so = os.path.join(HERE, "analysis", f"_{name}.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "analysis", f"{name}.c"), "-lm"], check=True)
    return C.CDLL(so)


def _meshes():
    yield "icosphere", mg.icosphere(24, dtype=np.float32), [100]
    yield "noisy icosphere", mg.icosphere(20, noise_sigma=0.2 * mg.mean_edge_icosphere(20), seed=3, dtype=np.float32), [7, 900]
    yield "grid", mg.grid(48, dtype=np.float32), [0]


def _p(a, t=C.c_uint32):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.mark.parametrize("name,mesh,src", list(_meshes()), ids=[m[0] for m in _meshes()])
def test_skip_rules_hold_over_a_whole_run(name, mesh, src):
    orc = ol.Oracle()
    src = np.asarray(src, dtype=np.uint32)
    _, srt, lim = orc.compute_toplesets(mesh, src)
    n = int(lim[-1])
    inv = np.full(mesh.n_vertices, 0xFFFFFFFF, dtype=np.uint32)
    inv[srt[:n][::-1]] = np.arange(n - 1, -1, -1, dtype=np.uint32)
    GT = np.ascontiguousarray(mesh.GT, dtype=np.float32)
    args = (mesh.n_vertices, _p(GT, C.c_float), _p(mesh.VT), _p(mesh.OT), _p(mesh.EVT), _p(src), src.size, _p(lim), lim.size, _p(srt))

    out = (C.c_uint64 * 10)()
    _build("changed_corners").analyze_f32(*args, out)
    assert out[0] > 0 and out[4] == 0 and out[7] == 0, list(out)   # restricted-minimum / skipped-vertex mismatches

    out = (C.c_uint64 * 8)()
    _build("two_sided_skip").analyze2_f32(*args, _p(inv), out)
    assert out[1] > 0 and out[7] == 0, list(out)                   # violations of the two-sided rule
    assert out[3] < out[2]                                         # and the rule does fire
