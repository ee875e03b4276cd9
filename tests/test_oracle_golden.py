"""CPU: pin the oracle (oracle/ptp_oracle.c) to the golden vectors generated from the reference itself
(tests/golden/make_golden.py). Bit-exact for tables and distances."""
import glob
import os

import numpy as np
import pytest

from gproshan_b200.meshgen import CheMesh

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def test_golden_files_present():
    assert len(FILES) >= 7


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_reference_golden(path, oracle):
    g = np.load(path)
    GT, faces, src, k = g["GT"], g["faces"], g["sources"], int(g["k"])
    OT, EVT, _ = oracle.che_build(GT.shape[0], faces)
    assert np.array_equal(OT, g["OT"]) and np.array_equal(EVT, g["EVT"])
    mesh = CheMesh(GT, faces, OT, EVT)
    top, srt, lim = oracle.compute_toplesets(mesh, src, k)
    assert np.array_equal(lim, g["limits"])
    assert np.array_equal(srt[:lim[-1]], g["sorted"])
    assert np.array_equal(top, g["toplesets"])
    for dt, tag, bits in ((np.float64, "f64", np.uint64), (np.float32, "f32", np.uint32)):
        d, _, _ = oracle.ptp_cpu(mesh.astype(dt), src, lim, srt)
        assert np.array_equal(d.view(bits), g["dist_" + tag].view(bits)), tag
