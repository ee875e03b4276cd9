"""CPU: OFF reader / writer against the reference's own che_off (oracle/_ref) — config C1's stated input format."""
import numpy as np
import pytest

from gproshan_b200 import meshgen as mg
from gproshan_b200.off_io import read_off, write_off
from oracle_lib import Reference, ref_available

needs_ref = pytest.mark.skipif(not ref_available(np.float64), reason="oracle/_ref not built")


def test_roundtrip_is_exact_with_17_digits(tmp_path):
    m = mg.icosphere(7, 1e-2, seed=3)
    p = tmp_path / "m.off"
    write_off(p, m.GT, m.VT)
    xyz, faces = read_off(p)
    assert np.array_equal(xyz, m.GT) and np.array_equal(faces, m.VT)


def test_coff_noff_and_quads(tmp_path):
    (tmp_path / "c.off").write_text("COFF\n4 1 0\n0 0 0 255 0 0 255\n1 0 0 0 255 0 255\n1 1 0 0 0 255 255\n0 1 0 9 9 9 9\n4 0 1 2 3\n")
    xyz, faces = read_off(tmp_path / "c.off")
    assert xyz.shape == (4, 3) and faces.tolist() == [0, 1, 2, 3, 0, 2]  # (a b c) (d a c), src/che_off.cpp:68-76
    (tmp_path / "n.off").write_text("NOFF\n3 1 0\n0 0 0 0 0 1\n1 0 0 0 0 1\n0 1 0 0 0 1\n3 0 1 2\n")
    xyz, faces = read_off(tmp_path / "n.off")
    assert xyz.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0]] and faces.tolist() == [0, 1, 2]


@needs_ref
def test_reader_matches_reference_che_off(tmp_path):
    ref = Reference(np.float64)
    m = mg.punch_hole(mg.grid(13), 6 * 13 + 6, 1)
    p = tmp_path / "grid.off"
    write_off(p, m.GT, m.VT)
    rc = ref.read_off(p)
    GT, VT, OT, EVT = rc.tables()
    xyz, faces = read_off(p)
    assert np.array_equal(GT, xyz) and np.array_equal(VT, faces)
    assert np.array_equal(OT, m.OT) and np.array_equal(EVT, m.EVT)
    # quads: same split as the reference
    (tmp_path / "q.off").write_text("OFF\n6 2 0\n0 0 0\n1 0 0\n2 0 0\n0 1 0\n1 1 0\n2 1 0\n4 0 1 4 3\n4 1 2 5 4\n")
    rq = ref.read_off(tmp_path / "q.off")
    _, VTq, _, _ = rq.tables()
    assert np.array_equal(VTq, read_off(tmp_path / "q.off")[1])


@needs_ref
def test_reference_writer_is_readable_but_lossy(tmp_path):
    """che_off::write_file prints 6 significant digits: readable by read_off, equal to write_off(digits=6)."""
    ref = Reference(np.float64)
    m = mg.icosphere(3, 1e-2, seed=1)
    rc = ref.che(m.GT, m.VT)
    rc.write_off(tmp_path / "ref")
    xyz, faces = read_off(tmp_path / "ref.off")
    assert np.array_equal(faces, m.VT)
    assert not np.array_equal(xyz, m.GT) and np.allclose(xyz, m.GT, rtol=1e-5, atol=1e-6)
    write_off(tmp_path / "mine.off", m.GT, m.VT, digits=6)
    assert np.array_equal(read_off(tmp_path / "mine.off")[0], xyz)
