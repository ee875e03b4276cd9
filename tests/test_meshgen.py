"""CPU: synthetic mesh generators and the host CHE builder."""
import numpy as np
import pytest

from gproshan_b200 import meshgen as mg
from oracle_lib import NIL


def euler(mesh):
    he = mesh.n_half_edges
    borders = int((mesh.OT == NIL).sum())
    edges = (he - borders) // 2 + borders
    return mesh.n_vertices - edges + mesh.n_faces


@pytest.mark.parametrize("f", [1, 2, 3, 7, 16])
def test_icosphere(f):
    m = mg.icosphere(f)
    assert m.n_vertices == 10 * f * f + 2 and m.n_faces == 20 * f * f
    assert (m.OT != NIL).all() and euler(m) == 2
    assert np.allclose(np.linalg.norm(m.GT, axis=1), 1.0)
    deg = np.bincount(m.VT, minlength=m.n_vertices)
    assert (deg == 5).sum() == 12 and ((deg == 6) | (deg == 5)).all()
    # OT is an involution pairing a->b with b->a
    nxt = lambda he: 3 * (he // 3) + (he + 1) % 3
    he = np.arange(m.n_half_edges)
    assert np.array_equal(m.OT[m.OT], he)
    assert np.array_equal(m.VT[m.OT], m.VT[nxt(he)])
    # outward orientation
    tri = m.GT[m.VT.reshape(-1, 3)]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert (np.einsum("ij,ij->i", nrm, tri.mean(axis=1)) > 0).all()


def test_grid_and_torus_topology():
    g = mg.grid(317)
    assert g.n_vertices == 100489 and g.n_faces == 199712  # config C1
    assert euler(g) == 1
    t = mg.torus(40, 16)
    assert euler(t) == 0 and (t.OT != NIL).all()


def test_che_build_matches_oracle(oracle):
    for m in (mg.grid(9, 14), mg.torus(12, 7), mg.icosphere(4), mg.punch_hole(mg.grid(15), 7 * 15 + 7, 2)):
        OT, EVT, manifold = oracle.che_build(m.n_vertices, m.VT)
        assert np.array_equal(OT, m.OT) and np.array_equal(EVT, m.EVT)


def test_non_manifold_rejected():
    faces = np.array([0, 1, 2, 0, 1, 3, 0, 1, 4], dtype=np.uint32)  # directed edge 0->1 three times
    with pytest.raises(ValueError):
        mg.che_from_faces(np.zeros((5, 3)), faces)


def test_mt19937_known_answers():
    # first outputs of mt19937(5489) — the C++ standard's default-seeded engine
    assert list(mg.mt19937(5489, 3)) == [3499211612, 581869302, 3890346734]
    s = mg.random_sources(1024, 1024, 1998092, unique=True)
    assert len(set(s.tolist())) == s.size and s.max() < 1998092


def test_radial_noise_is_seeded():
    a, b = mg.icosphere(5, 1e-2, seed=12345), mg.icosphere(5, 1e-2, seed=12345)
    c = mg.icosphere(5, 1e-2, seed=1)
    assert np.array_equal(a.GT, b.GT) and not np.array_equal(a.GT, c.GT)
    r = np.linalg.norm(a.GT, axis=1)
    assert r.min() >= 0.99 and r.max() <= 1.01 and r.std() > 1e-3
