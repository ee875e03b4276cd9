"""GPU parity at the sizes BASELINE.json states (configs 3, 4, 5), against the reference's own CPU PTP (oracle/_ref, the
reference compiled unmodified) where the question is distances, and against the oracle port where it is clusters (the
reference's CPU cluster code is racy). The CPU side costs ~3 s (C3), ~30 s (C4) and ~4 s per C5 row on a
16-thread host; run once per round on the GPU box."""
import numpy as np
import pytest

from cases import assert_dist_parity
from gproshan_b200 import api
from gproshan_b200 import meshgen as mg
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def cpu_ptp(mesh, src, oracle):
    """-> (distances, sorted, limits) from the reference CPU build when present, else the port"""
    src = np.ascontiguousarray(src, dtype=np.uint32)
    if ol.ref_available(mesh.GT.dtype):
        rc = ol.Reference(mesh.GT.dtype).che_raw(mesh)
        top, srt, lim = rc.compute_toplesets(src)
        return rc.ptp_cpu(src, lim, srt), srt, lim
    top, srt, lim = oracle.compute_toplesets(mesh, src)
    return oracle.ptp_cpu(mesh, src, lim, srt)[0], srt, lim


def test_config_c3_noisy_sphere_10m_double(oracle):
    """configs[2]: 10 000 002-vertex noisy icosphere, double, source = vertex 0: distances bit-equal, BFS order exact."""
    f = 1000
    m = mg.icosphere(f, noise_sigma=0.2 * mg.mean_edge_icosphere(f), seed=12345, dtype=np.float64)
    assert m.n_vertices == 10_000_002
    want, srt, lim = cpu_ptp(m, [0], oracle)
    with api.DeviceMesh(m, 0) as dm:
        got, _, s_gpu = dm.geodesics([0], want_sorted=True)
        st = dict(dm.last_stats)
        top, s2, l2 = dm.compute_toplesets([0], want_toplesets=False)
    assert np.array_equal(s_gpu, srt[:lim[-1]]) and np.array_equal(s2, srt[:lim[-1]]) and np.array_equal(l2, lim)
    assert_dist_parity(got, want, np.float64, "C3")
    assert st["n_levels"] == len(lim) - 1 and st["n_reached"] == lim[-1]


def test_config_c4_torus_5m_64_sources_clusters_double(oracle):
    """configs[3] at full size: torus 3780 x 1323 (5 000 940 vertices), 64 sources mt19937(7) % V, Voronoi clusters."""
    m = mg.torus(3780, 1323, 1.0, 0.35, dtype=np.float64)
    assert m.n_vertices == 5_000_940
    src = mg.random_sources(7, 64, m.n_vertices)
    t0, s0, l0 = oracle.compute_toplesets(m, src)
    want, want_cl, st = oracle.ptp_cpu(m, src, l0, s0, clusters=True)
    with api.DeviceMesh(m, 0) as dm:
        got, cl, srt = dm.geodesics(src, clusters=True, want_sorted=True)
        stats = dict(dm.last_stats)
    assert np.array_equal(srt, s0[:l0[-1]])
    assert_dist_parity(got, want, np.float64, "C4")
    assert np.array_equal(cl, want_cl) and cl.min() >= 1 and cl.max() <= 64
    assert stats["iterations"] == st["iterations"] and stats["vertex_updates"] == st["vertex_updates"]


def test_config_c5_rows_2m_float(oracle):
    """configs[4]: rows of the 1024-source distance matrix on the 1 998 092-vertex icosphere, float: a sample of rows
    (first, middle, last source of the job, solved in one batched call with the rest of a wave) against the CPU."""
    m = mg.icosphere(447, dtype=np.float32)
    assert m.n_vertices == 1_998_092
    srcs = mg.random_sources(1024, 1024, m.n_vertices, unique=True)
    pick = [0, 511, 1023]
    batch = np.ascontiguousarray(np.concatenate([srcs[pick], srcs[1:150]]))  # a full wave of CTAs + elastic helpers
    with api.DeviceMesh(m, 0) as dm:
        rows = dm.solve_batched(batch)
    for k, b in enumerate(pick):
        want, _, _ = cpu_ptp(m, [srcs[b]], oracle)
        assert_dist_parity(rows[k], want, np.float32, f"C5 row of source #{b}")
    # size-independent properties over the whole wave: zero exactly at the source, symmetric-ish triangle bound, finite
    assert np.isfinite(rows).all()
    assert all(rows[k, batch[k]] == 0 for k in range(batch.size))
    assert (rows.max(axis=1) < 3.6).all() and (rows.max(axis=1) > 3.0).all()  # unit sphere: antipode at ~pi
