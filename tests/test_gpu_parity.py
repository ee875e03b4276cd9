"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Distances: bit-equal (stated tolerance 1e-5 f32 / 1e-10 f64 checked first). Toplesets / sorted / limits: bit-exact."""
import numpy as np
import pytest

from cases import assert_dist_parity, small_cases
from gproshan_b200 import api
from gproshan_b200 import meshgen as mg
from oracle_lib import NIL

pytestmark = pytest.mark.gpu

CASES = small_cases()
IDS = [c[0] for c in CASES]


@pytest.fixture(scope="module")
def gpu():
    assert api.device_count() > 0, "no CUDA device: the PTP path has no fallback"
    return 0


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_toplesets_bit_exact(case, oracle, gpu):
    name, mesh, src = case
    with api.DeviceMesh(mesh, gpu) as dm:
        top, srt, lim = dm.compute_toplesets(src)
    t0, s0, l0 = oracle.compute_toplesets(mesh, src)
    assert np.array_equal(lim, l0)
    assert np.array_equal(srt, s0[:l0[-1]])
    assert np.array_equal(top, t0)


@pytest.mark.parametrize("k", [0, 1, 3, 7])
def test_toplesets_level_cap(k, oracle, gpu):
    mesh = mg.grid(25)
    src = [12 * 25 + 12, 3]
    with api.DeviceMesh(mesh, gpu) as dm:
        top, srt, lim = dm.compute_toplesets(src, k=k)
    t0, s0, l0 = oracle.compute_toplesets(mesh, src, k=k)
    assert np.array_equal(lim, l0)
    assert np.array_equal(srt, s0[:l0[-1]])
    assert np.array_equal(top, t0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_geodesics_matches_cpu_ptp(case, dtype, oracle, gpu):
    name, mesh, src = case
    m = mesh.astype(dtype)
    t0, s0, l0 = oracle.compute_toplesets(m, src)
    want, want_cl, st = oracle.ptp_cpu(m, src, l0, s0, clusters=True)
    with api.DeviceMesh(m, gpu) as dm:
        got, _, srt = dm.geodesics(src, want_sorted=True)
        stats = dict(dm.last_stats)
        got_c, cl, _ = dm.geodesics(src, clusters=True)
    assert_dist_parity(got, want, dtype, name)
    assert_dist_parity(got_c, want, dtype, name + " (clusters variant)")
    assert np.array_equal(srt, s0[:l0[-1]])
    assert np.array_equal(cl, want_cl), f"{name}: clusters differ"
    assert stats["iterations"] == st["iterations"]
    assert stats["vertex_updates"] == st["vertex_updates"]
    assert stats["n_levels"] == len(l0) - 1


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("case", CASES[:9], ids=IDS[:9])
def test_solve_with_host_toplesets(case, dtype, oracle, gpu):
    """parallel_toplesets_propagation_gpu signature: toplesets come from the caller (here: the oracle's BFS)."""
    name, mesh, src = case
    m = mesh.astype(dtype)
    t0, s0, l0 = oracle.compute_toplesets(m, src)
    want, want_cl, _ = oracle.ptp_cpu(m, src, l0, s0, clusters=True)
    with api.DeviceMesh(m, gpu) as dm:
        out = api.ptp_out_t(np.empty(m.n_vertices, dtype=dtype), np.empty(m.n_vertices, dtype=np.uint32))
        secs = api.parallel_toplesets_propagation_gpu(out, dm, src, api.toplesets_t(l0, s0))
    assert secs > 0
    assert_dist_parity(out.dist, want, dtype, name)
    assert np.array_equal(out.clusters, want_cl)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_solve_with_capped_toplesets(dtype, oracle, gpu):
    """partial toplesets (level cap): neighbours outside `sorted` behave as INF, as in the reference's index-space kernel."""
    m = mg.icosphere(10, 5e-3, seed=9).astype(dtype)
    src = [7]
    t0, s0, l0 = oracle.compute_toplesets(m, src, k=6)
    want, _, _ = oracle.ptp_cpu(m, src, l0, s0)
    with api.DeviceMesh(m, gpu) as dm:
        got, _ = dm.solve(src, l0, s0)
    assert_dist_parity(got, want, dtype, "capped toplesets")


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_batched_rows(dtype, oracle, gpu):
    m = mg.icosphere(14, 2e-3, seed=5).astype(dtype)
    srcs = mg.random_sources(1024, 40, m.n_vertices, unique=True)
    with api.DeviceMesh(m, gpu) as dm:
        rows = dm.solve_batched(srcs)
        stats = dict(dm.last_stats)
    upd = 0
    for b, s in enumerate(srcs):
        t0, s0, l0 = oracle.compute_toplesets(m, [s])
        want, _, st = oracle.ptp_cpu(m, [s], l0, s0)
        upd += st["vertex_updates"]
        assert_dist_parity(rows[b], want, dtype, f"row {b}")
    assert stats["vertex_updates"] == upd


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_wide_windows_whole_gpu(dtype, oracle, gpu):
    """torus whose band blows up (max window > 100k): exercises the compacted (sparse) iterations of the whole-GPU
    sweep and the change-driven skipping; still bit-equal, and far fewer relaxations than vertex-updates."""
    m = mg.torus(600, 300).astype(dtype)
    src = [17]
    t0, s0, l0 = oracle.compute_toplesets(m, src)
    want, want_cl, st = oracle.ptp_cpu(m, src, l0, s0, clusters=True)
    assert st["max_window"] > 60000
    with api.DeviceMesh(m, gpu) as dm:
        got, _, _ = dm.geodesics(src)
        stats = dict(dm.last_stats)
        got_c, cl, _ = dm.geodesics(src, clusters=True)
    assert_dist_parity(got, want, dtype, "torus 600x300")
    assert_dist_parity(got_c, want, dtype, "torus 600x300 (clusters variant)")
    assert np.array_equal(cl, want_cl)
    assert stats["iterations"] == st["iterations"] and stats["vertex_updates"] == st["vertex_updates"]
    assert 0 < stats["relaxations"] < stats["vertex_updates"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_wide_windows_batched(dtype, oracle, gpu):
    """one CTA per solve with windows far wider than the CTA (compacted iterations of the batched kernel)"""
    m = mg.torus(240, 120).astype(dtype)
    srcs = np.array([17, 5000, 28799, 12345], dtype=np.uint32)
    with api.DeviceMesh(m, gpu) as dm:
        rows = dm.solve_batched(srcs)
        stats = dict(dm.last_stats)
    assert stats["max_window"] > 4096 and stats["relaxations"] < stats["vertex_updates"]
    for b, s in enumerate(srcs):
        t0, s0, l0 = oracle.compute_toplesets(m, [s])
        want, _, _ = oracle.ptp_cpu(m, [s], l0, s0)
        assert_dist_parity(rows[b], want, dtype, f"row {b}")


def test_batched_source_sets(oracle, gpu):
    m = mg.torus(48, 20).astype(np.float32)
    sets = [[1, 500], [77], [3, 3, 900, 20], [959]]
    flat = np.concatenate(sets).astype(np.uint32)
    off = np.cumsum([0] + [len(s) for s in sets]).astype(np.uint64)
    with api.DeviceMesh(m, gpu) as dm:
        rows = dm.solve_batched(flat, off)
    for b, s in enumerate(sets):
        t0, s0, l0 = oracle.compute_toplesets(m, s)
        want, _, _ = oracle.ptp_cpu(m, s, l0, s0)
        assert_dist_parity(rows[b], want, np.float32, f"set {b}")


def test_geodesics_class_and_normalize(oracle, gpu):
    m = mg.grid(33)
    src = [16 * 33 + 16]
    t0, s0, l0 = oracle.compute_toplesets(m, src)
    want, _, _ = oracle.ptp_cpu(m, src, l0, s0)
    with api.DeviceMesh(m, gpu) as dm:
        g = api.geodesics(dm, src, api.geodesics.PTP_GPU, cluster=True)
    assert_dist_parity(g.dist, want, np.float64, "geodesics class")
    assert g(0) == src[0] and g.n_sorted_index() == 0
    assert np.array_equal(g.sorted_index, s0[:m.n_vertices])
    assert set(np.unique(g.clusters)) == {1}
    norm = want.copy()
    oracle.normalize_ptp(norm)
    g.normalize()
    assert np.array_equal(g.dist, norm)


def test_farthest_point_sampling(oracle, gpu):
    """FPS = repeated multi-source solves + first arg-max (cublasI?amax order)."""
    m = mg.icosphere(10, 4e-3, seed=2).astype(np.float64)
    samples = [0]
    with api.DeviceMesh(m, gpu) as dm:
        md, secs = api.farthest_point_sampling_ptp_gpu(dm, samples, 8)
    want = [0]
    for _ in range(7):
        t0, s0, l0 = oracle.compute_toplesets(m, want)
        d, _, _ = oracle.ptp_cpu(m, want, l0, s0)
        want.append(int(np.argmax(np.abs(d))))
    assert samples == want
    assert md == d[want[-1]]


def test_errors_are_reported(gpu):
    m = mg.grid(8)
    with api.DeviceMesh(m, gpu) as dm:
        with pytest.raises(api.PtpError):
            dm.geodesics([10 ** 6])
        with pytest.raises(api.PtpError):
            dm.geodesics([])
    bad = mg.grid(8)
    bad.OT = bad.OT.copy()
    bad.OT[:] = 5  # every walk cycles without returning to its start
    with pytest.raises(api.PtpError):
        api.DeviceMesh(bad, gpu)
