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


def test_batched_host_rows_in_several_launches(oracle, gpu):
    """host rows larger than the device staging buffer: several launches, each streaming its finished rows to the host
    while it runs (`stream_rows`), against one launch with the rows assembled on the device"""
    import torch
    m = mg.icosphere(16, 2e-3, seed=11).astype(np.float32)
    srcs = mg.random_sources(5, 150, m.n_vertices, unique=True)
    dev = torch.empty((srcs.size, m.n_vertices), dtype=torch.float32, device="cuda")
    try:
        api.set_option("rows_chunk_mb", 1)     # 1 MiB / (4 B x 2 562 vertices) = 102 rows per launch
        with api.DeviceMesh(m, gpu) as dm:
            dm.solve_batched(srcs, rows_device_ptr=dev.data_ptr())
            want = dev.cpu().numpy()
            got = dm.solve_batched(srcs)
            assert dm.last_stats["gpu_launches"] == 2
            api.set_option("stream_rows", 0)
            got_plain = dm.solve_batched(srcs)
    finally:
        api.set_option("rows_chunk_mb", 8192)
        api.set_option("stream_rows", 1)
    assert np.array_equal(got, want) and np.array_equal(got_plain, want)
    assert srcs.size > 102 + 32    # two launches, the second still large enough to stream
    t0, s0, l0 = oracle.compute_toplesets(m, [srcs[-1]])
    assert_dist_parity(got[-1], oracle.ptp_cpu(m, [srcs[-1]], l0, s0)[0], np.float32, "last row")


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


# ---------------------------------------------------------------- CHE construction on the device

@pytest.mark.parametrize("name", ["grid", "torus", "ico", "hole", "fan_open"])
def test_che_build_device_matches_reference_tables(name, oracle, gpu):
    from cases import fan_mesh
    m = {"grid": mg.grid(23, 31), "torus": mg.torus(30, 12), "ico": mg.icosphere(9),
         "hole": mg.punch_hole(mg.grid(21), 10 * 21 + 10, 2), "fan_open": fan_mesh(11, False)}[name]
    OT, EVT, manifold, ms = api.che_build(m.VT, m.n_vertices, gpu)
    OTo, EVTo, _ = oracle.che_build(m.n_vertices, m.VT)   # restatement of che::update_evt_ot_et, pinned to the reference
    assert manifold and ms > 0
    assert np.array_equal(OT, OTo) and np.array_equal(EVT, EVTo)


def test_che_build_device_flags_non_manifold(gpu):
    faces = np.array([0, 1, 2, 0, 1, 3, 0, 1, 4], dtype=np.uint32)  # directed edge 0->1 three times
    _, _, manifold, _ = api.che_build(faces, 5, gpu)
    assert not manifold
    with pytest.raises(api.PtpError):
        api.DeviceMesh(api.FaceMesh(np.zeros((5, 3)), faces), gpu)
    bowtie = np.array([0, 1, 2, 0, 3, 4], dtype=np.uint32)  # two border fans meeting at vertex 0
    _, _, manifold, _ = api.che_build(bowtie, 5, gpu)
    assert not manifold


def test_mesh_from_faces_equals_mesh_from_tables(oracle, gpu):
    m = mg.icosphere(11, 4e-3, seed=8).astype(np.float32)
    with api.DeviceMesh(m, gpu) as a, api.DeviceMesh(api.FaceMesh(m.GT, m.VT), gpu) as b:
        da, _, sa = a.geodesics([3, 99], want_sorted=True)
        db, _, sb = b.geodesics([3, 99], want_sorted=True)
    assert np.array_equal(da, db) and np.array_equal(sa, sb)


# ---------------------------------------------------------------- BASELINE configs at their stated sizes

@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_config_c1_grid_through_off(dtype, oracle, gpu, tmp_path):
    """configs[0]: 317x317 grid (100 489 vertices) written to / read from OFF, source = centre vertex."""
    from gproshan_b200.off_io import read_off, write_off
    g = mg.grid(317)
    write_off(tmp_path / "grid317.off", g.GT, g.VT)
    xyz, faces = read_off(tmp_path / "grid317.off", dtype=dtype)
    src = [158 * 317 + 158]
    with api.DeviceMesh(api.FaceMesh(xyz, faces), gpu) as dm:
        got, _, srt = dm.geodesics(src, want_sorted=True)
        stats = dict(dm.last_stats)
    m = g.astype(dtype)
    t0, s0, l0 = oracle.compute_toplesets(m, src)
    want, _, st = oracle.ptp_cpu(m, src, l0, s0)
    assert np.array_equal(srt, s0[:l0[-1]])
    assert_dist_parity(got, want, dtype, "C1")
    assert stats["vertex_updates"] == st["vertex_updates"]
    # sanity (not parity): on the flat grid PTP is within a few % of the Euclidean distance
    eu = np.linalg.norm(g.GT - g.GT[src[0]], axis=1)
    assert np.abs(got - eu).max() < 0.05


def test_config_c2_icosphere_1m_float(oracle, gpu):
    """configs[1]: 998 562-vertex icosphere, float, single source."""
    m = mg.icosphere(316, dtype=np.float32)
    assert m.n_vertices == 998562
    with api.DeviceMesh(m, gpu) as dm:
        got, _, srt = dm.geodesics([0], want_sorted=True)
    t0, s0, l0 = oracle.compute_toplesets(m, [0])
    want, _, _ = oracle.ptp_cpu(m, [0], l0, s0)
    assert np.array_equal(srt, s0[:l0[-1]])
    assert_dist_parity(got, want, np.float32, "C2")
    gc = np.arccos(np.clip(m.GT.astype(np.float64) @ m.GT[0].astype(np.float64), -1, 1))
    assert np.abs(got - gc).max() < 0.08  # sanity: great-circle distance on the unit sphere


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_three_device_paths_agree_at_scale(dtype, gpu):
    """Size-independent property, no oracle needed: the fused two-team kernel (ptp_geodesics), the stand-alone
    sweep fed with toplesets from ptp_toplesets (ptp_solve) and the one-CTA-per-solve kernel (ptp_solve_batched)
    are three different schedules of the same arithmetic and must agree bit for bit; so must repeated runs."""
    m = mg.icosphere(400, noise_sigma=0.2 * mg.mean_edge_icosphere(400), seed=12345, dtype=dtype)  # 1.6 M vertices
    src = [123456]
    with api.DeviceMesh(m, gpu) as dm:
        a, _, _ = dm.geodesics(src)
        a2, _, _ = dm.geodesics(src)
        top, srt, lim = dm.compute_toplesets(src)
        b, _ = dm.solve(src, lim, srt)
        c = dm.solve_batched(np.array(src, dtype=np.uint32))[0]
    assert np.array_equal(a, a2) and np.array_equal(a, b) and np.array_equal(a, c)
    assert np.isfinite(a).all() and a[src[0]] == 0 and (top != NIL).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_config_c4_shape_multi_source_voronoi(dtype, oracle, gpu):
    """configs[3] at 1/9 of its size (same aspect 3780:1323, R=1, r=0.35): 64 sources `mt19937(7)() % V`, clusters on.
    Distances bit-equal, Voronoi clusters equal, every reached vertex labelled 1..64."""
    m = mg.torus(1260, 441).astype(dtype)
    src = mg.random_sources(7, 64, m.n_vertices)
    t0, s0, l0 = oracle.compute_toplesets(m, src)
    want, want_cl, st = oracle.ptp_cpu(m, src, l0, s0, clusters=True)
    with api.DeviceMesh(m, gpu) as dm:
        got, cl, srt = dm.geodesics(src, clusters=True, want_sorted=True)
        stats = dict(dm.last_stats)
        g2 = api.geodesics(dm, src, api.geodesics.PTP_GPU, cluster=True)
    assert np.array_equal(srt, s0[:l0[-1]])
    assert_dist_parity(got, want, dtype, "C4-shape")
    assert np.array_equal(cl, want_cl) and cl.min() >= 1 and cl.max() <= 64
    assert stats["iterations"] == st["iterations"] and stats["vertex_updates"] == st["vertex_updates"]
    assert stats["relaxations"] < stats["vertex_updates"]
    assert np.array_equal(g2.dist, got) and np.array_equal(g2.clusters, cl)
    # each source is its own nearest source
    assert np.array_equal(cl[src], 1 + np.array([np.nonzero(src == s)[0].max() for s in src]))


@pytest.mark.gpu
@pytest.mark.parametrize("real_size", [4, 8])
def test_inv_gram_shared_reciprocal_is_ieee_division(real_size):
    """The inverse Gram matrix of update_step (src/geodesics_ptp.cpp:212-231: q11/det, -q01/det, q00/det) is computed with
    one reciprocal shared by the three divisions; every result must equal the plain IEEE division in every bit, over
    Gram matrices of random edge pairs at all scales, raw bit patterns and specials — and the shared path must be the one
    that normally runs."""
    import ctypes as C
    from gproshan_b200 import _lib
    L = _lib.lib()
    bad, fast = C.c_uint64(0), C.c_uint64(0)
    n = 400_000_000
    _lib.check(L.ptp_debug_inv_gram_check(n, 20261017 + real_size, real_size, C.byref(bad), C.byref(fast), None))
    assert bad.value == 0
    assert fast.value > n // 8  # the well-shaped geometric cases at moderate scales take the shared path


@pytest.mark.gpu
@pytest.mark.parametrize("real_size", [4, 8])
def test_short_sign_test_agrees_with_reference_chain(real_size):
    """update_step keeps the planar value iff c0 < 0 and c1 < 0 (src/geodesics_ptp.cpp:239-253). The batched sweep reads the
    signs off the two-term form e = Q (t - p) when |e| is above a proven bound; wherever it decides, the decision (and the
    returned value) must be the reference chain's — on random triangles up to the edge of the admitted shapes and on
    adversarial distances that make e nearly cancel."""
    import ctypes as C
    from gproshan_b200 import _lib
    L = _lib.lib()
    bad, dec, flg = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    n = 300_000_000
    _lib.check(L.ptp_debug_sign_short_check(n, 4242 + real_size, real_size, C.byref(bad), C.byref(dec), C.byref(flg)))
    assert bad.value == 0
    assert flg.value > n // 4 and dec.value > flg.value // 4  # the test must exercise the short form


@pytest.mark.gpu
def test_range_free_sqrt_is_ieee_sqrt():
    """every float in [2^-96, 2^96]: the square root without the range test (the PTP_FLAG_RANGE build variant uses it where a
    triangle's flag bounds the argument; in the default build sqrt_n IS __fsqrt_rn) equals __fsqrt_rn bit for bit"""
    import ctypes as C
    from gproshan_b200 import _lib
    bad, n = C.c_uint64(0), C.c_uint64(0)
    _lib.check(_lib.lib().ptp_debug_sqrt_check(C.byref(bad), C.byref(n)))
    assert bad.value == 0 and n.value == 192 * (1 << 23) + 1


@pytest.mark.gpu
@pytest.mark.parametrize("real_size", [4, 8])
def test_two_sided_skip_never_hides_an_improvement(real_size):
    """whenever the two-sided causal skip fires, update_step (reference chain) returns a value that is not below the
    vertex's — on random triangles and on inputs placed at the edges of the rule"""
    import ctypes as C
    from gproshan_b200 import _lib
    bad, fired, flg = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    n = 300_000_000
    _lib.check(_lib.lib().ptp_debug_two_sided_check(n, 777 + real_size, real_size, C.byref(bad), C.byref(fired), C.byref(flg)))
    assert bad.value == 0
    assert flg.value > n // 8 and fired.value > flg.value // 8  # the test must exercise the rule
