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


GEO_SYMBOLS = [  # the `geodesics` class of include/geodesics.h:18-72 (f64 mangling; f32 swaps d for f in the distance_t slots)
    "_ZN8gproshan9geodesicsC1EPNS_3cheERKSt6vectorIjSaIjEERKNS0_8option_tERKPdRKbRKmRKd",
    "_ZN8gproshan9geodesicsD1Ev", "_ZNK8gproshan9geodesicsixERKj", "_ZNK8gproshan9geodesicsclERKj",
    "_ZNK8gproshan9geodesics14n_sorted_indexEv", "_ZN8gproshan9geodesics9normalizeEv",
    "_ZNK8gproshan9geodesics17copy_sorted_indexEPjRKm", "_ZNK8gproshan9geodesics5radioEv", "_ZNK8gproshan9geodesics8farthestEv",
]


@needs_shim
def test_shim_exports_geodesics_class():
    out = subprocess.run(["nm", "-D", "--defined-only", shim_path(np.float64)], capture_output=True, text=True, check=True).stdout
    names = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for e in GEO_SYMBOLS:
        assert e in names, e


@needs_shim
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", CASES[:11], ids=[c[0] for c in CASES[:11]])
def test_geodesics_class_ptp_gpu(case, dtype, oracle):
    """gproshan::geodesics(mesh, sources, PTP_GPU, e_dist, cluster) from the drop-in class (shim/geodesics_b200.cpp) against the
    reference's OWN PTP_CPU arm run through the same class (reference compute_toplesets + parallel_toplesets_propagation_cpu)."""
    name, mesh, src = case
    if len(set(src)) != len(src):
        pytest.skip("duplicate sources overflow the reference's V-entry sorted_index (src/che.cpp:591)")
    sh = Shim(dtype)
    h, n_v = sh.che(mesh.GT, mesh.VT)
    try:
        cpu_opt = sh.L.shim_option_ptp_cpu()
        assert sh.L.shim_option_ptp_gpu() == 1 and sh.L.shim_option_fm() == 0   # enum layout with GPROSHAN_CUDA (geodesics.h:21-28)
        d_cpu, s_cpu, _, ns_cpu, dn_cpu = sh.geodesics_class(h, n_v, src, opt=cpu_opt)
        for ext in (False, True):
            d, srt, cl, ns, dn = sh.geodesics_class(h, n_v, src, cluster=True, external_dist=ext)
            assert ns == ns_cpu == 0                      # n_sorted stays 0 for PTP (src/geodesics.cpp:25)
            assert_dist_parity(d, d_cpu, dtype, name + " class PTP_GPU vs PTP_CPU")
            reached = np.isfinite(d_cpu)
            n_reached = int(reached.sum())
            assert np.array_equal(srt[:n_reached], s_cpu[:n_reached])     # operator(): BFS order
            assert (srt[n_reached:] == 0xFFFFFFFF).all()
            assert (cl[reached] >= 1).all() and (cl[reached] <= len(src)).all()
            assert np.array_equal(dn, dn_cpu) and np.array_equal(dn, sh.normalize_ptp(d_cpu))   # normalize() -> normalize_ptp
        m = mesh.astype(dtype)
        t0, s0, l0 = oracle.compute_toplesets(m, src)
        assert np.array_equal(cl, oracle.ptp_cpu(m, src, l0, s0, clusters=True)[1])
    finally:
        sh.destroy(h)


@needs_shim
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_shim_notices_in_place_edits(dtype):
    """the resident-mesh cache must not serve stale geometry after che::set_vertices on the same object."""
    from gproshan_b200 import meshgen as mg
    mesh = mg.icosphere(16, 2e-3, seed=1).astype(dtype)
    sh = Shim(dtype)
    h, n_v = sh.che(mesh.GT, mesh.VT)
    try:
        _, d0, c0, _ = sh.gpu_vs_cpu(h, n_v, [5])
        assert np.array_equal(d0, c0)
        xyz = np.asarray(mesh.GT, dtype=dtype) * dtype(1.5)
        xyz[7] *= dtype(1.01)                     # a single-vertex edit on top of a global one
        sh.set_vertices(h, xyz)
        _, d1, c1, _ = sh.gpu_vs_cpu(h, n_v, [5])
        assert np.array_equal(d1, c1), "solve after an in-place edit must use the new positions"
        assert not np.array_equal(d1, d0)
    finally:
        sh.destroy(h)


@needs_shim
@pytest.mark.gpu
def test_shim_coalescence_leaves_unreached_entries_untouched():
    """src/cuda/geodesics_ptp_coalescence.cu:84-86 writes dist[sorted[i]] for i < limits.back() only; the plain entry
    (geodesics_ptp.cu:60-66) overwrites the whole array."""
    from gproshan_b200 import meshgen as mg
    mesh = mg.punch_hole(mg.grid(25), 12 * 25 + 12, 2)   # the punched vertices are isolated: never reached
    sh = Shim(np.float64)
    h, n_v = sh.che(mesh.GT, mesh.VT)
    try:
        for coal in (True, False):
            d = np.full(n_v, -7.0)
            cl = np.full(n_v, 12345, dtype=np.uint32)
            assert sh.ptp_gpu_prefilled(h, [3], d, cl, coal) > 0
            unreached = ~np.isfinite(sh.gpu_vs_cpu(h, n_v, [3])[1])
            assert unreached.any()
            if coal:
                assert (d[unreached] == -7.0).all() and (cl[unreached] == 12345).all()
            else:
                assert np.isinf(d[unreached]).all()
            assert (d[~unreached] >= 0).all() and (cl[~unreached] == 1).all()
    finally:
        sh.destroy(h)


@needs_shim
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_shim_fps_reference_signature(dtype, oracle):
    """farthest_point_sampling_ptp_gpu through its reference signature: sample sequence vs the oracle loop, the `radio`
    early stop and the n >= V -> V / 2 clamp (src/cuda/geodesics_ptp.cu:121-148)."""
    from gproshan_b200 import meshgen as mg
    mesh = mg.icosphere(8, 5e-3, seed=4).astype(dtype)
    sh = Shim(dtype)
    h, n_v = sh.che(mesh.GT, mesh.VT)

    def oracle_fps(n, radio=0.0):
        s, md = [3], np.inf
        while len(s) < n and md > radio:
            t0, s0, l0 = oracle.compute_toplesets(mesh, s)
            d = oracle.ptp_cpu(mesh, s, l0, s0)[0]
            f = int(np.argmax(np.abs(d)))
            md = d[f]
            s.append(f)
        return s, md
    try:
        got, md, secs = sh.fps(h, [3], 10)
        want, wmd = oracle_fps(10)
        assert list(got) == want and md == wmd and secs > 0
        # radio: the loop stops after the first sample whose distance is <= radio (that sample is still appended)
        radio = float(wmd) * 1.5
        got_r, md_r, _ = sh.fps(h, [3], 200, radio)
        want_r, wmd_r = oracle_fps(200, radio)
        assert list(got_r) == want_r and md_r == wmd_r and md_r <= radio and len(got_r) < 200
        # clamp: n >= n_vertices -> n_vertices / 2 samples
        got_c, _, _ = sh.fps(h, [3], n_v + 5)
        assert len(got_c) == n_v // 2 and len(set(got_c.tolist())) == len(got_c)
    finally:
        sh.destroy(h)


@needs_shim
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_batched_callers_distance_rows_and_sampling_shape(dtype, oracle):
    """The batched mode behind caller-shaped C++ entry points (shim): distance-matrix rows for a point list over all
    visible devices, and the sampling_shape interface (src/sampling.cpp:16-38: per point, the vertices within `radio`
    by increasing geodesic distance) — against rows from the reference's CPU PTP."""
    from gproshan_b200 import meshgen as mg
    mesh = mg.icosphere(12, 3e-3, seed=6).astype(dtype)
    pts = mg.random_sources(5, 9, mesh.n_vertices, unique=True)
    sh = Shim(dtype)
    h, n_v = sh.che(mesh.GT, mesh.VT)
    try:
        secs, rows = sh.distance_rows(h, n_v, pts)
        assert secs > 0
        want = np.empty_like(rows)
        for k, p in enumerate(pts):
            _, _, want[k], _ = sh.gpu_vs_cpu(h, n_v, [int(p)])      # [2] = the reference's parallel_toplesets_propagation_cpu
        assert np.array_equal(rows, want)
        radio = float(np.median(want))
        patches = sh.sampling_shape(h, n_v, pts, radio)
        for k in range(pts.size):
            inside = np.nonzero(want[k] <= dtype(radio))[0]
            order = inside[np.lexsort((inside, want[k][inside]))]
            assert np.array_equal(patches[k], order)
            assert patches[k][0] == pts[k]
    finally:
        sh.destroy(h)


@needs_shim
@pytest.mark.gpu
def test_key_components_on_ptp_distances(oracle):
    """key_components::compute_kcs (src/key_components.cpp:51-63) routed onto the GPU solve: against the same construction in
    numpy on the reference's CPU PTP distances (star order from the CHE tables)."""
    from gproshan_b200 import meshgen as mg
    dtype = np.float64
    mesh = mg.icosphere(14, 3e-3, seed=9).astype(dtype)
    kps = mg.random_sources(21, 6, mesh.n_vertices, unique=True)
    frac = 0.18
    sh = Shim(dtype)
    h, n_v = sh.che(mesh.GT, mesh.VT)
    try:
        n_comp, comp = sh.key_components(h, n_v, kps, frac)
        _, _, dist, _ = sh.gpu_vs_cpu(h, n_v, kps)          # [2]: parallel_toplesets_propagation_cpu of the reference
    finally:
        sh.destroy(h)
    VT, OT, EVT = mesh.VT, mesh.OT, mesh.EVT
    nxt = lambda he: 3 * (he // 3) + (he + 1) % 3
    prv = lambda he: 3 * (he // 3) + (he + 2) % 3
    parent = np.arange(n_v)
    size = np.ones(n_v, dtype=np.int64)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    radio = frac * dist[np.isfinite(dist)].max()
    for v in np.lexsort((np.arange(n_v), dist)):
        if not dist[v] <= radio:
            break
        he = EVT[v]
        while he != 0xFFFFFFFF:
            x, y = find(v), find(VT[nxt(he)])
            if x != y:
                size[x] += size[y]
                parent[y] = x
            he = OT[prv(he)]
            if he == EVT[v]:
                break
    want = np.full(n_v, 0xFFFFFFFF, dtype=np.uint32)
    number, k = {}, 0
    for i in range(n_v):
        if parent[i] == i and size[i] > 1:
            number[i] = k
            k += 1
    for i in range(n_v):
        r = find(i)
        if r in number:
            want[i] = number[r]
    assert n_comp == k and 1 <= k <= len(kps)
    assert np.array_equal(comp, want)
