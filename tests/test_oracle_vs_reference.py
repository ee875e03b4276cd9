"""CPU: the plain-C oracle against the reference's own CPU code compiled unmodified into oracle/_ref
(present in the build container and shipped to the GPU box as a prebuilt .so; skipped when absent)."""
import numpy as np
import pytest

from cases import small_cases
from oracle_lib import NIL, Reference, ref_available

pytestmark = pytest.mark.skipif(not (ref_available(np.float64) and ref_available(np.float32)),
                                reason="oracle/_ref not built (needs /root/reference at build time)")

CASES = small_cases()
IDS = [c[0] for c in CASES]


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_equals_reference(case, dtype, oracle):
    name, mesh, src = case
    m = mesh.astype(dtype)
    ref = Reference(dtype)
    rc = ref.che(m.GT, m.VT)
    _, VT, OT, EVT = rc.tables()
    OTo, EVTo, _ = oracle.che_build(m.n_vertices, m.VT)
    assert np.array_equal(OT, OTo) and np.array_equal(EVT, EVTo)
    # the product-side host builder (gproshan_b200/csrc/meshgen.c) must agree on manifold input
    assert np.array_equal(OT, m.OT) and np.array_equal(EVT, m.EVT)

    t0, s0, l0 = rc.compute_toplesets(src)
    t1, s1, l1 = oracle.compute_toplesets(m, src)
    assert np.array_equal(l0, l1) and np.array_equal(t0, t1) and np.array_equal(s0[:l0[-1]], s1[:l1[-1]])
    if len(l0) < 3:
        return  # the reference reads limits[2] unconditionally; nothing to compare
    d0 = rc.ptp_cpu(src, l0, s0)
    d1, _, _ = oracle.ptp_cpu(m, src, l1, s1)
    bits = np.uint64 if dtype == np.float64 else np.uint32
    assert np.array_equal(d0.view(bits), d1.view(bits))
    if l0[-1] <= m.n_vertices:  # the coalescence variant overruns its buffers with duplicate sources
        d0c = rc.ptp_cpu(src, l0, s0, coalescence=True)
        assert np.array_equal(d0.view(bits), d0c.view(bits))


@pytest.mark.parametrize("k", [0, 2, 5])
def test_level_cap_equals_reference(k, oracle):
    name, mesh, src = CASES[4]
    rc = Reference(np.float64).che(mesh.GT, mesh.VT)
    t0, s0, l0 = rc.compute_toplesets(src, k)
    t1, s1, l1 = oracle.compute_toplesets(mesh, src, k)
    assert np.array_equal(l0, l1) and np.array_equal(t0, t1) and np.array_equal(s0[:l0[-1]], s1[:l1[-1]])


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_update_step_equals_reference(dtype, oracle):
    name, mesh, src = CASES[4]
    m = mesh.astype(dtype)
    rc = Reference(dtype).che(m.GT, m.VT)
    rng = np.random.default_rng(0)
    dist = rng.uniform(0, 2, m.n_vertices).astype(dtype)
    dist[rng.integers(0, m.n_vertices, 200)] = np.inf
    for he in rng.integers(0, m.n_half_edges, 400):
        a, b = rc.update_step(dist, int(he)), oracle.update_step(m, dist, int(he))
        assert (np.isnan(a) and np.isnan(b)) or a == b
