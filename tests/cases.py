"""Seeded synthetic cases shared by the CPU (oracle vs reference) and GPU (kernel vs oracle) parity tests."""
from __future__ import annotations

import numpy as np

from gproshan_b200 import meshgen as mg


def fan_mesh(n_spokes=14, closed=True):
    """A hub vertex of degree n_spokes (> 8: exercises the one-ring overflow rows), surrounded by a second ring."""
    ang = 2 * np.pi * np.arange(n_spokes) / n_spokes
    r1 = np.stack([np.cos(ang), np.sin(ang), 0.05 * np.sin(3 * ang)], 1)
    r2 = 2.0 * np.stack([np.cos(ang + np.pi / n_spokes), np.sin(ang + np.pi / n_spokes), 0.1 * np.cos(2 * ang)], 1)
    xyz = np.concatenate([[[0, 0, 0.2]], r1, r2])
    f = []
    last = n_spokes if closed else n_spokes - 1
    for k in range(last):
        a, b = 1 + k, 1 + (k + 1) % n_spokes
        f.append((0, a, b))
        c = 1 + n_spokes + k
        f.append((a, c, b))
        if closed or k + 1 < last:
            c2 = 1 + n_spokes + (k + 1) % n_spokes
            f.append((b, c, c2))
    return mg.che_from_faces(xyz, np.array(f, dtype=np.uint32).reshape(-1))


def small_cases():
    """(name, mesh (float64), sources)"""
    g = mg.grid(41)
    cases = [
        ("grid41_center", g, [20 * 41 + 20]),
        ("grid41_corner", g, [0]),
        ("grid_rect_23x57", mg.grid(23, 57), [5]),
        ("ico12", mg.icosphere(12), [0]),
        ("ico20_noise", mg.icosphere(20, 3e-3, seed=12345), [0]),
        ("ico9_noise_multi", mg.icosphere(9, 8e-3, seed=3), [5, 100, 333, 5]),          # duplicate source
        ("torus_60x24_multi", mg.torus(60, 24), list(mg.random_sources(7, 6, 60 * 24))),
        ("torus_aniso_1src", mg.torus(90, 12), [17]),                                   # band blows up (iterations > levels)
        ("grid_hole", mg.punch_hole(mg.grid(37), 18 * 37 + 18, 2), [3, 700]),           # open one-rings, isolated vertices
        ("grid_hole_dup", mg.punch_hole(mg.grid(30), 15 * 30 + 15, 3), [3, 3, 850]),
        ("fan14_closed", fan_mesh(14, True), [0]),                                      # degree 14 hub (overflow ring)
        ("fan14_from_rim", fan_mesh(14, True), [20]),
        ("fan11_open", fan_mesh(11, False), [3]),                                       # open overflow ring
        ("two_triangles", mg.grid(2), [0]),                                             # limits.size() == 3
        ("single_triangle_all_sources", mg.che_from_faces(np.eye(3), np.array([0, 1, 2], dtype=np.uint32)), [0, 1, 2]),  # limits.size() == 2
    ]
    return cases


def dtype_tol(dtype):
    # north_star tolerances; the design target (and what the tests assert first) is bit equality
    return 1e-5 if np.dtype(dtype) == np.float32 else 1e-10


def assert_dist_parity(got, want, dtype, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.dtype == want.dtype == np.dtype(dtype)
    inf_g, inf_w = np.isinf(got), np.isinf(want)
    assert np.array_equal(inf_g, inf_w), f"{what}: INF pattern differs"
    assert not np.isnan(got).any(), f"{what}: NaN in result"
    fin = ~inf_w
    denom = np.maximum(np.abs(want[fin]), np.finfo(want.dtype).tiny)
    rel = np.abs(got[fin] - want[fin]) / denom
    tol = dtype_tol(dtype)
    assert rel.size == 0 or rel.max() <= tol, f"{what}: max rel err {rel.max():.3e} > {tol}"
    bits = np.uint32 if np.dtype(dtype) == np.float32 else np.uint64
    nbad = int((got.view(bits) != want.view(bits)).sum())
    assert nbad == 0, f"{what}: within tolerance but {nbad} entries not bit-equal (design target is bit equality)"
