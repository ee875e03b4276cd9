"""Synthetic meshes for the PTP configs of BASELINE.json plus a host CHE-table builder.

Thin ctypes front-end over ``csrc/meshgen.c`` (built in-tree as ``libptp_meshgen.so``).
Host-side input preparation only; nothing here runs on the GPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

NIL = 0xFFFFFFFF
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libptp_meshgen.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -m gproshan_b200.build` (or __graft_entry__.build())")
        L = C.CDLL(path)
        u32p, f64p = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
        L.mg_mt19937.argtypes = [C.c_uint32, C.c_size_t, u32p]
        L.mg_radial_noise.argtypes = [f64p, C.c_size_t, C.c_double, C.c_uint64]
        L.mg_grid.argtypes = [C.c_uint32, C.c_uint32, f64p, u32p]
        L.mg_torus.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_double, f64p, u32p]
        L.mg_icosphere_counts.argtypes = [C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.mg_icosphere.argtypes = [C.c_uint32, f64p, u32p]
        L.mg_che_build.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.mg_che_build.restype = C.c_int
        for f in (L.mg_mt19937, L.mg_radial_noise, L.mg_grid, L.mg_torus, L.mg_icosphere_counts, L.mg_icosphere):
            f.restype = None
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@dataclass
class CheMesh:
    """Host mirror of the tables the reference's ``CHE`` POD exposes (include/che.h:134-146):
    GT (V x 3 reals), VT / OT (H = 3F half-edges), EVT (V)."""
    GT: np.ndarray
    VT: np.ndarray
    OT: np.ndarray
    EVT: np.ndarray

    @property
    def n_vertices(self) -> int:
        return self.GT.shape[0]

    @property
    def n_half_edges(self) -> int:
        return self.VT.shape[0]

    @property
    def n_faces(self) -> int:
        return self.VT.shape[0] // 3

    def astype(self, dtype) -> "CheMesh":
        return CheMesh(np.ascontiguousarray(self.GT, dtype=dtype), self.VT, self.OT, self.EVT)


def che_from_faces(xyz: np.ndarray, faces: np.ndarray) -> CheMesh:
    """Build OT/EVT for an oriented edge-manifold triangle list (host, OpenMP)."""
    xyz = np.ascontiguousarray(xyz)
    VT = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1)
    n_v, n_f = xyz.shape[0], VT.shape[0] // 3
    OT = np.empty(3 * n_f, dtype=np.uint32)
    EVT = np.empty(n_v, dtype=np.uint32)
    ok = _lib().mg_che_build(n_v, n_f, _p(VT, C.c_uint32), _p(OT, C.c_uint32), _p(EVT, C.c_uint32))
    if not ok:
        raise ValueError("face list is not an oriented edge-manifold triangle mesh")
    return CheMesh(xyz, VT, OT, EVT)


def mt19937(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    _lib().mg_mt19937(seed, n, _p(out, C.c_uint32))
    return out


def random_sources(seed: int, n: int, n_vertices: int, unique: bool = False) -> np.ndarray:
    """``mt19937(seed)() % V`` as SURVEY.md §8d specifies; ``unique`` de-duplicates keeping first occurrences."""
    s = (mt19937(seed, n).astype(np.uint64) % np.uint64(n_vertices)).astype(np.uint32)
    if unique:
        _, first = np.unique(s, return_index=True)
        s = s[np.sort(first)]
    return s


def grid(nx: int, ny: int | None = None, dtype=np.float64) -> CheMesh:
    ny = nx if ny is None else ny
    xyz = np.empty((nx * ny, 3), dtype=np.float64)
    faces = np.empty(2 * (nx - 1) * (ny - 1) * 3, dtype=np.uint32)
    _lib().mg_grid(nx, ny, _p(xyz, C.c_double), _p(faces, C.c_uint32))
    return che_from_faces(xyz.astype(dtype), faces)


def torus(nu: int, nv: int, R: float = 1.0, r: float = 0.35, dtype=np.float64) -> CheMesh:
    xyz = np.empty((nu * nv, 3), dtype=np.float64)
    faces = np.empty(2 * nu * nv * 3, dtype=np.uint32)
    _lib().mg_torus(nu, nv, R, r, _p(xyz, C.c_double), _p(faces, C.c_uint32))
    return che_from_faces(xyz.astype(dtype), faces)


def icosphere(f: int, noise_sigma: float = 0.0, seed: int = 12345, dtype=np.float64) -> CheMesh:
    """Class-I icosphere of frequency f (10f^2+2 vertices); optional radial noise 1+sigma*U(-1,1)."""
    nv, nf = C.c_uint64(), C.c_uint64()
    _lib().mg_icosphere_counts(f, C.byref(nv), C.byref(nf))
    xyz = np.empty((nv.value, 3), dtype=np.float64)
    faces = np.empty(nf.value * 3, dtype=np.uint32)
    _lib().mg_icosphere(f, _p(xyz, C.c_double), _p(faces, C.c_uint32))
    if noise_sigma:
        _lib().mg_radial_noise(_p(xyz, C.c_double), nv.value, noise_sigma, seed)
    return che_from_faces(xyz.astype(dtype), faces)


def mean_edge_icosphere(f: int) -> float:
    """Approximate mean edge length of the unit class-I icosphere (used to scale the C3 noise)."""
    return 1.1071487177940904 / f * 1.0  # arc of an icosahedron edge (atan 2) split f ways


def punch_hole(mesh: CheMesh, center: int, rings: int = 2) -> CheMesh:
    """Drop every face within `rings` edge-hops of `center` (gives border vertices / open one-rings)."""
    VT = mesh.VT.reshape(-1, 3)
    mark = np.zeros(mesh.n_vertices, dtype=bool)
    mark[center] = True
    for _ in range(rings):
        touched = mark[VT].any(axis=1)
        mark[VT[touched].reshape(-1)] = True
    keep = ~mark[VT].any(axis=1)
    return che_from_faces(mesh.GT, VT[keep].reshape(-1))
