"""ctypes access to the CPU oracle (oracle/libptp_oracle.so) and, when built, the reference's own
CPU code (oracle/_ref/libgproshan_ref_{f32,f64}.so). TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NIL = 0xFFFFFFFF
u32p = C.POINTER(C.c_uint32)
_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
_CT = {np.dtype(np.float32): C.c_float, np.dtype(np.float64): C.c_double}


def _p(a, t=C.c_uint32):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Oracle:
    """Plain-C restatement (oracle/ptp_oracle.c)."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "libptp_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle oracle`")
        self.L = L = C.CDLL(path)
        L.orc_che_build.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.orc_che_build.restype = C.c_int
        L.orc_compute_toplesets.argtypes = [C.c_uint32, u32p, u32p, u32p, u32p, C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.orc_compute_toplesets.restype = C.c_uint32
        for dt, suf in _SUF.items():
            rp = C.POINTER(_CT[dt])
            f = getattr(L, f"orc_ptp_cpu_{suf}")
            f.argtypes = [C.c_uint32, rp, u32p, u32p, u32p, u32p, C.c_uint32, u32p, C.c_uint32, u32p, rp, u32p,
                          C.c_uint32, C.POINTER(C.c_uint64)]
            f.restype = None
            g = getattr(L, f"orc_update_step_{suf}")
            g.argtypes = [rp, u32p, rp, C.c_uint32]
            g.restype = _CT[dt]
            h = getattr(L, f"orc_normalize_ptp_{suf}")
            h.argtypes = [rp, C.c_size_t]
            h.restype = None

    def che_build(self, n_v, faces):
        VT = _u32(faces).reshape(-1)
        OT = np.empty_like(VT)
        EVT = np.empty(n_v, dtype=np.uint32)
        manifold = self.L.orc_che_build(n_v, VT.size // 3, _p(VT), _p(OT), _p(EVT))
        return OT, EVT, bool(manifold)

    def compute_toplesets(self, mesh, sources, k=NIL):
        src = _u32(sources)
        n_v = mesh.n_vertices
        top = np.empty(n_v, dtype=np.uint32)
        srt = np.full(n_v + src.size, NIL, dtype=np.uint32)
        lim = np.empty(n_v + 2, dtype=np.uint32)
        nl = self.L.orc_compute_toplesets(n_v, _p(mesh.VT), _p(mesh.OT), _p(mesh.EVT), _p(src), src.size, k,
                                          _p(top), _p(srt), _p(lim))
        return top, srt, lim[:nl].copy()

    def ptp_cpu(self, mesh, sources, limits, sorted_, clusters=False, cluster_fill=NIL):
        dt = mesh.GT.dtype
        suf, ct = _SUF[dt], _CT[dt]
        src, lim, srt = _u32(sources), _u32(limits), _u32(sorted_)
        dist = np.empty(mesh.n_vertices, dtype=dt)
        cl = np.empty(mesh.n_vertices, dtype=np.uint32) if clusters else None
        stats = np.zeros(4, dtype=np.uint64)
        getattr(self.L, f"orc_ptp_cpu_{suf}")(mesh.n_vertices, _p(mesh.GT, ct), _p(mesh.VT), _p(mesh.OT), _p(mesh.EVT),
                                              _p(src), src.size, _p(lim), lim.size, _p(srt), _p(dist, ct), _p(cl),
                                              cluster_fill, _p(stats, C.c_uint64))
        st = dict(iterations=int(stats[0]), vertex_updates=int(stats[1]), max_window=int(stats[2]), d=int(stats[3]))
        return dist, cl, st

    def update_step(self, mesh, dist, he):
        dt = mesh.GT.dtype
        return getattr(self.L, f"orc_update_step_{_SUF[dt]}")(_p(mesh.GT, _CT[dt]), _p(mesh.VT), _p(dist, _CT[dt]), he)

    def normalize_ptp(self, dist):
        dt = dist.dtype
        getattr(self.L, f"orc_normalize_ptp_{_SUF[dt]}")(_p(dist, _CT[dt]), dist.size)


def ref_path(dtype):
    return os.path.join(ROOT, "oracle", "_ref", f"libgproshan_ref_{_SUF[np.dtype(dtype)]}.so")


def ref_available(dtype=np.float64):
    return os.path.exists(ref_path(dtype))


class Reference:
    """The reference's own CPU code, compiled unmodified (oracle/Makefile `ref`)."""

    def __init__(self, dtype):
        self.dt = np.dtype(dtype)
        self.ct = ct = _CT[self.dt]
        rp = C.POINTER(ct)
        self.L = L = C.CDLL(ref_path(dtype))
        assert L.ref_sizeof_real() == self.dt.itemsize
        L.ref_che_create.argtypes = [rp, C.c_uint32, u32p, C.c_uint32]
        L.ref_che_create.restype = C.c_void_p
        L.ref_che_create_raw.argtypes = [rp, C.c_uint32, u32p, u32p, u32p, C.c_uint32]
        L.ref_che_create_raw.restype = C.c_void_p
        L.ref_che_read_off.argtypes = [C.c_char_p]
        L.ref_che_read_off.restype = C.c_void_p
        L.ref_che_write_off.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_che_n_vertices.argtypes = [C.c_void_p]
        L.ref_che_n_half_edges.argtypes = [C.c_void_p]
        L.ref_che_destroy.argtypes = [C.c_void_p]
        L.ref_che_tables.argtypes = [C.c_void_p, rp, u32p, u32p, u32p]
        L.ref_compute_toplesets.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.ref_compute_toplesets.restype = C.c_uint32
        for n in ("ref_ptp_cpu", "ref_ptp_coalescence_cpu"):
            getattr(L, n).argtypes = [C.c_void_p, u32p, C.c_uint32, u32p, C.c_uint32, u32p, rp, u32p]
            getattr(L, n).restype = None
        L.ref_update_step.argtypes = [C.c_void_p, rp, C.c_uint32]
        L.ref_update_step.restype = ct
        L.ref_normalize_ptp.argtypes = [rp, C.c_size_t]

    def che(self, xyz, faces):
        """reference constructor: builds OT/EVT with the reference's own update_evt_ot_et"""
        xyz = np.ascontiguousarray(xyz, dtype=self.dt)
        VT = _u32(faces).reshape(-1)
        return RefChe(self, self.L.ref_che_create(_p(xyz, self.ct), xyz.shape[0], _p(VT), VT.size // 3), xyz.shape[0], VT.size)

    def read_off(self, path):
        """the reference's own OFF reader (che_off)"""
        h = self.L.ref_che_read_off(str(path).encode())
        return RefChe(self, h, self.L.ref_che_n_vertices(h), self.L.ref_che_n_half_edges(h))

    def che_raw(self, mesh):
        """inject prebuilt tables (skips the serial construction; for large meshes)"""
        GT = np.ascontiguousarray(mesh.GT, dtype=self.dt)
        h = self.L.ref_che_create_raw(_p(GT, self.ct), mesh.n_vertices, _p(mesh.VT), _p(mesh.OT), _p(mesh.EVT), mesh.n_faces)
        return RefChe(self, h, mesh.n_vertices, mesh.n_half_edges)


class RefChe:
    def __init__(self, ref, handle, n_v, n_he):
        self.ref, self.h, self.n_v, self.n_he = ref, handle, n_v, n_he

    def __del__(self):
        if self.h:
            self.ref.L.ref_che_destroy(self.h)
            self.h = None

    def write_off(self, path_without_ext):
        self.ref.L.ref_che_write_off(self.h, str(path_without_ext).encode())

    def tables(self):
        GT = np.empty((self.n_v, 3), dtype=self.ref.dt)
        VT = np.empty(self.n_he, dtype=np.uint32)
        OT = np.empty(self.n_he, dtype=np.uint32)
        EVT = np.empty(self.n_v, dtype=np.uint32)
        self.ref.L.ref_che_tables(self.h, _p(GT, self.ref.ct), _p(VT), _p(OT), _p(EVT))
        return GT, VT, OT, EVT

    def compute_toplesets(self, sources, k=NIL):
        src = _u32(sources)
        top = np.empty(self.n_v, dtype=np.uint32)
        srt = np.full(self.n_v + src.size, NIL, dtype=np.uint32)  # the reference only needs n_v; slack for duplicates
        lim = np.empty(self.n_v + 2, dtype=np.uint32)
        nl = self.ref.L.ref_compute_toplesets(self.h, _p(src), src.size, k, _p(top), _p(srt), _p(lim))
        return top, srt, lim[:nl].copy()

    def ptp_cpu(self, sources, limits, sorted_, coalescence=False):
        src, lim, srt = _u32(sources), _u32(limits), _u32(sorted_)
        dist = np.empty(self.n_v, dtype=self.ref.dt)
        f = self.ref.L.ref_ptp_coalescence_cpu if coalescence else self.ref.L.ref_ptp_cpu
        f(self.h, _p(src), src.size, _p(lim), lim.size, _p(srt), _p(dist, self.ref.ct), None)
        return dist

    def update_step(self, dist, he):
        return self.ref.L.ref_update_step(self.h, _p(dist, self.ref.ct), he)
