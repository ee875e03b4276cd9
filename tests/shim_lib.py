"""ctypes access to the C++ drop-in shim built against the reference (gproshan_b200/shim/_build). TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
_CT = {np.dtype(np.float32): C.c_float, np.dtype(np.float64): C.c_double}
u32p = C.POINTER(C.c_uint32)


def shim_path(dtype):
    return os.path.join(ROOT, "gproshan_b200", "shim", "_build", f"libgproshan_shim_{_SUF[np.dtype(dtype)]}.so")


def shim_available():
    return all(os.path.exists(shim_path(d)) for d in (np.float32, np.float64))


class Shim:
    def __init__(self, dtype):
        self.dt = np.dtype(dtype)
        self.ct = ct = _CT[self.dt]
        rp = C.POINTER(ct)
        self.L = L = C.CDLL(shim_path(dtype))
        assert L.shim_sizeof_real() == self.dt.itemsize
        L.shim_che_create.argtypes = [rp, C.c_uint32, u32p, C.c_uint32]
        L.shim_che_create.restype = C.c_void_p
        L.shim_che_destroy.argtypes = [C.c_void_p]
        L.shim_ptp_gpu_vs_cpu.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_int, rp, rp, u32p]
        L.shim_ptp_gpu_vs_cpu.restype = C.c_double
        L.shim_geodesics.argtypes = [C.c_void_p, u32p, C.c_uint32, rp, u32p, u32p]
        L.shim_geodesics.restype = C.c_double
        L.shim_fps.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_uint32, ct, rp, C.POINTER(C.c_double)]
        L.shim_fps.restype = C.c_uint32

    def che(self, xyz, faces):
        xyz = np.ascontiguousarray(xyz, dtype=self.dt)
        VT = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1)
        h = self.L.shim_che_create(xyz.ctypes.data_as(C.POINTER(self.ct)), xyz.shape[0], VT.ctypes.data_as(u32p), VT.size // 3)
        return h, xyz.shape[0]

    def destroy(self, h):
        self.L.shim_che_destroy(h)

    def gpu_vs_cpu(self, h, n_v, sources, coalescence=False, clusters=False):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        dg = np.empty(n_v, dtype=self.dt)
        dc = np.full(n_v, np.nan, dtype=self.dt)
        cl = np.empty(n_v, dtype=np.uint32) if clusters else None
        rp = C.POINTER(self.ct)
        secs = self.L.shim_ptp_gpu_vs_cpu(h, src.ctypes.data_as(u32p), src.size, int(coalescence), dg.ctypes.data_as(rp),
                                          dc.ctypes.data_as(rp), None if cl is None else cl.ctypes.data_as(u32p))
        return secs, dg, dc, cl

    def geodesics(self, h, n_v, sources):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        d = np.empty(n_v, dtype=self.dt)
        srt = np.full(n_v, 0xFFFFFFFF, dtype=np.uint32)
        secs = self.L.shim_geodesics(h, src.ctypes.data_as(u32p), src.size, d.ctypes.data_as(C.POINTER(self.ct)), None,
                                     srt.ctypes.data_as(u32p))
        return secs, d, srt
