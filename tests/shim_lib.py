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
        L.shim_geodesics_class.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_int, C.c_int, C.c_int, rp, u32p, u32p, rp]
        L.shim_geodesics_class.restype = C.c_uint32
        L.shim_che_set_vertices.argtypes = [C.c_void_p, rp]
        L.shim_ptp_gpu_prefilled.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_int, rp, u32p]
        L.shim_ptp_gpu_prefilled.restype = C.c_double
        L.shim_normalize_ptp.argtypes = [rp, C.c_uint32]
        L.shim_distance_rows.argtypes = [C.c_void_p, u32p, C.c_uint32, rp, C.c_int]
        L.shim_distance_rows.restype = C.c_double
        L.shim_sampling_shape.argtypes = [C.c_void_p, u32p, C.c_uint32, ct, C.POINTER(C.c_ulong), u32p, C.c_ulong]
        L.shim_sampling_shape.restype = C.c_ulong
        L.shim_key_components.argtypes = [C.c_void_p, u32p, C.c_uint32, ct, u32p]
        L.shim_key_components.restype = C.c_ulong

    def che(self, xyz, faces):
        xyz = np.ascontiguousarray(xyz, dtype=self.dt)
        VT = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1)
        h = self.L.shim_che_create(xyz.ctypes.data_as(C.POINTER(self.ct)), xyz.shape[0], VT.ctypes.data_as(u32p), VT.size // 3)
        return h, xyz.shape[0]

    def destroy(self, h):
        self.L.shim_che_destroy(h)

    def gpu_vs_cpu(self, h, n_v, sources, coalescence=False, clusters=False):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        dg = np.full(n_v, np.inf, dtype=self.dt)   # like geodesics::geodesics (src/geodesics.cpp:27-28): the coalescence arm
        dc = np.full(n_v, np.nan, dtype=self.dt)   # only writes the vertices it reached
        cl = np.zeros(n_v, dtype=np.uint32) if clusters else None
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

    def geodesics_class(self, h, n_v, sources, opt=None, cluster=False, external_dist=False):
        """gproshan::geodesics(mesh, sources, opt, e_dist, cluster) -> (dist, sorted_index, clusters, n_sorted, normalized)"""
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        rp = C.POINTER(self.ct)
        d = np.empty(n_v, dtype=self.dt)
        dn = np.empty(n_v, dtype=self.dt)
        srt = np.empty(n_v, dtype=np.uint32)
        cl = np.zeros(n_v, dtype=np.uint32)
        opt = self.L.shim_option_ptp_gpu() if opt is None else opt
        ns = self.L.shim_geodesics_class(h, src.ctypes.data_as(u32p), src.size, opt, int(cluster), int(external_dist),
                                         d.ctypes.data_as(rp), srt.ctypes.data_as(u32p), cl.ctypes.data_as(u32p), dn.ctypes.data_as(rp))
        return d, srt, (cl if cluster else None), ns, dn

    def set_vertices(self, h, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=self.dt)
        self.L.shim_che_set_vertices(h, xyz.ctypes.data_as(C.POINTER(self.ct)))

    def ptp_gpu_prefilled(self, h, sources, dist_io, clusters_io, coalescence):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        return self.L.shim_ptp_gpu_prefilled(h, src.ctypes.data_as(u32p), src.size, int(coalescence), dist_io.ctypes.data_as(C.POINTER(self.ct)),
                                             None if clusters_io is None else clusters_io.ctypes.data_as(u32p))

    def fps(self, h, samples, n, radio=0.0):
        init = np.ascontiguousarray(samples, dtype=np.uint32)
        buf = np.zeros(max(n, init.size) + 8, dtype=np.uint32)
        buf[:init.size] = init
        md, secs = self.ct(0), C.c_double(0)
        cnt = self.L.shim_fps(h, buf.ctypes.data_as(u32p), init.size, n, self.ct(radio), C.byref(md), C.byref(secs))
        return buf[:cnt].copy(), md.value, secs.value

    def normalize_ptp(self, dist):
        d = np.ascontiguousarray(dist, dtype=self.dt).copy()
        self.L.shim_normalize_ptp(d.ctypes.data_as(C.POINTER(self.ct)), d.size)
        return d

    def distance_rows(self, h, n_v, points, n_devices=0):
        p = np.ascontiguousarray(points, dtype=np.uint32)
        rows = np.empty((p.size, n_v), dtype=self.dt)
        secs = self.L.shim_distance_rows(h, p.ctypes.data_as(u32p), p.size, rows.ctypes.data_as(C.POINTER(self.ct)), n_devices)
        return secs, rows

    def sampling_shape(self, h, n_v, points, radio):
        p = np.ascontiguousarray(points, dtype=np.uint32)
        sizes = np.zeros(p.size, dtype=np.uint64)
        flat = np.empty(p.size * n_v, dtype=np.uint32)
        tot = self.L.shim_sampling_shape(h, p.ctypes.data_as(u32p), p.size, self.ct(radio), sizes.ctypes.data_as(C.POINTER(C.c_ulong)),
                                         flat.ctypes.data_as(u32p), flat.size)
        out, at = [], 0
        for s in sizes:
            out.append(flat[at:at + int(s)].copy())
            at += int(s)
        assert at == tot
        return out

    def key_components(self, h, n_v, key_points, radio_fraction):
        k = np.ascontiguousarray(key_points, dtype=np.uint32)
        comp = np.empty(n_v, dtype=np.uint32)
        n = self.L.shim_key_components(h, k.ctypes.data_as(u32p), k.size, self.ct(radio_fraction), comp.ctypes.data_as(u32p))
        return int(n), comp
