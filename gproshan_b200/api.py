"""Host-side mirror of the reference's PTP interface over the C ABI (include/ptp_b200.h).

Names, argument meaning and results follow larc/gproshan (file:line relative to that repo):

  che::compute_toplesets                      src/che.cpp:546-593          -> DeviceMesh.compute_toplesets
  parallel_toplesets_propagation_gpu          include/geodesics_ptp.h:36   -> parallel_toplesets_propagation_gpu
  parallel_toplesets_propagation_coalescence_gpu  include/geodesics_ptp.h:34 -> same entry (single layout path here)
  farthest_point_sampling_ptp_gpu             include/geodesics_ptp.h:42   -> farthest_point_sampling_ptp_gpu
  class geodesics (option PTP_GPU)            include/geodesics.h:18-72    -> class geodesics
  normalize_ptp                               src/geodesics_ptp.cpp:264-276 -> geodesics.normalize

This is glue for tests, bench.py and Python callers; the compiled drop-in for gproshan itself is the C++
shim under gproshan_b200/shim. Everything computes on the GPU through libptp_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import PTP_NIL, PtpError, Stats, check

NIL = PTP_NIL
_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
_CT = {np.dtype(np.float32): C.c_float, np.dtype(np.float64): C.c_double}


def _p(a, t=C.c_uint32):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def device_count() -> int:
    return _lib.lib().ptp_device_count()


def set_option(name: str, value: int) -> None:
    """ptp_set_option: kernel-variant switches of the library (see `options()`); same bits whatever the choice."""
    check(_lib.lib().ptp_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    return _lib.lib().ptp_get_option(name.encode())


def options() -> dict:
    """{name: (value, doc)} of every switch the library has."""
    L, out, i = _lib.lib(), {}, 0
    while (n := L.ptp_option_name(i)) is not None:
        out[n.decode()] = (L.ptp_get_option(n), L.ptp_option_doc(i).decode())
        i += 1
    return out


def che_build(faces, n_vertices: int, device: int = 0):
    """OT / EVT from a face list on the GPU (che::update_evt_ot_et, src/che.cpp:1295-1362).
    -> (OT, EVT, manifold, device_ms)"""
    VT = _u32(faces).reshape(-1)
    OT = np.empty_like(VT)
    EVT = np.empty(n_vertices, dtype=np.uint32)
    mf, ms = C.c_int(), C.c_double()
    check(_lib.lib().ptp_che_build(_p(VT), n_vertices, VT.size, _p(OT), _p(EVT), device, C.byref(mf), C.byref(ms)))
    return OT, EVT, bool(mf.value), ms.value


class FaceMesh:
    """A bare face list + positions: DeviceMesh(FaceMesh(xyz, faces)) builds OT / EVT on the device."""

    def __init__(self, xyz, faces):
        self.GT = np.asarray(xyz)
        self.VT = _u32(faces).reshape(-1)
        self.OT = self.EVT = None


class DeviceMesh:
    """A CHE mesh resident on one GPU (replaces the per-call CHE(mesh) + cuda_create_CHE upload,
    src/che.cpp:36-46, src/cuda/che.cu:29-48). Accepts anything with GT / VT / OT / EVT arrays."""

    def __init__(self, mesh, device: int = 0, dtype=None):
        L = _lib.lib()
        GT = np.asarray(mesh.GT)
        self.dtype = np.dtype(dtype if dtype is not None else GT.dtype)
        if self.dtype not in _SUF:
            raise TypeError("real_t must be float32 or float64")
        GT = np.ascontiguousarray(GT, dtype=self.dtype)
        VT = _u32(mesh.VT)
        OT = None if mesh.OT is None else _u32(mesh.OT)
        EVT = None if mesh.EVT is None else _u32(mesh.EVT)
        if GT.ndim != 2 or GT.shape[1] != 3 or (OT is not None and (OT.shape != VT.shape or EVT.shape[0] != GT.shape[0])):
            raise ValueError("inconsistent CHE table shapes")
        self.n_vertices, self.n_half_edges = GT.shape[0], VT.shape[0]
        self.suf, self.ct = _SUF[self.dtype], _CT[self.dtype]
        self._h = C.c_void_p()
        check(getattr(L, f"ptp_mesh_create_{self.suf}")(_p(GT, self.ct), _p(VT), _p(OT), _p(EVT), self.n_vertices,
                                                        self.n_half_edges, device, C.byref(self._h)))
        self.device = device
        self.last_stats: dict = {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().ptp_mesh_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def update_positions(self, xyz):
        """new vertex positions, same connectivity (ptp_mesh_update_positions_*)"""
        GT = np.ascontiguousarray(xyz, dtype=self.dtype)
        assert GT.shape == (self.n_vertices, 3)
        check(getattr(_lib.lib(), f"ptp_mesh_update_positions_{self.suf}")(self._h, _p(GT, self.ct)))

    @property
    def last_kernel(self) -> str:
        """dominant kernel of the last call on this mesh (which single-solve variant ran)"""
        return _lib.lib().ptp_mesh_last_kernel(self._h).decode()

    @property
    def device_bytes(self) -> int:
        return _lib.lib().ptp_mesh_device_bytes(self._h)

    # ---- che::compute_toplesets
    def compute_toplesets(self, sources, k: int = NIL, want_toplesets: bool = True):
        """-> (toplesets[V] (NIL = unreached), sorted[limits[-1]], limits)"""
        src = _u32(sources)
        V = self.n_vertices
        top = np.empty(V, dtype=np.uint32) if want_toplesets else None
        srt = np.empty(V + src.size, dtype=np.uint32)
        lim = np.empty(V + 2, dtype=np.uint32)
        nl = C.c_uint32()
        st = Stats()
        check(_lib.lib().ptp_toplesets(self._h, _p(src), src.size, k, _p(top), _p(srt), srt.size, _p(lim), lim.size,
                                       C.byref(nl), C.byref(st)))
        self.last_stats = st.as_dict()
        lim = lim[:nl.value].copy()
        return top, srt[:lim[-1]].copy(), lim

    # ---- parallel_toplesets_propagation_gpu with caller-provided toplesets
    def solve(self, sources, limits, sorted_, clusters: bool = False, cluster_fill: int = NIL, out=None):
        src, lim, srt = _u32(sources), _u32(limits), _u32(sorted_)
        dist = out if out is not None else np.empty(self.n_vertices, dtype=self.dtype)
        cl = np.empty(self.n_vertices, dtype=np.uint32) if clusters else None
        st = Stats()
        check(getattr(_lib.lib(), f"ptp_solve_{self.suf}")(self._h, _p(src), src.size, _p(lim), lim.size, _p(srt),
                                                           _p(dist, self.ct), _p(cl), cluster_fill, C.byref(st)))
        self.last_stats = st.as_dict()
        return dist, cl

    # ---- geodesics::run_parallel_toplesets_propagation_gpu (toplesets on device + solve)
    def geodesics(self, sources, clusters: bool = False, cluster_fill: int = NIL, want_sorted: bool = False, out=None):
        src = _u32(sources)
        dist = out if out is not None else np.empty(self.n_vertices, dtype=self.dtype)
        cl = np.empty(self.n_vertices, dtype=np.uint32) if clusters else None
        srt = np.empty(self.n_vertices + src.size, dtype=np.uint32) if want_sorted else None
        st = Stats()
        check(getattr(_lib.lib(), f"ptp_geodesics_{self.suf}")(self._h, _p(src), src.size, _p(dist, self.ct), _p(cl),
                                                               cluster_fill, _p(srt), 0 if srt is None else srt.size,
                                                               C.byref(st)))
        self.last_stats = st.as_dict()
        if srt is not None:
            srt = srt[:st.n_reached]
        return dist, cl, srt

    # ---- per-iteration error (iter_error_run_ptp_gpu, src/cuda/test_geodesics_ptp.cu:164-211)
    def error_per_iteration(self, sources, exact, capacity: int = 4096):
        """-> (iterations, errors in %, final distances): one record per iteration whose window ends at the last topleset"""
        src = _u32(sources)
        ex = np.ascontiguousarray(exact, dtype=self.dtype)
        assert ex.size == self.n_vertices
        dist = np.empty(self.n_vertices, dtype=self.dtype)
        it = np.empty(capacity, dtype=np.uint32)
        er = np.empty(capacity, dtype=self.dtype)
        n = C.c_uint32()
        st = Stats()
        check(getattr(_lib.lib(), f"ptp_geodesics_error_iter_{self.suf}")(self._h, _p(src), src.size, _p(ex, self.ct), _p(dist, self.ct),
                                                                          _p(it), _p(er, self.ct), capacity, C.byref(n), C.byref(st)))
        self.last_stats = st.as_dict()
        return it[:n.value].copy(), er[:n.value].copy(), dist

    # ---- batched independent solves (distance-matrix rows)
    def solve_batched(self, sources, offsets=None, rows=None, rows_device_ptr: int | None = None, stream: int | None = None):
        """One solve per source (offsets=None) or per source set sources[offsets[b]:offsets[b+1]].
        rows: host array (B, V) to fill (allocated if None) — or rows_device_ptr: device pointer to B*V reals."""
        src = _u32(sources)
        off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.uint64)
        B = src.size if off is None else off.size - 1
        st = Stats()
        fn = getattr(_lib.lib(), f"ptp_solve_batched_{self.suf}")
        if rows_device_ptr is not None:
            check(fn(self._h, _p(src), _p(off, C.c_uint64), B, src.size, C.c_void_p(rows_device_ptr), 1,
                     C.c_void_p(stream or 0), C.byref(st)))
            self.last_stats = st.as_dict()
            return None
        if rows is None:
            rows = np.empty((B, self.n_vertices), dtype=self.dtype)
        assert rows.dtype == self.dtype and rows.flags.c_contiguous and rows.size == B * self.n_vertices
        check(fn(self._h, _p(src), _p(off, C.c_uint64), B, src.size, C.c_void_p(rows.ctypes.data), 0,
                 C.c_void_p(stream or 0), C.byref(st)))
        self.last_stats = st.as_dict()
        return rows

    # ---- farthest_point_sampling_ptp_gpu
    def farthest_point_sampling(self, samples, n: int, radio: float = 0.0):
        """-> (samples (grown to n), max_dist). Mirrors farthest_point_sampling_ptp_gpu (src/cuda/geodesics_ptp.cu:87-172)."""
        init = _u32(samples)
        buf = np.zeros(max(n, init.size), dtype=np.uint32)
        buf[:init.size] = init
        n_out = C.c_uint32()
        md = self.ct()
        st = Stats()
        check(getattr(_lib.lib(), f"ptp_farthest_point_sampling_{self.suf}")(self._h, _p(buf), init.size, n, self.ct(radio),
                                                                            C.byref(n_out), C.byref(md), C.byref(st)))
        self.last_stats = st.as_dict()
        return buf[:n_out.value].copy(), md.value


def solve_batched_multi(meshes, sources, offsets=None, rows=None, rows_device_ptr: int | None = None):
    """ptp_solve_batched_multi_*: one process, several devices. `meshes` = DeviceMesh objects of the same mesh on different
    devices; rows land in a host array (allocated if None) or, with rows_device_ptr, assembled on meshes[0]'s device (NCCL)."""
    m0 = meshes[0]
    src = _u32(sources)
    off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.uint64)
    B = src.size if off is None else off.size - 1
    hs = (C.c_void_p * len(meshes))(*[m._h for m in meshes])
    st = Stats()
    fn = getattr(_lib.lib(), f"ptp_solve_batched_multi_{m0.suf}")
    if rows_device_ptr is not None:
        check(fn(hs, len(meshes), _p(src), _p(off, C.c_uint64), B, src.size, C.c_void_p(rows_device_ptr), 1, C.byref(st)))
        m0.last_stats = st.as_dict()
        return None
    if rows is None:
        rows = np.empty((B, m0.n_vertices), dtype=m0.dtype)
    assert rows.dtype == m0.dtype and rows.flags.c_contiguous and rows.size == B * m0.n_vertices
    check(fn(hs, len(meshes), _p(src), _p(off, C.c_uint64), B, src.size, C.c_void_p(rows.ctypes.data), 0, C.byref(st)))
    m0.last_stats = st.as_dict()
    return rows


@dataclass
class ptp_out_t:
    """include/geodesics_ptp.h:18-24"""
    dist: np.ndarray
    clusters: np.ndarray | None = None


@dataclass
class toplesets_t:
    """include/geodesics_ptp.h:26-30"""
    limits: np.ndarray
    index: np.ndarray


def parallel_toplesets_propagation_gpu(ptp_out: ptp_out_t, mesh: DeviceMesh, sources, toplesets: toplesets_t) -> float:
    """Fills ptp_out.dist (and .clusters when given) in place; returns elapsed seconds like the reference."""
    _, cl = mesh.solve(sources, toplesets.limits, toplesets.index, clusters=ptp_out.clusters is not None, out=ptp_out.dist)
    if ptp_out.clusters is not None:
        ptp_out.clusters[:] = cl
    return mesh.last_stats["ms_total"] / 1e3


parallel_toplesets_propagation_coalescence_gpu = parallel_toplesets_propagation_gpu


def farthest_point_sampling_ptp_gpu(mesh: DeviceMesh, samples: list, n: int, radio: float = 0.0):
    """-> (max_dist, time_fps seconds); `samples` grows in place like the reference's vector."""
    out, md = mesh.farthest_point_sampling(samples, n, radio)
    samples[:] = [int(x) for x in out]
    return md, mesh.last_stats["ms_total"] / 1e3


class geodesics:
    """include/geodesics.h:18-72 with the PTP_GPU arm only (FM / heat arms are other algorithms, out of scope)."""

    FM, PTP_GPU, HEAT_FLOW_GPU, PTP_CPU, HEAT_FLOW = range(5)  # option_t with GPROSHAN_CUDA defined (geodesics.h:21-28)

    def __init__(self, mesh: DeviceMesh, sources, opt: int = PTP_GPU, e_dist=None, cluster: bool = False,
                 n_iter: int = 0, radio: float = float("inf")):
        if opt != geodesics.PTP_GPU:
            raise NotImplementedError("only option_t::PTP_GPU is provided by this library")
        if len(sources) == 0:
            raise ValueError("sources must be non-empty")  # assert(sources.size() > 0), src/geodesics.cpp:31
        self.n_vertices = mesh.n_vertices
        self.n_sorted = 0  # stays 0 for PTP (src/geodesics.cpp:25, 225-240)
        self.dist = e_dist if e_dist is not None else np.empty(mesh.n_vertices, dtype=mesh.dtype)
        self.sorted_index = np.full(mesh.n_vertices, NIL, dtype=np.uint32)
        _, self.clusters, srt = mesh.geodesics(sources, clusters=cluster, want_sorted=True, out=self.dist)
        n = min(srt.size, mesh.n_vertices)
        self.sorted_index[:n] = srt[:n]
        self.stats = dict(mesh.last_stats)

    def __getitem__(self, i):  # operator[]
        return self.dist[i]

    def __call__(self, i):  # operator()
        return self.sorted_index[i]

    def n_sorted_index(self):
        return self.n_sorted

    def normalize(self):
        """n_sorted == 0 -> normalize_ptp: divide by the largest finite distance (src/geodesics.cpp:76-90)."""
        finite = self.dist[self.dist < np.inf]
        max_d = finite.max() if finite.size else self.dist.dtype.type(0)
        max_d = max(max_d, self.dist.dtype.type(0))
        with np.errstate(divide="ignore", invalid="ignore"):
            self.dist /= max_d
