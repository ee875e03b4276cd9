"""ctypes binding of the C ABI in include/ptp_b200.h (libptp_b200.so, built in-tree by build.py).

Fails loudly when the library is missing: there is no CPU or eager fallback for the PTP path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PTP_B200_LIB: development hook for A/B runs of differently compiled builds of the same library
LIB_PATH = os.environ.get("PTP_B200_LIB") or os.path.join(_HERE, "libptp_b200.so")

PTP_OK = 0
PTP_NIL = 0xFFFFFFFF


class PtpError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"ptp_b200 error {code}: {text}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("n_reached", C.c_uint64), ("n_levels", C.c_uint64), ("iterations", C.c_uint64),
                ("vertex_updates", C.c_uint64), ("max_window", C.c_uint64), ("relaxations", C.c_uint64),
                ("gpu_launches", C.c_uint64),
                ("ms_toplesets", C.c_double), ("ms_solve", C.c_double), ("ms_total", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/ptp_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "ptp_last_error", "ptp_device_count", "ptp_version", "ptp_host_alloc", "ptp_host_free",
    "ptp_set_option", "ptp_get_option", "ptp_option_name", "ptp_option_doc",
    "ptp_mesh_update_positions_f32", "ptp_mesh_update_positions_f64",
    "ptp_mesh_create_f32", "ptp_mesh_create_f64", "ptp_mesh_destroy", "ptp_mesh_last_kernel", "ptp_che_build", "ptp_mesh_n_vertices",
    "ptp_mesh_n_half_edges", "ptp_mesh_real_size", "ptp_mesh_device", "ptp_mesh_device_bytes",
    "ptp_toplesets", "ptp_solve_f32", "ptp_solve_f64", "ptp_geodesics_f32", "ptp_geodesics_f64",
    "ptp_geodesics_error_iter_f32", "ptp_geodesics_error_iter_f64",
    "ptp_solve_batched_f32", "ptp_solve_batched_f64", "ptp_solve_batched_multi_f32", "ptp_solve_batched_multi_f64",
    "ptp_farthest_point_sampling_f32", "ptp_farthest_point_sampling_f64", "ptp_debug_barrier_ns", "ptp_debug_inv_gram_check", "ptp_debug_sign_short_check", "ptp_debug_sqrt_check", "ptp_debug_two_sided_check",
]

_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing — build it with `python -m gproshan_b200.build`; "
                           "the PTP path has no fallback implementation")
    L = C.CDLL(LIB_PATH)
    u32p, u64p, vp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_void_p
    sp = C.POINTER(Stats)
    L.ptp_last_error.restype = C.c_char_p
    L.ptp_version.restype = C.c_char_p
    L.ptp_device_count.restype = C.c_int
    L.ptp_host_alloc.restype = vp
    L.ptp_host_alloc.argtypes = [C.c_size_t]
    L.ptp_host_free.argtypes = [vp]
    L.ptp_host_free.restype = None
    L.ptp_set_option.argtypes = [C.c_char_p, C.c_long]
    L.ptp_get_option.argtypes = [C.c_char_p]
    L.ptp_get_option.restype = C.c_long
    for n in ("ptp_option_name", "ptp_option_doc"):
        getattr(L, n).argtypes = [C.c_int]
        getattr(L, n).restype = C.c_char_p
    for suf, ct in (("f32", C.c_float), ("f64", C.c_double)):
        rp = C.POINTER(ct)
        getattr(L, f"ptp_mesh_update_positions_{suf}").argtypes = [vp, rp]
        f = getattr(L, f"ptp_mesh_create_{suf}")
        f.argtypes = [rp, u32p, u32p, u32p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(vp)]
        f = getattr(L, f"ptp_solve_{suf}")
        f.argtypes = [vp, u32p, C.c_uint32, u32p, C.c_uint32, u32p, rp, u32p, C.c_uint32, sp]
        f = getattr(L, f"ptp_geodesics_{suf}")
        f.argtypes = [vp, u32p, C.c_uint32, rp, u32p, C.c_uint32, u32p, C.c_uint64, sp]
        f = getattr(L, f"ptp_solve_batched_{suf}")
        f.argtypes = [vp, u32p, u64p, C.c_uint32, C.c_uint64, vp, C.c_int, vp, sp]
        f = getattr(L, f"ptp_geodesics_error_iter_{suf}")
        f.argtypes = [vp, u32p, C.c_uint32, rp, rp, u32p, rp, C.c_uint32, u32p, sp]
        f = getattr(L, f"ptp_solve_batched_multi_{suf}")
        f.argtypes = [C.POINTER(vp), C.c_int, u32p, u64p, C.c_uint32, C.c_uint64, vp, C.c_int, sp]
        f = getattr(L, f"ptp_farthest_point_sampling_{suf}")
        f.argtypes = [vp, u32p, C.c_uint32, C.c_uint32, ct, u32p, rp, sp]
    L.ptp_che_build.argtypes = [u32p, C.c_uint64, C.c_uint64, u32p, u32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.ptp_debug_barrier_ns.argtypes = [C.c_int, C.c_int, C.c_int]
    L.ptp_debug_barrier_ns.restype = C.c_double
    L.ptp_debug_sign_short_check.argtypes = [C.c_uint64, C.c_uint64, C.c_int, u64p, u64p, u64p]
    L.ptp_debug_sqrt_check.argtypes = [u64p, u64p]
    L.ptp_debug_two_sided_check.argtypes = [C.c_uint64, C.c_uint64, C.c_int, u64p, u64p, u64p]
    L.ptp_debug_inv_gram_check.argtypes = [C.c_uint64, C.c_uint64, C.c_int, u64p, u64p, C.POINTER(C.c_double)]
    L.ptp_mesh_last_kernel.argtypes = [vp]
    L.ptp_mesh_last_kernel.restype = C.c_char_p
    L.ptp_mesh_destroy.argtypes = [vp]
    L.ptp_mesh_destroy.restype = None
    for n in ("ptp_mesh_n_vertices", "ptp_mesh_n_half_edges", "ptp_mesh_device_bytes"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = C.c_uint64
    for n in ("ptp_mesh_real_size", "ptp_mesh_device"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = C.c_int
    L.ptp_toplesets.argtypes = [vp, u32p, C.c_uint32, C.c_uint32, u32p, u32p, C.c_uint64, u32p, C.c_uint64, u32p, sp]
    _LIB = L
    return L


def check(rc: int):
    if rc != PTP_OK:
        raise PtpError(rc, lib().ptp_last_error().decode(errors="replace"))
