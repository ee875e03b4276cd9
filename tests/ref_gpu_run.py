"""TEST INFRASTRUCTURE. Runs the reference's OWN CUDA PTP (oracle/_ref/libgproshan_ref_cuda_*.so: the reference's
src/cuda/geodesics_ptp*.cu compiled unmodified for sm_100a) in a process of its own — it calls cudaDeviceReset()
(src/cuda/geodesics_ptp.cu:22,89) — and stores what it returns, for the tests that pin this repo's cluster rule,
`newest` buffer option and farthest-point sampling to the code they replace.

    python tests/ref_gpu_run.py clusters <f32|f64> <out.npz>      multi-source solve with Voronoi labels
    python tests/ref_gpu_run.py fps <f32|f64> <n> <radio> <out.npz>  farthest_point_sampling_ptp_gpu
    python tests/ref_gpu_run.py iter_error <f32|f64> <out.npz>    iter_error_parallel_toplesets_propagation_gpu (harness)
The meshes are the seeded ones of `case_mesh` below (the test rebuilds the same)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def case_mesh(kind, dtype):
    from gproshan_b200 import meshgen as mg
    if kind == "clusters":
        m = mg.icosphere(48, noise_sigma=4e-3, seed=77, dtype=dtype)       # 23 042 vertices, irregular: no exact ties
        src = mg.random_sources(11, 9, m.n_vertices, unique=True)
    elif os.environ.get("REF_FPS_F"):  # timing runs (tools/run_fps.py --ref): the plain icosphere of that frequency, sample 0 first
        m = mg.icosphere(int(os.environ["REF_FPS_F"]), dtype=dtype)
        src = np.array([0], dtype=np.uint32)
    else:
        m = mg.icosphere(30, noise_sigma=3e-3, seed=5, dtype=dtype)
        src = np.array([17], dtype=np.uint32)
    return m, src


def exact_sphere(mesh, source):
    """stand-in for the `.exact` files of the reference's harness: great-circle distance on the unit sphere"""
    P = mesh.GT.astype(np.float64)
    u = P / np.linalg.norm(P, axis=1, keepdims=True)
    return np.arccos(np.clip(u @ u[source], -1.0, 1.0))


def ref_cuda_path(dtype):
    return os.path.join(ROOT, "oracle", "_ref", f"libgproshan_ref_cuda_{'f32' if np.dtype(dtype) == np.float32 else 'f64'}.so")


def main():
    import oracle_lib as ol
    mode, dts = sys.argv[1], sys.argv[2]
    dtype = np.float32 if dts == "f32" else np.float64
    ol.ref_path = ref_cuda_path
    ref = ol.Reference(dtype)
    rp, u32p = C.POINTER(ref.ct), ol.u32p
    ref.L.ref_ptp_gpu.argtypes = [C.c_void_p, u32p, C.c_uint32, u32p, C.c_uint32, u32p, rp, u32p]
    ref.L.ref_ptp_gpu.restype = C.c_double
    ref.L.ref_fps_gpu.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_uint32, ref.ct, rp, C.POINTER(C.c_double)]
    ref.L.ref_fps_gpu.restype = C.c_uint32
    ref.L.ref_iter_error_gpu.argtypes = [C.c_void_p, u32p, C.c_uint32, u32p, C.c_uint32, u32p, rp, u32p, rp, C.c_uint32]
    ref.L.ref_iter_error_gpu.restype = C.c_uint32
    mesh, src = case_mesh(mode, dtype)
    rc = ref.che_raw(mesh)
    if mode == "clusters":
        top, srt, lim = rc.compute_toplesets(src)
        dist = np.full(rc.n_v, np.inf, dtype=dtype)
        cl = np.zeros(rc.n_v, dtype=np.uint32)
        srt = np.ascontiguousarray(srt[:rc.n_v])
        ref.L.ref_ptp_gpu(rc.h, ol._p(src), src.size, ol._p(lim), lim.size, ol._p(srt), ol._p(dist, ref.ct), ol._p(cl))
        cpu = rc.ptp_cpu(src, lim, srt)
        np.savez(sys.argv[3], dist=dist, clusters=cl, cpu=cpu, sources=src)
    elif mode == "iter_error":
        top, srt, lim = rc.compute_toplesets(src)
        srt = np.ascontiguousarray(srt[:rc.n_v])
        exact = exact_sphere(mesh, int(src[0])).astype(dtype)
        it = np.zeros(4096, dtype=np.uint32)
        er = np.zeros(4096, dtype=dtype)
        n = ref.L.ref_iter_error_gpu(rc.h, ol._p(src), src.size, ol._p(lim), lim.size, ol._p(srt), ol._p(exact, ref.ct), ol._p(it), ol._p(er, ref.ct), 4096)
        np.savez(sys.argv[3], iters=it[:n].copy(), errors=er[:n].copy(), n_limits=np.array(lim.size))
    else:
        n, radio = int(sys.argv[3]), float(sys.argv[4])
        buf = np.zeros(max(n, src.size, mesh.n_vertices), dtype=np.uint32)
        buf[:src.size] = src
        md, secs = ref.ct(0), C.c_double(0)
        t = time.perf_counter()
        cnt = ref.L.ref_fps_gpu(rc.h, ol._p(buf), src.size, n, ref.ct(radio), C.byref(md), C.byref(secs))
        np.savez(sys.argv[5], samples=buf[:cnt].copy(), max_dist=np.array(md.value), seconds=np.array(secs.value),
                 wall=np.array(time.perf_counter() - t))


if __name__ == "__main__":
    main()
