#!/usr/bin/env python
"""Second baseline: the reference's OWN CUDA PTP (unmodified sources, compiled for sm_100a by `make -C oracle
refgpu` into oracle/_ref/libgproshan_ref_cuda_{f32,f64}.so) run on this box's GPU, on the bench workloads.

BENCH INFRASTRUCTURE, not product code. Runs in a process of its own (no torch, no libptp_b200): the reference
calls cudaDeviceReset() at the top of every solve (src/cuda/geodesics_ptp.cu:22). bench.py launches it as a
subprocess and copies the JSON line it prints into `single_source.reference_gpu` / `reference_gpu`.

What is timed (per solve, as the reference's `geodesics` class runs it, src/geodesics.cpp:223-239):
  toplesets_cpu_ms  che::compute_toplesets on the host (the reference has no device BFS)
  gpu_ms            the reference's own CUDA-event timer around upload + window loop + download
  wall_ms           wall clock around the whole parallel_toplesets_propagation[_coalescence]_gpu call
                    (the coalescence arm builds a re-ordered che on the host inside the call, outside its timer)
Distances are compared with the reference CPU PTP of the same library (max relative error; the reference's own
GPU kernels are compiled with FMA contraction, so they are close to, not bit-equal with, its CPU path).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def lib_path(dtype):
    return os.path.join(ROOT, "oracle", "_ref", f"libgproshan_ref_cuda_{'f32' if np.dtype(dtype) == np.float32 else 'f64'}.so")


def bind(dtype):
    import oracle_lib as ol
    ol.ref_path = lib_path            # same C driver + the GPU entry points, one library
    ref = ol.Reference(dtype)
    rp, u32p = C.POINTER(ref.ct), ol.u32p
    for n in ("ref_ptp_gpu", "ref_ptp_coalescence_gpu"):
        f = getattr(ref.L, n)
        f.argtypes = [C.c_void_p, u32p, C.c_uint32, u32p, C.c_uint32, u32p, rp, u32p]
        f.restype = C.c_double
    return ol, ref


def gpu_solve(ol, ref, rc, src, lim, srt, coalescence):
    dist = np.full(rc.n_v, np.inf, dtype=ref.dt)
    f = ref.L.ref_ptp_coalescence_gpu if coalescence else ref.L.ref_ptp_gpu
    srt = np.ascontiguousarray(srt[:rc.n_v])
    t = time.perf_counter()
    sec = f(rc.h, ol._p(src), src.size, ol._p(lim), lim.size, ol._p(srt), ol._p(dist, ref.ct), None)
    return dist, sec * 1e3, (time.perf_counter() - t) * 1e3


def rel_err(a, b):
    fin = np.isfinite(b)
    if not np.array_equal(fin, np.isfinite(a)):
        return float("inf")
    d, ref = np.abs(a[fin] - b[fin]), b[fin]
    nz = ref > 0                       # sources: 0 vs 0
    if not np.array_equal(a[fin][~nz], ref[~nz]):
        return float("inf")
    return float((d[nz] / ref[nz]).max()) if nz.any() else 0.0


def run(workload, quick, n_sources, coalescence, check):
    from gproshan_b200 import meshgen as mg
    if workload == "c3":
        f = 100 if quick else 1000
        mesh = mg.icosphere(f, noise_sigma=0.2 * mg.mean_edge_icosphere(f), seed=12345, dtype=np.float64)
        srcs = np.array([0], dtype=np.uint32)
    else:
        f = 60 if quick else 447
        mesh = mg.icosphere(f, dtype=np.float32)
        srcs = mg.random_sources(1024, 1024, mesh.n_vertices, unique=True)
    ol, ref = bind(mesh.GT.dtype)
    rc = ref.che_raw(mesh)
    out = {"workload": workload, "V": mesh.n_vertices, "dtype": "f32" if mesh.GT.dtype == np.float32 else "f64",
           "impl": "reference CUDA sources (src/cuda/geodesics_ptp*.cu), unmodified, sm_100a", "solves": []}
    # warm-up: CUDA context + module load of the reference library are not charged to its first solve
    top, srt, lim = rc.compute_toplesets(srcs[:1])
    gpu_solve(ol, ref, rc, srcs[:1], lim, srt, False)
    for k in range(n_sources):
        src = np.ascontiguousarray(srcs[k % srcs.size:k % srcs.size + 1])  # (c3 has one source: repeated solves)
        t = time.perf_counter()
        top, srt, lim = rc.compute_toplesets(src)
        top_ms = (time.perf_counter() - t) * 1e3
        d, gpu_ms, wall_ms = gpu_solve(ol, ref, rc, src, lim, srt, False)
        rec = {"toplesets_cpu_ms": top_ms, "gpu_ms": gpu_ms, "wall_ms": wall_ms, "levels": int(lim.size - 1)}
        if coalescence and k == 0:
            dc, g2, w2 = gpu_solve(ol, ref, rc, src, lim, srt, True)
            rec["coalescence"] = {"gpu_ms": g2, "wall_ms": w2, "max_rel_vs_plain_gpu": rel_err(dc, d)}
        if check and k == 0:
            t = time.perf_counter()
            cpu = rc.ptp_cpu(src, lim, srt)
            rec["cpu_ptp_ms"] = (time.perf_counter() - t) * 1e3
            rec["max_rel_err_vs_reference_cpu"] = rel_err(d, cpu)
        out["solves"].append(rec)
    g = [s["gpu_ms"] for s in out["solves"]]
    tt = [s["toplesets_cpu_ms"] + s["wall_ms"] for s in out["solves"]]
    out["gpu_ms_per_solve"] = float(np.median(g))
    out["gpu_ms_best"] = float(min(g))
    out["ms_per_solve_with_cpu_toplesets"] = float(np.median(tt))
    out["sources_per_s"] = 1e3 * len(tt) / float(sum(tt))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"])
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--sources", type=int, default=1)
    ap.add_argument("--coalescence", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    if not os.path.exists(lib_path(np.float64)):
        print(json.dumps({"unavailable": "oracle/_ref/libgproshan_ref_cuda_*.so not built (make -C oracle refgpu)"}))
        return
    print(json.dumps(run(a.workload, a.quick, a.sources, a.coalescence, not a.no_check)), flush=True)


if __name__ == "__main__":
    main()
