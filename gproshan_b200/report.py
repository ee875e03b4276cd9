"""Benchmark / accuracy harness of the PTP path — the part of gproshan's `test_geodesics` executable
(src/test_geodesics_ptp.cpp:16-241) that concerns PTP on the GPU, driven through this repo's API.

For every mesh it writes what the reference writes (same file names, same line formats):
  ptp_results.tex / ptp_results_double.tex   one LaTeX row per mesh (:103-156): name, |V|, then the PTP GPU cell
                                              `& time & (speed-up) & error%` — the cells of the arms that are out of
                                              scope here (fast marching, PTP CPU, heat method) are left as `--`
  <name>.deg                                  degree histogram (:159-172)
  <name>_toplesets.dist, _toplesets_sorted.dist   topleset sizes, in order and sorted (:175-193)
  <name>.fps                                  farthest-point-sampling time against the number of samples (:218-229;
                                              cumulative device time at a few sample counts instead of per sample)
`compute_error` is the reference's (:361-370): 100 / (n - s) * sum over exact > 0 of |dist - exact| / exact.
  <name>_error.iter / _error_double.iter      per-iteration error (:198-214; iter_error_run_ptp_gpu,
                                              src/cuda/test_geodesics_ptp.cu:164-211) from ptp_geodesics_error_iter_*: the
                                              sums are formed on the device, no per-iteration copy of the distances

Exact distances: `<name>.exact` files as the reference reads them (:345-359, one value per vertex), or — for the
synthetic meshes of this repo — the analytic distance on the smooth surface (great circle on the unit sphere,
Euclidean on the planar grid).
"""
from __future__ import annotations

import os

import numpy as np

from . import api


def compute_error(dist: np.ndarray, exact: np.ndarray, n_sources: int) -> float:
    """src/test_geodesics_ptp.cpp:361-370"""
    dist, exact = np.asarray(dist, dtype=np.float64), np.asarray(exact, dtype=np.float64)
    m = exact > 0
    return float((np.abs(dist[m] - exact[m]) / exact[m]).sum() * 100.0 / (exact.size - n_sources))


def load_exact_geodesics(path: str, n: int):
    """src/test_geodesics_ptp.cpp:345-359: None when the file is missing"""
    if not os.path.exists(path):
        return None
    return np.loadtxt(path, dtype=np.float64).reshape(-1)[:n]


def degree_histogram(mesh) -> dict:
    """:159-172 — degree = triangles in the star, +1 for border vertices (ot_evt(v) == NIL)"""
    star = np.bincount(mesh.VT, minlength=mesh.n_vertices)
    evt = mesh.EVT
    border = np.zeros(mesh.n_vertices, dtype=np.int64)
    ok = evt != api.NIL
    border[ok] = (mesh.OT[evt[ok]] == api.NIL).astype(np.int64)
    deg, cnt = np.unique(star + border, return_counts=True)
    return {int(d): int(c) for d, c in zip(deg, cnt)}


def toplesets_distribution(limits: np.ndarray):
    """:175-193 -> (sizes in topleset order, sizes sorted ascending)"""
    sizes = np.diff(np.asarray(limits, dtype=np.int64))
    return sizes, np.sort(sizes)


def latex_row(name: str, n_vertices: int, ptp_gpu_s: float, ptp_gpu_err: float, fm_s: float | None = None, double: bool = False) -> str:
    """One row of ptp_results[_double].tex with the reference's printf formats (:36-40, :103-156). Arms not measured
    here are `--`; the speed-up is against fast marching like the reference's, when a time for it is supplied."""
    dash_t, dash_e = "& %6s %6ss " % ("", "--"), "& %6s %6s\\%% " % ("", "--")
    speed = "& \\bf (%.1lfx) " % (fm_s / ptp_gpu_s) if fm_s else "& \\bf (--) "
    cell = "& %6s %6.3lfs " % ("\\bf", ptp_gpu_s) + speed + "& %6s %6.2lf\\%% " % ("\\bf", ptp_gpu_err)
    row = "%20s " % ("\\verb|" + name + "|") + "& %12lu " % n_vertices
    row += dash_t + dash_e                      # FM
    row += dash_t + "& \\bf (--) " + dash_e     # PTP CPU
    if not double:
        return row + cell + "\\\\\n"
    row += "& OpenMP " + "& %6ss " % "--" + dash_t + "& \\bf (--) " + dash_e + "& Cholmod \\\\\n"
    row += "&&& " + cell + "& Cuda " + "& %6ss " % "--" + dash_t + "& \\bf (--) " + dash_e + "& cusolverSp \\\\\\hline\n"
    return row


def analytic_exact(kind: str, mesh, source: int) -> np.ndarray:
    P = mesh.GT.astype(np.float64)
    if kind == "sphere":
        u = P / np.linalg.norm(P, axis=1, keepdims=True)
        return np.arccos(np.clip(u @ u[source], -1.0, 1.0))
    if kind == "plane":
        return np.linalg.norm(P - P[source], axis=1)
    raise ValueError(kind)


def run(meshes, out_dir: str, n_test: int = 10, device: int = 0, fps_counts=(2, 4, 8, 16, 32, 64)) -> list:
    """meshes: iterable of (name, CheMesh, exact array or None). Source = vertex 0 like the reference (:49).
    Returns one dict per mesh; writes the files listed in the module docstring into out_dir."""
    os.makedirs(out_dir, exist_ok=True)
    results = []
    tex = {}
    for name, mesh, exact in meshes:
        double = mesh.GT.dtype == np.float64
        source = [0]
        with api.DeviceMesh(mesh, device) as dm:
            top, srt, lim = dm.compute_toplesets(source)
            best, dist = float("inf"), None
            for _ in range(n_test):                      # min over n_test runs, like test_ptp_gpu (:298-314)
                dist, _, _ = dm.geodesics(source)
                best = min(best, dm.last_stats["ms_total"] / 1e3)
            stats = dict(dm.last_stats)
            iter_err = dm.error_per_iteration(source, exact)[:2] if exact is not None else None
            fps = []
            for n in fps_counts:
                if n < mesh.n_vertices // 2:
                    dm.farthest_point_sampling(source, n)
                    fps.append((n, dm.last_stats["ms_total"] / 1e3))
        err = compute_error(dist, exact, len(source)) if exact is not None else float("nan")
        tex.setdefault(double, []).append(latex_row(name, mesh.n_vertices, best, err, double=double))
        with open(os.path.join(out_dir, name + ".deg"), "w") as f:
            for d, c in sorted(degree_histogram(mesh).items()):
                f.write(f"{d} {c}\n")
        sizes, ssorted = toplesets_distribution(lim)
        with open(os.path.join(out_dir, name + "_toplesets.dist"), "w") as f:
            f.writelines(f"{i} {s}\n" for i, s in enumerate(sizes))
        with open(os.path.join(out_dir, name + "_toplesets_sorted.dist"), "w") as f:
            f.writelines(f"{i} {s}\n" for i, s in enumerate(ssorted))
        if iter_err is not None:
            with open(os.path.join(out_dir, name + ("_error_double.iter" if double else "_error.iter")), "w") as f:
                f.writelines(f"{int(i)} {e}\n" for i, e in zip(*iter_err))
        with open(os.path.join(out_dir, name + ".fps"), "w") as f:
            f.writelines(f"{n} {t}\n" for n, t in fps)
        results.append({"name": name, "n_vertices": mesh.n_vertices, "seconds": best, "error_pct": err,
                        "levels": int(lim.size - 1), "iterations": stats["iterations"], "kernel": None})
    for double, rows in tex.items():
        with open(os.path.join(out_dir, "ptp_results_double.tex" if double else "ptp_results.tex"), "w") as f:
            f.writelines(rows)
    return results
