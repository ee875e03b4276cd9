#!/usr/bin/env python
"""PTP geodesics benchmark (contract: one JSON line on stdout from rank 0).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the reference's own CPU PTP on the host cores
    torchrun ... bench.py --gpus N ...                        # one process per GPU (batched sources sharded)

Workloads (BASELINE.json configs):
  batched  C5: icosphere f=447 (1 998 092 vertices, float), the 1024 independent single-source solves of the
           distance-matrix job at EVERY N: 1024 / N sources per GPU (strong scaling), rows gathered with NCCL to the
           full 1024 x V matrix. This is the line's `metric` / `value`.
  single   C3: noisy icosphere f=1000 (10 000 002 vertices, double), one source; reported on the N=1 line under
           "single_source" (ms/solve, vertex-updates/s, its own roofline, e2e and CPU baseline).

A step = one pass of the hot path over the whole job: one batched call over this rank's sources + the gather.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_UPDATE = {4: 72, 8: 92}  # 52 + 5*sizeof(real) algorithmic bytes per vertex-update


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ncu_traffic(kernel_key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from THIS round's
    `ncu --set full` capture of the same workload (profiles/r2_traffic.json: {key: {"bytes": .., "commit": ..,
    "command": ..}}, written by tools/ncu_traffic.py); None when the kernel has not been captured this round."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        e = json.load(open(p)).get(kernel_key)
        return e if e is None else e["bytes"]
    except Exception:
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def pin_openmp(n):
    """The CPU arms run the reference's OpenMP code on ALL host threads, whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1 for its workers). The variable covers libgomp instances not loaded yet, omp_set_num_threads the
    one already in the process. -> omp_get_max_threads() as seen afterwards."""
    import ctypes
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        g = ctypes.CDLL("libgomp.so.1")
        g.omp_set_num_threads(int(n))
        return int(g.omp_get_max_threads())
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark(self):
        """the timed region starts here: only samples taken from now on are reported"""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines[getattr(self, "first", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workloads

def workload_meshes(quick):
    from gproshan_b200 import meshgen as mg
    f5, f3 = (60, 100) if quick else (447, 1000)
    return mg, f5, f3


def build_c5(quick):
    mg, f5, _ = workload_meshes(quick)
    t = time.perf_counter()
    mesh = mg.icosphere(f5, dtype=np.float32)
    srcs = mg.random_sources(1024, 1024, mesh.n_vertices, unique=True)
    log(f"[bench] C5 icosphere f={f5}: {mesh.n_vertices} vertices, {srcs.size} unique sources, built in {time.perf_counter()-t:.1f}s")
    return mesh, srcs, f"C5 batched single-source solves, icosphere f={f5} V={mesh.n_vertices} float"


def build_c3(quick):
    mg, _, f3 = workload_meshes(quick)
    t = time.perf_counter()
    sigma = 0.2 * mg.mean_edge_icosphere(f3)
    mesh = mg.icosphere(f3, noise_sigma=sigma, seed=12345, dtype=np.float64)
    log(f"[bench] C3 noisy icosphere f={f3}: {mesh.n_vertices} vertices, sigma={sigma:.3e}, built in {time.perf_counter()-t:.1f}s")
    return mesh, np.array([0], dtype=np.uint32), f"C3 single-source, noisy icosphere f={f3} V={mesh.n_vertices} double"


def cpu_runner(dtype):
    """The CPU arm: the reference's own code (oracle/_ref) when it was built, else the oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    if ol.ref_available(dtype):
        ref = ol.Reference(dtype)

        def make(mesh):
            rc = ref.che_raw(mesh)

            def solve(src):
                top, srt, lim = rc.compute_toplesets(src)
                return rc.ptp_cpu(src, lim, srt), lim, srt
            return solve
        return "reference", make
    orc = ol.Oracle()

    def make(mesh):
        def solve(src):
            top, srt, lim = orc.compute_toplesets(mesh, src)
            return orc.ptp_cpu(mesh, src, lim, srt)[0], lim, srt
        return solve
    return "port", make


def cpu_sources_per_s(mesh, srcs, budget_s, max_n):
    """-> (kind, rows solved on the CPU (list of arrays), seconds)"""
    kind, make = cpu_runner(mesh.GT.dtype)
    solve = make(mesh)
    rows, t0 = [], time.perf_counter()
    while len(rows) < max_n and (not rows or time.perf_counter() - t0 < budget_s):
        rows.append(solve(srcs[len(rows):len(rows) + 1])[0])
    dt = time.perf_counter() - t0
    return kind, rows, dt


N_SOURCES = 1024  # BASELINE.json config 5: the distance-matrix job


def c5_config(wl, world):
    """The line's `config`: identical for both arms (the reference arm times a bounded sample of the same job)."""
    per = N_SOURCES // world
    return {"workload": wl, "sources_total": per * world, "sources_per_gpu": per, "gpus": world,
            "sharding": ("sources block-partitioned over the GPUs, mesh replicated, rows gathered with NCCL (all_gather) into the "
                         "full matrix on every GPU" if world > 1 else "single GPU"),
            "l2": "inputs larger than L2: per step 1024 x V distance rows (8.2 GB) are produced and every solve rebuilds its "
                  "own workspace; no explicit flush"}


def reference_gpu_leg(workload, quick, sources, coalescence):
    """Second baseline: the reference's own CUDA PTP (unmodified, sm_100a build in oracle/_ref) on this GPU, in a
    process of its own (it calls cudaDeviceReset()). See tools/ref_gpu_bench.py for what is timed."""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "ref_gpu_bench.py"), "--workload", workload, "--sources", str(sources)]
    if quick:
        cmd.append("--quick")
    if coalescence:
        cmd.append("--coalescence")
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": f"exit {r.returncode}: {r.stderr.strip()[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:  # the baseline must never take the bench line down
        return {"unavailable": repr(e)}


# ------------------------------------------------------------------------------------------------ arms

def run_reference(args, rank, world):
    """--impl reference: the reference's CPU PTP (compute_toplesets + parallel_toplesets_propagation_cpu) on the
    host cores, same config and metric; each step is a bounded sample (`--ref-sources-per-step` sources) of the job.
    Rank 0 only; the OpenMP thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    if rank != 0:
        return
    cores = host_threads()
    omp = pin_openmp(cores)
    mesh, srcs, wl = build_c5(args.quick)
    kind, make = cpu_runner(mesh.GT.dtype)
    solve = make(mesh)
    per_step = max(1, args.ref_sources_per_step)
    k = 0
    for _ in range(args.warmup):
        solve(srcs[k:k + 1]); k += 1
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        for _ in range(per_step):
            solve(srcs[k % srcs.size:k % srcs.size + 1]); k += 1; done += 1
        if time.perf_counter() - t0 > args.ref_budget_s:
            break
    dt = time.perf_counter() - t0
    steps_done = max(1, done // per_step)
    val = done / dt
    line = {
        "impl": "reference", "metric": "ptp_batched_sources_per_s", "value": val, "unit": "sources/s", "n_gpus": args.gpus,
        "steps": steps_done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps_done, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": c5_config(wl, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": "sources/s", "cores": omp or cores, "kind": kind,
                         "sample": f"{done} of the {N_SOURCES} single-source solves ({per_step} per step; compute_toplesets + "
                                   f"parallel_toplesets_propagation_cpu), OpenMP: omp_get_max_threads() = {omp} on {cores} host threads"},
        "e2e": {"value": val, "unit": "sources/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch
    from gproshan_b200 import api

    if not torch.cuda.is_available() or api.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — the PTP path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peak, peak_src = measured_peaks()

    # ---------------- batched (the line's metric): the whole 1024-source job at every N
    mesh, srcs_all, wl = build_c5(args.quick)
    V = mesh.n_vertices
    from gproshan_b200.sharding import gather_rows, shard_sources
    n_total = min(args.sources, srcs_all.size) // world * world
    per_gpu = n_total // world
    mine = np.ascontiguousarray(shard_sources(srcs_all[:n_total], rank, world, per_rank=per_gpu))
    t = time.perf_counter()
    dm = api.DeviceMesh(mesh, device=local_rank)
    torch.cuda.synchronize()
    upload_s = time.perf_counter() - t
    rows = torch.empty((per_gpu, V), dtype=torch.float32, device="cuda")
    gathered = torch.empty((n_total, V), dtype=torch.float32, device="cuda") if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    def step_resident():
        dm.solve_batched(mine, rows_device_ptr=rows.data_ptr(), stream=stream)
        if world > 1:
            gather_rows(rows, world, out=gathered)
        return dm.last_stats

    # (the sampler is started BEFORE the warm-up: nvidia-smi's own start-up — NVML initialisation, ~0.3 s, during which CUDA
    # calls of this process can stall — must not fall into the timed region; only samples taken after mark() are reported)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kern_ms, updates, launches = [], 0, 0
    for _ in range(args.steps):
        st = step_resident()
        kern_ms.append(st["ms_total"]); updates = st["vertex_updates"]; launches += st["gpu_launches"]
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    upd = torch.tensor([float(updates)], device="cuda")
    if dist:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(upd, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    ms_step = ms_total / args.steps
    value = n_total / (ms_step / 1e3)

    # e2e: host sources in, the assembled n_total x V matrix out in HOST memory on rank 0, every step.
    #   N = 1: the C-ABI call with host pointers (H2D of the sources, D2H of the rows inside the call), pinned rows.
    #   N > 1: the matrix lives in a shared-memory host buffer (/dev/shm, mapped by every rank, each rank's slice pinned);
    #          every rank makes the same C-ABI call with a host pointer to ITS slice, so each GPU copies its rows over its
    #          own PCIe link and the assembled matrix is complete on the host when the last rank's call returns. No
    #          collective on this path (the NCCL gather is part of `value`, where the matrix stays on the GPUs). If the shared
    #          buffer cannot be set up: per-rank solve, NCCL all_gather, rank 0 copies the assembled matrix to its host.
    e2e_steps = max(1, min(args.steps, 3))
    lo = rank * per_gpu
    shared, mine_slice, shm_path, e2e_mode = None, None, None, "single"
    if world > 1:
        shm_path = f"/dev/shm/ptp_b200_rows_{os.environ.get('MASTER_PORT', '0')}"
        ok = torch.zeros(1, device="cuda")
        try:
            if rank == 0:
                with open(shm_path, "wb") as f:
                    f.truncate(n_total * V * 4)
            dist.barrier()
            shared = torch.from_file(shm_path, shared=True, size=n_total * V, dtype=torch.float32).view(n_total, V)
            mine_slice = shared[lo:lo + per_gpu]
            rc = torch.cuda.cudart().cudaHostRegister(mine_slice.data_ptr(), mine_slice.numel() * 4, 0)
            if int(rc) != 0:
                log(f"[bench] rank {rank}: cudaHostRegister of the shared slice failed ({rc}); rows go through pageable memory")
            ok += 1
        except Exception as e:  # no /dev/shm, not enough room, ...
            log(f"[bench] rank {rank}: shared host matrix unavailable ({e!r})")
        dist.all_reduce(ok)
        e2e_mode = "shared-host" if int(ok.item()) == world else "nccl-then-d2h"
    host_rows = torch.empty((n_total, V), dtype=torch.float32, pin_memory=True) if rank == 0 and e2e_mode != "shared-host" else None

    def step_e2e():
        if e2e_mode == "single":
            dm.solve_batched(mine, rows=host_rows.numpy())
            return float(host_rows[0, :8].sum())
        if e2e_mode == "shared-host":
            dm.solve_batched(mine, rows=shared[lo:lo + per_gpu].numpy())
            dist.barrier()  # every rank's rows have landed in the host matrix
            return float(shared[0, :8].sum()) + float(shared[n_total - 1, :8].sum()) if rank == 0 else 0.0
        dm.solve_batched(mine, rows_device_ptr=rows.data_ptr(), stream=stream)
        gather_rows(rows, world, out=gathered)
        if rank == 0:
            host_rows.copy_(gathered, non_blocking=True)
        torch.cuda.synchronize()
        return float(host_rows[0, :8].sum()) if rank == 0 else 0.0

    step_e2e()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(e2e_steps):
        checksum = step_e2e()
        if dist:
            dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t], device="cuda")
    if dist:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = n_total * e2e_steps / float(e2e_s.item())
    e2e_what = {"single": "C-ABI call, host sources in, pinned host rows out",
                "shared-host": "per rank: the C-ABI call with host pointers — sources in, its slice of ONE shared host matrix (mapped by all ranks, "
                               "owned by rank 0) out over its own PCIe link; no collective on this path (the NCCL gather is in `value`)",
                "nccl-then-d2h": "per rank: host sources in, rows on device; NCCL all_gather; rank 0 copies the assembled matrix to pinned host memory"}[e2e_mode]
    if shared is not None:
        e2e_parity = None
        if rank == 0:  # the assembled host matrix against the device-resident gathered one of the timed steps
            e2e_parity = bool(torch.equal(shared[:64], gathered[:64].cpu())) and bool(torch.equal(shared[-64:], gathered[-64:].cpu()))
        try:
            torch.cuda.cudart().cudaHostUnregister(shared[lo:lo + per_gpu].data_ptr())
        except Exception:
            pass
        mine_slice = None
        shared = None
        dist.barrier()
        if rank == 0:
            try:
                os.unlink(shm_path)
            except OSError:
                pass
    else:
        e2e_parity = None

    kernel_ms = statistics.mean(kern_ms)
    kernel = dm.last_kernel
    achieved = updates * BYTES_PER_UPDATE[4] / (kernel_ms / 1e3) / 1e9
    line = {
        "metric": "ptp_batched_sources_per_s", "value": value, "unit": "sources/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": c5_config(wl, world) if n_total == N_SOURCES else dict(c5_config(wl, world), sources_total=n_total, sources_per_gpu=per_gpu),
        "vertex_updates_per_step_all_gpus": float(upd.item()), "mesh_upload_s": upload_s,
        "vertex_updates_per_s": float(upd.item()) / (ms_step / 1e3),
        "e2e": {"value": e2e_val, "unit": "sources/s", "h2d_bytes_per_step": int(mine.nbytes) * world,
                "d2h_bytes_per_step": int(n_total) * V * 4, "steps": e2e_steps, "checksum": checksum,
                "what": e2e_what, "host_matrix_equals_gathered": e2e_parity},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic("c5_batched_f32"), "peak_source": peak_src,
                     "algorithmic_bytes_per_vertex_update": BYTES_PER_UPDATE[4], "vertex_updates_per_launch": updates,
                     "achieved_counting_executed_relaxations_only": st["relaxations"] * BYTES_PER_UPDATE[4] / (kernel_ms / 1e3) / 1e9,
                     "relaxations_per_launch": st["relaxations"], "kernel_ms": kernel_ms,
                     "note": "rank 0's launch (its share of the job); includes BFS + layout + sweep of every source in the launch. "
                             "vertex_updates = the reference schedule's window sizes (skipped relaxations included)"},
        "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference's CPU PTP on a bounded sample of the SAME sources; its rows are the parity check of the GPU rows
        cores = host_threads()
        omp = pin_openmp(cores)
        kind, cpu_rows, dt = cpu_sources_per_s(mesh, mine, args.cpu_budget_s, 16)
        n = len(cpu_rows)
        got = rows[:n].cpu().numpy()
        worst, equal = 0.0, True
        for k in range(n):
            fin = np.isfinite(cpu_rows[k])
            equal = equal and bool(np.array_equal(got[k], cpu_rows[k]))
            if fin.any():
                worst = max(worst, float((np.abs(got[k][fin] - cpu_rows[k][fin]) / np.maximum(cpu_rows[k][fin], 1e-30)).max()))
        line["cpu_baseline"] = {"value": n / dt, "unit": "sources/s", "cores": omp or cores, "kind": kind,
                                "sample": f"{n} of the {n_total} sources of this workload, {dt:.1f}s, OpenMP: omp_get_max_threads() = {omp}"}
        line["parity"] = {"rows_checked": n, "bit_equal": equal, "max_rel_err": worst, "tolerance": 1e-5,
                          "against": "the reference's CPU PTP (oracle/_ref) on the same sources" if kind == "reference" else "the oracle port"}
    del rows, gathered, host_rows
    if rank == 0 and world == 1 and args.workload in ("auto", "both", "single"):
        dm.close()
        torch.cuda.empty_cache()
        line["single_source"] = run_single(args, api, torch, peak, peak_src)
    dm.close()
    if rank == 0 and world == 1 and not args.no_ref_gpu:
        torch.cuda.synchronize()
        if "single_source" in line:
            line["single_source"]["reference_gpu"] = reference_gpu_leg("c3", args.quick, 3, True)
        line["reference_gpu"] = reference_gpu_leg("c5", args.quick, 3, False)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_single(args, api, torch, peak, peak_src):
    mesh, src, wl = build_c3(args.quick)
    V = mesh.n_vertices
    t = time.perf_counter()
    dm = api.DeviceMesh(mesh, device=torch.cuda.current_device())
    upload_s = time.perf_counter() - t
    out = torch.empty(V, dtype=torch.float64, pin_memory=True).numpy()
    steps = max(3, min(args.steps, 10))
    for _ in range(max(1, min(args.warmup, 3))):
        dm.geodesics(src, out=out)
    dev_ms, top_ms, sol_ms, wall = [], [], [], []
    for _ in range(steps):
        t = time.perf_counter()
        dm.geodesics(src, out=out)            # host sources in, host distances out: the e2e call
        wall.append((time.perf_counter() - t) * 1e3)
        st = dm.last_stats
        dev_ms.append(st["ms_total"]); top_ms.append(st["ms_toplesets"]); sol_ms.append(st["ms_solve"])
    st = dm.last_stats
    kernel = dm.last_kernel
    ms = statistics.median(dev_ms)
    sweep_ms = statistics.median(sol_ms)
    achieved = st["vertex_updates"] * BYTES_PER_UPDATE[8] / (ms / 1e3) / 1e9
    t = time.perf_counter()
    _, _, manifold, che_ms = api.che_build(mesh.VT, V, torch.cuda.current_device())
    che_wall = time.perf_counter() - t
    res = {
        "workload": wl, "kernel": kernel, "ms_per_solve": ms, "ms_bfs_team": statistics.median(top_ms), "ms_until_sweep_team_done": sweep_ms,
        "che_build": {"device_ms": che_ms, "ms_with_h2d_d2h": che_wall * 1e3, "manifold": manifold,
                      "note": "OT/EVT from the face list on the device (reference: che::update_evt_ot_et, serial)"},
        "vertex_updates": st["vertex_updates"], "relaxations": st["relaxations"], "iterations": st["iterations"], "levels": st["n_levels"],
        "max_window": st["max_window"], "vertex_updates_per_s": st["vertex_updates"] / (ms / 1e3),
        "e2e": {"ms_per_solve": statistics.median(wall), "h2d_bytes": 4, "d2h_bytes": int(out.nbytes)},
        "mesh_upload_s": upload_s, "steps": steps, "gpu_launches_per_solve": st["gpu_launches"],
        "roofline": {"bound": "hbm", "kernel": kernel + " (BFS team + sweep team, one launch)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic("c3_single_f64"), "peak_source": peak_src,
                     "algorithmic_bytes_per_vertex_update": BYTES_PER_UPDATE[8],
                     "note": "latency-bound by construction: ~#toplesets dependent iterations (SURVEY.md §0.4)"},
    }
    if not args.no_cpu_baseline:
        omp = pin_openmp(host_threads())
        kind, make = cpu_runner(np.float64)
        solve = make(mesh)
        t = time.perf_counter()
        ref, lim, srt = solve(src)
        cpu_s = time.perf_counter() - t
        rel = np.abs(out - ref)[np.isfinite(ref)] / np.maximum(ref[np.isfinite(ref)], 1e-300)
        res["cpu_baseline"] = {"value": cpu_s * 1e3, "unit": "ms/solve", "cores": omp or host_threads(), "kind": kind,
                               "sample": "1 full solve (compute_toplesets + parallel_toplesets_propagation_cpu)"}
        res["parity_vs_cpu"] = {"max_rel_err": float(rel.max()), "bit_equal": bool(np.array_equal(out, ref))}
    dm.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "single", "batched", "both"])
    ap.add_argument("--sources", type=int, default=N_SOURCES, help="sources of the batched job (default: the full 1024)")
    ap.add_argument("--quick", action="store_true", help="small meshes (smoke / CI)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference's own CUDA PTP (second baseline)")
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--ref-sources-per-step", type=int, default=1)
    ap.add_argument("--ref-budget-s", type=float, default=150.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus:
            log(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
