"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + stall mix + hottest SASS lines.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_top=25]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
        "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_xu.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_inst0.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:100])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:62s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
ns = 0
for r in data:
    n = int(r[ix["# Samples"]] or 0)
    ns += n
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
print(f"== stall mix ({ns} samples)")
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:9]:
    print(f"  {s:28s} {100 * v / max(ns, 1):5.1f}%")
print("== hottest instructions (samples, executed, SASS, stalls)")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]:
    st = " ".join(f"{s[6:]}={r[ix[s]]}" for s in stalls if r[ix[s]] not in ("0", ""))
    print(f"  {r[ix['# Samples']]:>8s} {r[ix['Instructions Executed']]:>11s}  {r[ix['Source']].strip()[:70]:70s} | {st}")

# dynamic opcode mix: warp-level executed instructions per opcode (and the share of thread-level work)
mix, tmix = {}, {}
iT = ix.get("Thread Instructions Executed")
for r in data:
    op = r[ix["Source"]].strip().split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    o = o.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    mix[o] = mix.get(o, 0) + n
    if iT is not None:
        tmix[o] = tmix.get(o, 0) + int(r[iT] or 0)
tot = sum(mix.values()) or 1
print(f"== opcode mix ({tot} warp instructions)")
for o, n in sorted(mix.items(), key=lambda x: -x[1])[:28]:
    extra = f"  lanes/inst {tmix[o] / n:5.1f}" if iT is not None and n else ""
    print(f"  {o:12s} {100 * n / tot:5.1f}%{extra}")
