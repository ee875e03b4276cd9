export PATH=/usr/local/cuda/bin:$PATH
echo ==== BENCH
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1_a.json'))
s=d.get('single_source',{})
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','roofline','cpu_baseline','parity','clocks') if k in d})
print({k:s[k] for k in ('kernel','ms_per_solve','ms_bfs_team','e2e','roofline','cpu_baseline','parity_vs_cpu') if k in s})
PY
echo ==== VARIANTS
for v in _B640x2 _B768x2; do echo "== $v"; PTP_B200_LIB=$PWD/gproshan_b200/libptp_b200$v.so python tools/exp_team.py 447 296 1 1 2>&1 | tail -1; done
echo "== PAIR2 (C3 single)"; python tools/run_single.py 1000 3 2>&1 | tail -1; PTP_B200_LIB=$PWD/gproshan_b200/libptp_b200_PAIR2.so python tools/run_single.py 1000 3 2>&1 | tail -1
PTP_B200_LIB=$PWD/gproshan_b200/libptp_b200_PAIR2.so python tests/variant_check.py 2>&1 | tail -1
echo ==== NCU batched
timeout 600 ncu --set full --clock-control none -k regex:k_batched -s 1 -c 1 -f -o /tmp/r2_k_batched_f200 python tools/run_batched.py 200 148 2 2>&1 | tail -2
python tools/ncu_summary.py /tmp/r2_k_batched_f200.ncu-rep 40 > gpurun_out/r2_ncu_k_batched_f32_f200.txt 2>&1
python tools/ncu_lines.py /tmp/r2_k_batched_f200.ncu-rep 60 >> gpurun_out/r2_ncu_k_batched_f32_f200.txt 2>&1
ls -la /tmp/r2_k_batched_f200.ncu-rep
echo ==== NCU range single
PTP_PROFILE_RANGE=1 timeout 600 ncu --replay-mode app-range --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_fp64.sum -f -o /tmp/r2_c3_range python tools/run_single.py 1000 2 2>&1 | tail -3
ncu -i /tmp/r2_c3_range.ncu-rep --page raw --csv > gpurun_out/r2_c3_range_raw.csv 2>&1; ls -la /tmp/r2_c3_range.ncu-rep; head -c 3000 gpurun_out/r2_c3_range_raw.csv
