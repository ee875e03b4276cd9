export PATH=/usr/local/cuda/bin:$PATH
echo "== C3 default"; python tools/run_single.py 1000 4 2>&1 | tail -1
echo "== C3 staged"; PTP_STAGE=1 python tools/run_single.py 1000 4 2>&1 | tail -1
PTP_STAGE=1 python tests/variant_check.py 2>&1 | tail -1
echo ==== NCU batched
timeout 600 ncu --set full --clock-control none -k regex:k_batched -s 1 -c 1 -f -o /tmp/r2_k_batched_f200 python tools/run_batched.py 200 148 2 2>&1 | tail -1
python tools/ncu_summary.py /tmp/r2_k_batched_f200.ncu-rep 30 > gpurun_out/r2_ncu_k_batched_f32_f200.txt 2>&1
python tools/ncu_lines.py /tmp/r2_k_batched_f200.ncu-rep 50 >> gpurun_out/r2_ncu_k_batched_f32_f200.txt 2>&1
ncu -i /tmp/r2_k_batched_f200.ncu-rep --page raw --csv > gpurun_out/r2_k_batched_f200_raw.csv 2>&1
echo ==== NCU app-range C5 1024 sources
PTP_PROFILE_RANGE=1 timeout 900 ncu --replay-mode app-range --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed -f -o /tmp/r2_c5_range python tools/run_batched.py 447 1024 1 2>&1 | tail -3
ncu -i /tmp/r2_c5_range.ncu-rep --page raw --csv > gpurun_out/r2_c5_range_raw.csv 2>&1; ls -la /tmp/*.ncu-rep
