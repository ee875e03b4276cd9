export PATH=/usr/local/cuda/bin:$PATH
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo ==== BENCH
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; tail -c 1500 gpurun_out/r2_bench_n1_a.json; tail -5 gpurun_out/r2_bench_n1_a.err
echo ==== REF
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2>/dev/null; cat gpurun_out/r2_bench_ref_a.json | head -c 1200
echo ==== NCU batched
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batched -s 1 -c 1 -f -o gpurun_out/r2_k_batched_f200 python tools/run_batched.py 200 148 2 2>&1 | tail -3
echo ==== NCU range single
PTP_PROFILE_RANGE=1 timeout 600 ncu --replay-mode app-range --set full --clock-control none -f -o gpurun_out/r2_c3_range python tools/run_single.py 1000 2 2>&1 | tail -5
