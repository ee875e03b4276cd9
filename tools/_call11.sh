export PATH=/usr/local/cuda/bin:$PATH
echo "== C3 default"; python tools/run_single.py 1000 4 2>&1 | tail -1
echo "== C3 DYN8"; PTP_B200_LIB=$PWD/gproshan_b200/libptp_b200_DYN8.so python tools/run_single.py 1000 4 2>&1 | tail -1
PTP_B200_LIB=$PWD/gproshan_b200/libptp_b200_DYN8.so python tests/variant_check.py 2>&1 | tail -1
echo "== C2/C4-like default vs DYN8 (f32 1M icosphere)"; python tools/run_single.py 316 3 f32 2>&1 | tail -1; PTP_B200_LIB=$PWD/gproshan_b200/libptp_b200_DYN8.so python tools/run_single.py 316 3 f32 2>&1 | tail -1
echo ==== smoke; python __graft_entry__.py smoke 2>&1 | tail -2
echo ==== tests; python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo ==== launch list under ncu
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu > /tmp/b.log 2>&1; tail -c 300 /tmp/b.log; grep -c . gpurun_out/r2_launches_bench.csv
