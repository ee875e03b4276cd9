export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi -L
python -m pytest tests/test_multi_gpu_native.py -m gpu -q 2>&1 | tail -8
echo ==== native multi device rows; python tools/run_multi.py 447 1024 2 device 2>&1 | tail -1
echo ==== native multi host rows; python tools/run_multi.py 447 1024 1 host 2>&1 | tail -1
echo ==== torchrun N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_n2_a.json'))
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','e2e','gpu_launches','config') if k in d})
except Exception as e:
    print("bench n2 failed", e); print(open('gpurun_out/r2_bench_n2_a.err').read()[-3000:])
PY
echo ==== torchrun ref arm N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['cpu_baseline']['sample'])"
echo ==== ROLLED; PTP_B200_LIB=$PWD/gproshan_b200/libptp_b200_ROLLED.so python tools/exp_team.py 447 296 1 1 2>&1 | tail -1
