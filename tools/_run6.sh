python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -4
for v in "" _WD0_PF0 _WD1_PF0 _WD1_PF2 _C16 _C32; do echo "==== variant ${v:-default(WD1,PF1,C8)}"; PTP_B200_LIB=$PWD/gproshan_b200/libptp_b200$v.so python tools/exp_team.py 447 296 1 1 2>&1 | tail -1; done
