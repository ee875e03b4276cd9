#!/bin/bash
# A/B runs on the GPU box: tools/ab_run.sh "<variant names>" [f=447] [n_sources=296] [reps=3] [dtype]
# prints, per variant of gproshan_b200/_ab/, the sources/s of every repetition and the matrix checksum
cd "$(dirname "$0")/.."
for v in $1; do
  lib=gproshan_b200/_ab/libptp_b200_$v.so
  [ "$v" = "default" ] && lib=gproshan_b200/libptp_b200.so
  echo "== $v"
  PTP_B200_LIB=$PWD/$lib timeout 300 python tools/run_batched.py ${2:-447} ${3:-296} ${4:-3} ${5:-f32} 2>&1 | grep -o "sources/s [0-9.]*\|checksum.*\|Error.*\|error.*" | tr '\n' ' '
  echo
done
