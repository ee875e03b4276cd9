#!/bin/bash
# A/B builds of the CUDA library with extra -D switches:  tools/ab_build.sh <name> [-DPTP_X=1 ...]
# -> gproshan_b200/_ab/libptp_b200_<name>.so (git-ignored; travels to the GPU box); select with PTP_B200_LIB=<path>
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gproshan_b200/_ab
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
     -prec-div=true -prec-sqrt=true -ftz=false -Xptxas -v "$@" -o gproshan_b200/_ab/libptp_b200_$name.so gproshan_b200/csrc/ptp_api.cu \
     > gproshan_b200/_ab/$name.ptxas.log 2>&1
echo "built $name"
