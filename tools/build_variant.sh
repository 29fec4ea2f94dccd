#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc -D flags...>: tuning build of the rollout TU into variants/libmtgp_<name>.so
set -e
cd /root/repo/multitreegp_b200/csrc
name=$1; shift
mkdir -p /root/repo/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -Xptxas -v -c mtgp_rollout.cu -o /tmp/roll_$name.o 2>&1 | grep -A2 "k_rolloutILi2ELi1ELi1" | grep -E "Used|spill" | tr '\n' ' '
echo " <- $name"
nvcc -shared -o /root/repo/variants/libmtgp_$name.so /tmp/roll_$name.o ../build/mtgp_api.o ../build/mtgp_encode.o ../build/mtgp_pointwise.o -gencode arch=compute_100a,code=sm_100a -lcudart
