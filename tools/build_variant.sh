#!/bin/bash
# tools/build_variant.sh <name> <tu: rollout|pointwise> <extra nvcc -D flags...>: tuning build of one TU into variants/libmtgp_<name>.so
set -e
cd /root/repo/multitreegp_b200/csrc
name=$1; tu=$2; shift; shift
mkdir -p /root/repo/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --no-compress "$@" -Xptxas -v -c mtgp_$tu.cu -o /tmp/var_$name.o 2>&1 | grep -A2 -E "k_rolloutILi2ELi1ELi1ELi2ELb0|k_rollout_rkILi2ELi1ELb1ELi2|k_pointwiseILi8" | grep -E "Used|spill" | tr '\n' ' '
echo " <- $name"
objs=""
for o in api encode rollout rollout_rk pointwise select; do
  if [ "$o" == "$tu" ]; then objs="$objs /tmp/var_$name.o"; else objs="$objs ../build/mtgp_$o.o"; fi
done
nvcc -shared -o /root/repo/variants/libmtgp_$name.so $objs -gencode arch=compute_100a,code=sm_100a -lcudart
cd /root/repo && python -m multitreegp_b200.sass_patch variants/libmtgp_$name.so rollout pointwise 2>&1 | tail -1
