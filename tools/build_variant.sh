#!/bin/bash
# Kernel-tuning helper: build a variant of the CUDA library in which ONLY the per-size kernel objects named in
# SIZES (default 1024) are recompiled with extra flags; every other object is taken from the default build.
#   tools/build_variant.sh <name> "<nvcc -D flags>" [sizes...]   ->  torchfsm_b200/libfsm_<name>.so
# Used with FSM_B200_LIB=<that .so> by tools/bench_c3.py; never by the product.
set -e
name=$1; flags=$2; shift 2
sizes=${@:-1024}
root=$(cd "$(dirname "$0")/.." && pwd)
cs=$root/torchfsm_b200/csrc
bd=$cs/_build_$name
mkdir -p $bd
objs=""
for o in $cs/_build/*.o; do
  b=$(basename $o)
  keep=1
  for n in $sizes; do [ "$b" = "fsm_kernels_$n.o" ] && keep=0; done
  [ $keep = 1 ] && objs="$objs $o"
done
for n in $sizes; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$root/include -I$cs \
       $flags -DFSM_N=$n -c $cs/fsm_kernels.cu -o $bd/fsm_kernels_$n.o &
done
wait
for n in $sizes; do objs="$objs $bd/fsm_kernels_$n.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $root/torchfsm_b200/libfsm_$name.so $objs
echo built $root/torchfsm_b200/libfsm_$name.so
