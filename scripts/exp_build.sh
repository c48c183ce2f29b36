#!/bin/bash
# build kernel variants for scripts/exp_ab.py: exp_build.sh name "-DFLAG=1 ..." [name2 "flags" ...]   -> exp/lib_<name>.so
mkdir -p exp
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I include -Xptxas -v $f \
      -o exp/lib_$n.so evosops_b200/csrc/sops_host.cu > exp/build_$n.log 2>&1; echo "$n rc=$?" ) &
done
wait
