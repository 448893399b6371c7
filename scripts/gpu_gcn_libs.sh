#!/bin/bash
# A/B of experiment builds of the library (make variant NAME=...) on the order-2 GCN layer of cfg2.
# usage (under gpurun): bash scripts/gpu_gcn_libs.sh <tag> <name> [<name> ...]   ("default" = the shipped library)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=pathpyg_b200/_C/libpathpyg_b200_$v.so
  [ "$v" = default ] && lib=pathpyg_b200/_C/libpathpyg_b200.so
  echo "== $v"
  PATHPYG_B200_LIB=$lib timeout 120 python scripts/gcn_layer_ab.py --reps 20 --save /tmp/gcn_$v.pt $([ "$v" != default ] && echo --compare /tmp/gcn_default.pt) 2>&1 | tail -2
done > gpurun_out/${tag}_gcn_libs.log
cat gpurun_out/${tag}_gcn_libs.log
