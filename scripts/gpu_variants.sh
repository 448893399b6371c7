#!/bin/bash
# usage: bash scripts/gpu_variants.sh <tag> <variant> [<variant> ...]   (variant "base" = the default library)
# correctness (tests/test_sort_gpu.py) and per-pass timing (scripts/sort_probe.py) of experiment builds of the library
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/${tag}_variants.log
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then unset PATHPYG_B200_LIB; else export PATHPYG_B200_LIB=$PWD/pathpyg_b200/_C/libpathpyg_b200_$v.so; fi
  echo "=== $v" >> $out
  timeout 150 python -m pytest tests/test_sort_gpu.py -x -q 2>&1 | tail -2 >> $out
  timeout 120 python scripts/sort_probe.py ${SIZES:-1000000,1800000,20000000,64000000} >> $out 2>&1
done
cat $out
