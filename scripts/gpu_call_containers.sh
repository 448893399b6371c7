#!/bin/bash
# new container / io GPU tests + onesweep phase timelines (trace build) for the next tuning round
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_containers_gpu.py tests/test_containers.py tests/test_io_graph.py tests/test_ingest.py tests/test_lift_gpu.py -x -q > gpurun_out/r01p_pytest_containers.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r01p_pytest_containers.log
for n in 1800000 20000000 64000000; do
  PATHPYG_B200_LIB=$PWD/pathpyg_b200/_C/libpathpyg_b200_trace.so timeout 120 python scripts/sort_trace.py $n >> gpurun_out/r01p_sort_trace.log 2>&1
done
tail -5 gpurun_out/r01p_pytest_containers.log; cat gpurun_out/r01p_sort_trace.log
