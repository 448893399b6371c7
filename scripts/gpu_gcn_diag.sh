#!/bin/bash
# Diagnosis of the fused tcgen05 GCN layer at the order-2 graph of cfg2: per-tile phase time stamps (instrumented
# library), device time of the layer, one ncu --set full capture with source correlation.
# usage (under gpurun): bash scripts/gpu_gcn_diag.sh <tag>
tag=${1:-r02aa}
mkdir -p gpurun_out
PATHPYG_B200_LIB=pathpyg_b200/_C/libpathpyg_b200_trace.so timeout 300 python scripts/gcn_trace.py > gpurun_out/${tag}_gcn_trace.log 2>&1
timeout 300 python scripts/gcn_layer_ab.py --reps 20 > gpurun_out/${tag}_gcn_time.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gcn_tc" -s 3 -c 1 -o gpurun_out/${tag}_gcn \
    python scripts/gcn_layer_ab.py --reps 1 > gpurun_out/${tag}_ncu.log 2>&1
cat gpurun_out/${tag}_gcn_trace.log gpurun_out/${tag}_gcn_time.log; tail -3 gpurun_out/${tag}_ncu.log
