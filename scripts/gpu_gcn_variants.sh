#!/bin/bash
# Staged kernel of the fused tcgen05 GCN layer: parity tests, A/B against the other two kernels, phase trace.
# usage (under gpurun): bash scripts/gpu_gcn_ring.sh <tag>
tag=${1:-r02ab}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dbgnn_gpu.py -m gpu -x -q -k "tensor_core or fused_layers or dense_features or replay" > gpurun_out/${tag}_pytest_gcn.log 2>&1
tail -5 gpurun_out/${tag}_pytest_gcn.log
for v in staged single ws; do
  PPG_GCN_TC=$v timeout 120 python scripts/gcn_layer_ab.py --reps 20 --save /tmp/gcn_$v.pt $([ $v != staged ] && echo --compare /tmp/gcn_staged.pt) 2>&1 | tail -2
done > gpurun_out/${tag}_gcn_ab.log
cat gpurun_out/${tag}_gcn_ab.log
PATHPYG_B200_LIB=pathpyg_b200/_C/libpathpyg_b200_trace.so timeout 120 python scripts/gcn_trace.py > gpurun_out/${tag}_gcn_trace.log 2>&1
cat gpurun_out/${tag}_gcn_trace.log
