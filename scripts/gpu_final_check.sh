#!/bin/bash
# End-of-round evidence on one B200: full GPU test suite, smoke, the default bench line (with the extra workloads), the
# ncu launch list of the cfg2 step, A/B of the three GCN layer kernels.
# usage (under gpurun): bash scripts/gpu_final_check.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout ${PYTEST_LIMIT:-420} python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_full.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --cpu-budget 0.5 > /dev/null 2>&1
for v in staged single ws; do
  PPG_GCN_TC=$v timeout 120 python scripts/gcn_layer_ab.py --reps 20 --save /tmp/gcn_$v.pt $([ $v != staged ] && echo --compare /tmp/gcn_staged.pt) 2>&1 | tail -2
done > gpurun_out/${tag}_gcn_ab.log
for v in staged single; do PPG_GCN_TC=$v timeout 120 python scripts/gcn_layer_ab.py --layer 1 --reps 20 2>&1 | tail -1; done >> gpurun_out/${tag}_gcn_ab.log
tail -4 gpurun_out/${tag}_pytest_full.log; tail -1 gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_gcn_ab.log
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_1gpu.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_lift", "ms_per_step_dbgnn", "gpu_launches", "gpu_launches_lift")}, d["e2e"]["ms_per_step"], d["clocks"])
print({k: d["roofline"][k] for k in ("kernel", "frac", "traffic", "share_of_step", "launch_ms")})
for k, v in d.get("extra", {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "parity_ok", "error")}, (v.get("roofline") or {}).get("frac"))
PY
