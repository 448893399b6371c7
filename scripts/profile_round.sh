#!/bin/bash
# One round of GPU evidence: tests, bench, launch list, full ncu captures of the dominant kernels.
# usage (under gpurun): bash scripts/profile_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${tag}_clocks.csv 2>/dev/null &
smi=$!
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
kill $smi 2>/dev/null
timeout 120 python scripts/stage_times.py > gpurun_out/${tag}_stage_times_cfg2.log 2>&1
timeout 300 python scripts/stage_times.py --m 10000000 --T 250 --delta 5 --order 3 > gpurun_out/${tag}_stage_times_cfg3.log 2>&1
timeout 300 python scripts/sort_probe.py 1000000,1800000,20000000,64000000 > gpurun_out/${tag}_sort_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --cpu-budget 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass_kernel" -s 10 -c 5 -o gpurun_out/${tag}_sort1p8m \
    python scripts/sort_probe.py 1800000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass_kernel" -s 10 -c 5 -o gpurun_out/${tag}_sort64m \
    python scripts/sort_probe.py 64000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gcn_tc_kernel|gcn_fused_kernel" -s 10 -c 5 -o gpurun_out/${tag}_gcn \
    python scripts/stage_times.py > /dev/null 2>&1
tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_smoke.log | tail -1; cat gpurun_out/${tag}_sort_probe.log gpurun_out/${tag}_stage_times_cfg2.log gpurun_out/${tag}_stage_times_cfg3.log
