#!/bin/bash
# ncu evidence for the sort kernel after a change: launch list of the bench step + --set full captures at two sizes
tag=${1:-r01}
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --cpu-budget 1 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass_kernel" -s 10 -c 3 -o gpurun_out/${tag}_sort1p8m \
    python scripts/sort_probe.py 1800000 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass_kernel" -s 10 -c 3 -o gpurun_out/${tag}_sort64m \
    python scripts/sort_probe.py 64000000 > /dev/null 2>&1
ls -la gpurun_out
