#!/bin/bash
# full GPU test suite + smoke + bench line (device-timed, e2e, roofline, cpu baseline) + sort probe
tag=${1:-r01}
mkdir -p gpurun_out
timeout ${PYTEST_LIMIT:-420} python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${tag}_clocks.csv 2>/dev/null &
smi=$!
timeout 300 python bench.py --steps 30 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
kill $smi 2>/dev/null
timeout 120 python scripts/sort_probe.py 1000000,1800000,20000000,64000000 > gpurun_out/${tag}_sort_probe.log 2>&1
timeout 120 python scripts/stage_times.py > gpurun_out/${tag}_stage_times_cfg2.log 2>&1
tail -12 gpurun_out/${tag}_pytest.log; tail -1 gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_sort_probe.log gpurun_out/${tag}_stage_times_cfg2.log; cut -c1-600 gpurun_out/${tag}_bench.json
