#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider --timeout 300 -k "cwbvh or build or parity or convert or big or large" > gpurun_out/pytest_r2x.log 2>&1
tail -3 gpurun_out/pytest_r2x.log
for wl in terrain soup; do
  timeout 600 python scripts/trace_build.py $wl 10000000 2>&1 | awk '/--- build 2/,0' | grep -E "calculate_cost|convert_to|total"
done
