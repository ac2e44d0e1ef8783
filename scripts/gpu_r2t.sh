#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider --timeout 300 -k "reinsertion or build or parity or deep or dynamic" > gpurun_out/pytest_r2t.log 2>&1
tail -3 gpurun_out/pytest_r2t.log
timeout 600 python scripts/trace_build.py terrain 10000000 > gpurun_out/trace_t10m.out 2> gpurun_out/trace_t10m.err
timeout 300 python scripts/trace_build.py kitchen > gpurun_out/trace_kitchen.out 2> gpurun_out/trace_kitchen.err
awk '/--- build 2/,0' gpurun_out/trace_t10m.err | grep -E "round  ?(0|1|2|8|15) |reinsertion_optimize|total"
awk '/--- build 2/,0' gpurun_out/trace_kitchen.err | grep -E "round  ?(0|1|2|8|15) |reinsertion_optimize|total"
timeout 400 python bench.py --workload dynamic --steps 20 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dynamic', l['value'], l['ms_per_step'])"
