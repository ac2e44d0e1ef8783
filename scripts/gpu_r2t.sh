#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider --timeout 300 -k "reinsertion or build or parity or deep or dynamic or random or candidates" > gpurun_out/pytest_r2t.log 2>&1
tail -2 gpurun_out/pytest_r2t.log
timeout 300 python scripts/build_time.py kitchen 2>&1 | tail -1
timeout 300 python scripts/build_time.py terrain 10000000 2>&1 | tail -1
timeout 400 python bench.py --workload dynamic --steps 20 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dynamic', l['value'], l['ms_per_step'])"
timeout 400 python scripts/trace_dynamic.py 2>&1 | awk "/--- frame 2/,0" | grep -E "round  ?(0|1|5|10|15) |reinsertion_optimize"
timeout 300 python scripts/trace_build.py kitchen 2>&1 | awk '/--- build 2/,0' | grep -E "round  ?(0|1|2|4|8) |reinsertion_optimize"
