#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -p no:cacheprovider --timeout 150 -k "ploc_bvh2_bit_exact or deep_trees" > gpurun_out/pytest_r2l_first.log 2>&1 || { tail -40 gpurun_out/pytest_r2l_first.log; echo "first test failed / hung: stopping"; exit 0; }
tail -2 gpurun_out/pytest_r2l_first.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_r2l.log 2>&1
tail -4 gpurun_out/pytest_r2l.log
OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py terrain 2>&1 | grep -E "^\[obvhs trace\]   ploc|^\[obvhs trace\] |total" | tail -12
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2l.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_r2l.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('s3 value', d['value'], 'build', d['build']['ms'], 'kitchen build', d['kitchen']['build']['ms'], 'kitchen value', d['kitchen']['value'], 'launches', d['gpu_launches'])
else: print(open('gpurun_out/bench_r2l.log').read()[-2000:])
PY
