#!/bin/bash
# round 2, call H: reinsertion as one cooperative launch -- parity, trace, bench
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -p no:cacheprovider --timeout 100 -k "reinsertion_bit_exact" > gpurun_out/pytest_r2h_first.log 2>&1 || { tail -30 gpurun_out/pytest_r2h_first.log; echo "first reinsertion test failed / hung: stopping"; exit 0; }
tail -2 gpurun_out/pytest_r2h_first.log
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider --timeout 600 -x -k "reinsert or dynamic or end_to_end or golden or random or bvh2 or full_size or splits" > gpurun_out/pytest_r2h.log 2>&1
tail -8 gpurun_out/pytest_r2h.log
OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py kitchen 2>&1 | grep -E "reinsertion|total|build_ploc|bvh2_to_cwbvh" | tail -8
OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py terrain 2>&1 | grep -E "reinsertion|total|build_ploc|bvh2_to_cwbvh" | tail -8
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2h.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_r2h.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('s3 value', d['value'], 'build', d['build'], 'kitchen build', d['kitchen']['build'], 'kitchen value', d['kitchen']['value'])
else: print(open('gpurun_out/bench_r2h.log').read()[-2000:])
PY
