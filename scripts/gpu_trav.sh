#!/bin/bash
# traversal regression + the three traversal workloads
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 -p no:cacheprovider --timeout 300 -k "traversal or kitchen or icosphere or golden or pipelined or zero_copy" > gpurun_out/pytest_trav.log 2>&1
tail -3 gpurun_out/pytest_trav.log
for args in "--workload soup" "--workload bounce --samples 24" "--workload terrain" ""; do
  timeout 600 python bench.py $args --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:40], 'Mrays/s', round(d['value']), 'trav_ms', round(d['traverse_ms'],3), 'hits', d['hits'], 'frac', round(d['roofline']['frac'],3))"
done
