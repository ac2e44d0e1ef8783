#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/dbg_miss2.py 20 2>&1 | grep -v "closest\|miss\|count" | tail -5
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -p no:cacheprovider --timeout 120 > gpurun_out/pytest.log 2>&1; tail -8 gpurun_out/pytest.log
export VARIANTS=${VARIANTS:-static,persistent:8:32,auto}
for w in kitchen soup bounce; do
  timeout -s KILL 900 python scripts/trav_sweep.py $w > gpurun_out/sweep_$w.log 2>&1; tail -4 gpurun_out/sweep_$w.log
done
