#!/bin/bash
# round 2, call A: regression of the refactored traversal + persistent-kernel variant sweep on the three 10M workloads
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=10 -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_r2a.log 2>&1
tail -5 gpurun_out/pytest_r2a.log
for wl in soup "bounce 10000000 16" terrain; do
  timeout 900 python scripts/trav_sweep.py $wl 2>&1 | tee -a gpurun_out/sweep_r2a.log | tail -14
done
PACKED=1 VARIANTS=0:16,3:16 timeout 600 python scripts/trav_sweep.py soup 2>&1 | tee -a gpurun_out/sweep_r2a.log | tail -3
