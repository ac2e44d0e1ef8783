#!/bin/bash
# round 2, call B: fused-turn persistent kernel (POLICY 2) -- parity of the new variants + sweep
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -p no:cacheprovider --timeout 600 -x -k "persistent_kernel_variants or traversal_hits" > gpurun_out/pytest_r2b.log 2>&1
tail -3 gpurun_out/pytest_r2b.log
export VARIANTS=0:16,9,10,11,12,9/8,10/8,10/2,10/1
for wl in soup "bounce 10000000 16" terrain; do
  timeout 900 python scripts/trav_sweep.py $wl 2>&1 | tee -a gpurun_out/sweep_r2b.log | tail -10
done
