#!/bin/bash
# round 2, call D: POLICY 3 (fetched nodes held until the warp's node test is worth running) -- parity + sweep
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -p no:cacheprovider --timeout 600 -x -k "persistent_kernel_variants" > gpurun_out/pytest_r2d.log 2>&1
tail -3 gpurun_out/pytest_r2d.log
export VARIANTS=0:16,10,13:8,13:12,13:16,13:20,13:24,13:28,14:16,14:24,15:16,13:16/8,13:20/2
for wl in soup "bounce 10000000 16" terrain; do
  timeout 900 python scripts/trav_sweep.py $wl 2>&1 | tee -a gpurun_out/sweep_r2d.log | tail -14
done
