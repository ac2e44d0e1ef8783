#!/bin/bash
# round 2, call C: ncu --set full of the persistent kernel on the soup, baseline (variant 0) and fused-turn (variant 10): pipe utilisation
mkdir -p gpurun_out
for v in 0 10; do
  VARIANTS=$v:16 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:traverse_persistent_kernel -c 1 -f \
     -o gpurun_out/prof_r2c_soup_v$v python scripts/trav_sweep.py soup > gpurun_out/ncu_r2c_v$v.log 2>&1
  tail -2 gpurun_out/ncu_r2c_v$v.log
done
ls -la gpurun_out/*.ncu-rep
