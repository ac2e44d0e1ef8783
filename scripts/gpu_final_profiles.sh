#!/bin/bash
# end-of-round evidence: full GPU test suite, smoke, bench lines, launch lists, full ncu captures, per-kernel DRAM traffic
mkdir -p gpurun_out
PROFILE=1 bash scripts/gpu_check.sh
for wl in soup terrain; do
  timeout -s KILL 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}.log 2>&1
  tail -1 gpurun_out/bench_${wl}.log | cut -c1-300
done
timeout -s KILL 900 python bench.py --workload bounce --samples 24 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bounce.log 2>&1
timeout -s KILL 900 python bench.py --workload dynamic --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dynamic.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_soup.csv \
   python bench.py --workload soup --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_soup_ncu.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:traverse_persistent_kernel -s 3 -c 1 -f -o gpurun_out/prof_traverse_soup \
   python bench.py --workload soup --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_soup.log 2>&1
WORKLOADS="soup kitchen" bash scripts/gpu_prof_build.sh
