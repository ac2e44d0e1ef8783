#!/bin/bash
# round-end numbers without a profiler: full GPU suite, smoke, and the bench line of every workload
mkdir -p gpurun_out
bash scripts/gpu_check.sh
for wl in soup terrain demoscene; do
  timeout -s KILL 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}.log 2>&1
done
timeout -s KILL 900 python bench.py --workload bounce --samples 24 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bounce.log 2>&1
timeout -s KILL 900 python bench.py --workload dynamic --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dynamic.log 2>&1
for wl in soup terrain demoscene bounce dynamic; do grep '^{' gpurun_out/bench_${wl}.log | tail -1 | cut -c1-200; done
