#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 -p no:cacheprovider --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
for wl in soup terrain; do
  timeout -s KILL 900 python bench.py --workload $wl --tris 10000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}.log 2>&1
  tail -1 gpurun_out/bench_${wl}.log
done
