#!/bin/bash
# correctness of the sort / PLOC paths, then 10M build traces + bench lines
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 -p no:cacheprovider --timeout 300 -k "${KEXPR:-sort or morton or ploc or large or end_to_end or rebuild}" > gpurun_out/pytest_perf.log 2>&1
tail -3 gpurun_out/pytest_perf.log
for w in ${WORKLOADS:-soup}; do
  python scripts/trace_build.py $w ${TRIS:-10000000} > gpurun_out/trace_$w.log 2>&1
  awk '/--- build 2/{f=1} f' gpurun_out/trace_$w.log | grep -v "reins_\|emit level" | tail -12
  awk '/--- build 2/{f=1} f' gpurun_out/trace_$w.log | grep "reins_" | awk '{a[$3]+=$4; c[$3]+=$6} END {for (k in a) print k, a[k], "ms", c[k], "launches"}'
  timeout -s KILL 600 python bench.py --workload $w --tris ${TRIS:-10000000} --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1
  tail -1 gpurun_out/bench_$w.log | cut -c1-900
done
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kitchen.log 2>&1
tail -1 gpurun_out/bench_kitchen.log | cut -c1-900
