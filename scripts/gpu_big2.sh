#!/bin/bash
# 10M-triangle scenes: bench lines, stage traces, launch list, full ncu capture of the traversal kernel on incoherent rays
mkdir -p gpurun_out
for wl in soup terrain; do
  timeout -s KILL 900 python bench.py --workload $wl --tris ${TRIS:-10000000} --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}.log 2>&1
  echo "exit $?" >> gpurun_out/bench_${wl}.log
  tail -2 gpurun_out/bench_${wl}.log
  python scripts/trace_build.py $wl ${TRIS:-10000000} > gpurun_out/trace_$wl.log 2>&1
  grep -v "reins_" gpurun_out/trace_$wl.log | tail -11
done
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_soup.csv \
   python bench.py --workload soup --tris ${TRIS:-10000000} --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_soup_ncu.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 3 -c 1 -f -o gpurun_out/prof_traverse_soup \
   python bench.py --workload soup --tris ${TRIS:-10000000} --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_soup.log 2>&1
ls -la gpurun_out
