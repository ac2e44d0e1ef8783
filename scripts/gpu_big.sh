#!/bin/bash
# large-scene bench (10M triangle soup / terrain) + launch list; results in gpurun_out/
mkdir -p gpurun_out
for wl in soup terrain; do
  timeout -s KILL 900 python bench.py --workload $wl --tris ${TRIS:-10000000} --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}.log 2>&1
  echo "exit $?" >> gpurun_out/bench_${wl}.log
  tail -2 gpurun_out/bench_${wl}.log
done
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_soup.csv \
   python bench.py --workload soup --tris ${TRIS:-10000000} --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_soup_ncu.log 2>&1
