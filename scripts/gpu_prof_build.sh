#!/bin/bash
# DRAM traffic and duration of every kernel of one 10M-triangle build (north_star: "achieved HBM GB/s for build sweeps")
mkdir -p gpurun_out
for wl in ${WORKLOADS:-soup}; do
  timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
     --log-file gpurun_out/build_kernels_$wl.csv python scripts/trace_build.py $wl ${TRIS:-10000000} ${PRESET:-} > gpurun_out/build_kernels_$wl.log 2>&1
  echo "exit $?"; tail -3 gpurun_out/build_kernels_$wl.log
done
ls -la gpurun_out | head -20
