#!/bin/bash
# full ncu capture of the persistent traversal kernel on the 10M-triangle soup (incoherent rays) + terrain bench line
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:traverse_persistent_kernel -s 3 -c 1 -f -o gpurun_out/prof_traverse_soup \
   python bench.py --workload soup --tris ${TRIS:-10000000} --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_soup.log 2>&1
timeout -s KILL 900 python bench.py --workload terrain --tris ${TRIS:-10000000} --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_terrain.log 2>&1
tail -1 gpurun_out/bench_terrain.log | cut -c1-900
timeout -s KILL 900 python bench.py --workload bounce --samples 24 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bounce.log 2>&1
tail -1 gpurun_out/bench_bounce.log | cut -c1-1500
ls -la gpurun_out
