#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --workload cornell --steps 10 --warmup 3 > gpurun_out/bench_cornell.log 2>&1; echo "cornell rc=$?"; tail -c 900 gpurun_out/bench_cornell.log
timeout -s KILL 900 python bench.py --workload dynamic --steps 100 --warmup 3 > gpurun_out/bench_dynamic.log 2>&1; echo "dynamic rc=$?"; tail -c 1500 gpurun_out/bench_dynamic.log
timeout -s KILL 600 python bench.py --workload kitchen --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kitchen_legacy.log 2>&1; echo "kitchen rc=$?"; tail -c 300 gpurun_out/bench_kitchen_legacy.log
