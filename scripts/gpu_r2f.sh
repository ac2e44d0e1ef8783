#!/bin/bash
# round 2, call F: the new headline bench, small then full size, both arms
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --tris 200000 --rays 3000000 --steps 3 --warmup 3 > gpurun_out/bench_r2f_small.log 2>&1; echo "small rc=$?"; tail -c 1500 gpurun_out/bench_r2f_small.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2f_full.log 2>&1; echo "full rc=$?"; tail -c 6000 gpurun_out/bench_r2f_full.log
timeout -s KILL 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_r2f_ref.log 2>&1; echo "ref rc=$?"; tail -c 2500 gpurun_out/bench_r2f_ref.log
