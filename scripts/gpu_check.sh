#!/bin/bash
# Runs on the B200 box under gpurun: GPU parity tests, smoke, bench. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout -s KILL ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q --tb=short --maxfail=${MAXFAIL:-30} -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -s KILL 600 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log
