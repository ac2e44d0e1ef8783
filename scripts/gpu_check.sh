#!/bin/bash
# Runs on the B200 box under gpurun: GPU parity tests, smoke, bench. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout -s KILL ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q --tb=short --maxfail=${MAXFAIL:-30} -p no:cacheprovider --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -s KILL 600 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log
if [ "${PROFILE:-0}" = "1" ]; then
  # launch list of one bench run (cold-cache, serialised: compare SHARES) and one full capture of the top kernel
  timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-4000} --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 3 -c 1 -f -o gpurun_out/prof_traverse \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
