#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; tail -5 gpurun_out/pytest.log
export VARIANTS=${VARIANTS:-static,persistent:8:32,auto}
for w in kitchen soup bounce; do
  timeout -s KILL 900 python scripts/trav_sweep.py $w > gpurun_out/sweep_$w.log 2>&1; tail -4 gpurun_out/sweep_$w.log
done
for w in kitchen soup terrain; do
timeout -s KILL 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1
tail -1 gpurun_out/bench_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['build'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
done
timeout -s KILL 900 python bench.py --workload bounce --steps 5 --warmup 3 --samples 24 > gpurun_out/bench_bounce.log 2>&1; tail -1 gpurun_out/bench_bounce.log
