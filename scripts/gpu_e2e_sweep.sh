#!/bin/bash
# host-batch (e2e) pipeline: regression tests + slice-size sweep on the large-scene workloads
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -p no:cacheprovider --timeout 300 -k "ray_new or pipelined or zero_copy or two_compute" 2>&1 | tail -5
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print(d["config"]["workload"][:30], round(d["value"]), "e2e", round(e["value"]), "struct", round(e["ray_struct"]["value"]), "build", round(e["build_mtris_per_s"]))'
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
for slice in ${SLICES:-0}; do
  for w in "--workload soup" "--workload terrain" "--workload bounce --samples 24"; do
    echo -n "slice $slice: "
    OBVHS_BENCH_HOST_SLICE=$slice timeout 600 python bench.py $w --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
  done
done
