#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -q --tb=short -x -p no:cacheprovider --timeout 300 -k "traversal or kitchen or golden or variants or random" > gpurun_out/pytest_r2r.log 2>&1; tail -3 gpurun_out/pytest_r2r.log
for v in main noimad main noimad; do
  if [ $v != main ]; then export OBVHS_LIB_PATH=$PWD/obvhs_b200/lib_variants/$v/libobvhs_cuda.so; else unset OBVHS_LIB_PATH; fi
  echo "== $v"
  for wl in soup "bounce 10000000 16" terrain kitchen; do
    VARIANTS=0 timeout 600 python scripts/trav_sweep.py $wl 2>&1 | tail -1
  done
done
