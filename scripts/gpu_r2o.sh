#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q --tb=short -x -p no:cacheprovider --timeout 300 -k "sort or morton or ploc or reinsertion or full_size or end_to_end" > gpurun_out/pytest_r2o.log 2>&1; tail -3 gpurun_out/pytest_r2o.log
for v in main s12x4 s8x5 s8x4 s20x2; do
  if [ $v != main ]; then export OBVHS_LIB_PATH=$PWD/obvhs_b200/lib_variants/$v/libobvhs_cuda.so; else unset OBVHS_LIB_PATH; fi
  timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/kt_$v.csv python scripts/trace_build.py terrain 10008338 > /dev/null 2>&1
  echo "== $v"; python scripts/kernel_times.py gpurun_out/kt_$v.csv | head -8
done
