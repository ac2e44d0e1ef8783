#!/bin/bash
# quick GPU regression: parity tests + stage traces
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 -p no:cacheprovider --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
for w in ${WORKLOADS:-kitchen soup terrain}; do
  python scripts/trace_build.py $w ${TRIS:-10000000} > gpurun_out/trace_$w.log 2>&1
  grep -v "reins_" gpurun_out/trace_$w.log | tail -11
  grep "reins_" gpurun_out/trace_$w.log | tail -64 | awk '{a[$3]+=$4; c[$3]+=$6} END {for (k in a) print k, a[k], "ms", c[k], "launches"}'
done
