#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_bvh2.py tests/test_gpu_parity.py -m gpu -q --tb=short -x -p no:cacheprovider --timeout 200 -k "reorder or order_children" > gpurun_out/pytest_r2s_first.log 2>&1 || { tail -60 gpurun_out/pytest_r2s_first.log; echo "new tests failed / hung: stopping"; exit 0; }
tail -2 gpurun_out/pytest_r2s_first.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider --timeout 600 > gpurun_out/pytest_r2s.log 2>&1
tail -4 gpurun_out/pytest_r2s.log
