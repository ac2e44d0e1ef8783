#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider --timeout 120 > gpurun_out/pytest.log 2>&1; tail -8 gpurun_out/pytest.log
