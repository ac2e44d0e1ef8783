#!/bin/bash
OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py terrain 2>&1 | grep -vE "round|emit level|result cudaMalloc" | tail -14
OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py kitchen 2>&1 | grep -vE "round|emit level|result cudaMalloc" | tail -14
