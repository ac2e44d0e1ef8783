#!/bin/bash
# per-phase timings of the single-launch reinsertion run
OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py kitchen 2>&1 | grep -E "round|reinsertion run|total" | tail -19
OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py terrain 2>&1 | grep -E "round|reinsertion run|total" | tail -19
