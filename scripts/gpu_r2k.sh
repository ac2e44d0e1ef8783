#!/bin/bash
for v in 0 1 2 3; do
echo "=== variant $v"
OBVHS_REINS_VARIANT=$v OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py kitchen 2>&1 | grep -E "round  [0-2] |round 15|reinsertion_optimize" | tail -5
OBVHS_REINS_VARIANT=$v OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py terrain 2>&1 | grep -E "round  [0-2] |round 15|reinsertion_optimize" | tail -5
done
