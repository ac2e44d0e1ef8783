#!/bin/bash
# compute-sanitizer (memcheck, then racecheck / synccheck on a smaller selection) over parity tests on small scenes
mkdir -p gpurun_out
SEL=${SEL:-"(cornell or terrain32 or flat4 or ico_plane or splitty or degenerate or varying) and not large and not full_size and not 10m"}
timeout -s KILL ${T1:-900} compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 0 \
   python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider --timeout 800 -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | tail -8
SEL2=${SEL2:-"(cornell or terrain32) and (ploc or sort or reinsertion or cwbvh or rebuild or end_to_end) and not large"}
timeout -s KILL ${T2:-600} compute-sanitizer --tool racecheck --error-exitcode 7 --launch-timeout 0 \
   python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider --timeout 800 -k "$SEL2" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | tail -8
