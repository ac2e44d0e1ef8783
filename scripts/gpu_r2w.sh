#!/bin/bash
# ncu --set full of the bandwidth-side build kernels at 10 M triangles (third build of trace_build.py: warm arena)
mkdir -p gpurun_out
prof() {  # name regex skip
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_build_$1 \
     python scripts/trace_build.py terrain 10008338 > gpurun_out/ncu_build_$1.log 2>&1
  ncu -i gpurun_out/prof_build_$1.ncu-rep --page raw --csv > gpurun_out/prof_build_$1.csv 2>/dev/null
  ls -la gpurun_out/prof_build_$1.ncu-rep | awk '{print $5, $9}'
}
prof ploc_fused0 ploc_fused_kernel 16
prof ploc_fused3 ploc_fused_kernel 19
prof cost cwbvh_cost_frontier_kernel 2
prof emit cwbvh_emit_all_kernel 2
prof onesweep "onesweep_kernel" 16
