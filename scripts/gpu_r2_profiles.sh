#!/bin/bash
# Round-2 evidence (one GPU): full GPU test suite, smoke, the headline bench line (never under a profiler), then under ncu:
# launch list of one headline step, full captures of the S3 and kitchen traversal kernels, per-kernel DRAM traffic of the 10M build.
TAG=${TAG:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1; nproc >> gpurun_out/smi_$TAG.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=20 -p no:cacheprovider --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_$TAG.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_$TAG.log
timeout -s KILL 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log; tail -c 600 gpurun_out/bench_$TAG.log
# launch list: one timed step of the headline (S3 step = build + traversal; kitchen likewise); ray generation launches come first
timeout -s KILL 1200 ncu --nvtx --nvtx-include "obvhs_timed_step/" --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# full capture of the S3 traversal launch: skip the 272 primary-ray launches of the ray generation and the 3 warm-up steps
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:traverse_persistent_kernel -s 275 -c 1 -f -o gpurun_out/prof_s3_traverse_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_s3_$TAG.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"traverse_kernel" -s 3 -c 1 -f -o gpurun_out/prof_kitchen_traverse_$TAG \
    python bench.py --workload kitchen --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kitchen_$TAG.log 2>&1
# the other BASELINE configs and the layout passes (never under a profiler)
timeout -s KILL 600 python bench.py --workload passes --steps 3 > gpurun_out/bench_passes_$TAG.log 2>&1
timeout -s KILL 600 python bench.py --workload cornell --steps 10 > gpurun_out/bench_cornell_$TAG.log 2>&1
timeout -s KILL 900 python bench.py --workload dynamic --steps 10 > gpurun_out/bench_dynamic_$TAG.log 2>&1
# full captures of the two conversion kernels of the 10 M build (third build of trace_build.py)
for k in cwbvh_cost_frontier_kernel cwbvh_emit_all_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_$TAG \
     python scripts/trace_build.py terrain 10008338 > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
for wl in terrain kitchen; do
  timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
     --log-file gpurun_out/build_kernels_${wl}_$TAG.csv python scripts/trace_build.py $wl 10008338 > gpurun_out/build_kernels_${wl}_$TAG.log 2>&1
done
ls -la gpurun_out | grep $TAG
