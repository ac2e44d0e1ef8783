#!/bin/bash
# N-GPU runs exactly as the driver launches them (torchrun, one rank per GPU)
mkdir -p gpurun_out
N=${N:-2}
for args in "" "--workload bounce --samples ${SAMPLES:-32}"; do
  tag=$(echo "$args" | tr -d ' -' | cut -c1-20)
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 $args > gpurun_out/bench_n${N}_${tag}.log 2>&1
  echo "exit $?"; tail -1 gpurun_out/bench_n${N}_${tag}.log | cut -c1-2500
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 $args > gpurun_out/ref_n${N}_${tag}.log 2>&1
  echo "ref exit $?"; tail -1 gpurun_out/ref_n${N}_${tag}.log | cut -c1-800
done
