#!/bin/bash
# round 2, call G (2 GPUs): the 2-GPU broadcast test and the headline bench at N = 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_r2g.txt 2>&1
timeout -s KILL 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider --timeout 500 > gpurun_out/pytest_r2g.log 2>&1; tail -5 gpurun_out/pytest_r2g.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2g_n2.log 2>&1; echo "n2 rc=$?"; tail -c 3000 gpurun_out/bench_r2g_n2.log
