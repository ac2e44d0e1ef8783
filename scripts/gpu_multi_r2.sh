#!/bin/bash
# headline bench on N GPUs of one box (torchrun), plus the 2-GPU broadcast test
N=${N:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout -s KILL 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider --timeout 500 > gpurun_out/pytest_multi_n$N.log 2>&1; tail -2 gpurun_out/pytest_multi_n$N.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/bench_n$N.log | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'bcast', d['broadcast'], 'parity_ok', d['parity_ok'], 'kitchen', round(d['kitchen']['value']), round(d['kitchen']['e2e']['value']), d['kitchen']['broadcast'], 'link', d['e2e']['host_link'])
PY
