#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cpp_host.py -m gpu -q --tb=short -x -p no:cacheprovider --timeout 300 -k "origin_direction or ray_new or cpp_host or two_compute or persistent_kernel_variants" > gpurun_out/pytest_r2u.log 2>&1
tail -5 gpurun_out/pytest_r2u.log
timeout 900 python bench.py --steps 5 > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err
tail -3 gpurun_out/bench_r2u.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2u.json').read().strip().splitlines()[-1])
print('value',l['value'],'build',l['build']['ms'],'e2e',l['e2e']['value'],'ray_new',l['e2e']['ray_new']['value'],'struct',l['e2e']['ray_struct']['value'],'link',l['e2e']['host_link'])
k=l['kitchen']
print('kitchen',k['value'],'build',k['build']['ms'],'e2e',k['e2e']['value'],'ray_new',k['e2e']['ray_new']['value'],'parity',l['parity_ok'])
PY
