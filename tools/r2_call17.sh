#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c17_gpu_tests.log 2>&1
tail -12 gpurun_out/r2c17_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2c17_bench_256_full.json 2> gpurun_out/r2c17_bench_256_full.err
tail -3 gpurun_out/r2c17_bench_256_full.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c17_bench_256_full.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
print(d.get('ptap')); print(d.get('ptap_fused')); print(d.get('cpu_baseline')); print(d['roofline'])
P
