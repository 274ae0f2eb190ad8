#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c9_gpu_tests.log 2>&1
tail -6 gpurun_out/r2c9_gpu_tests.log
timeout 300 python tools/host_profile_extract.py 256 > gpurun_out/r2c9_host_extract.txt 2>&1
head -90 gpurun_out/r2c9_host_extract.txt
timeout 900 python bench.py > gpurun_out/r2c9_bench_256_full.json 2> gpurun_out/r2c9_bench_256_full.err
tail -3 gpurun_out/r2c9_bench_256_full.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c9_bench_256_full.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
print(d.get('ptap')); print(d.get('cpu_baseline'))
P
