#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c4_gpu_tests.log 2>&1
tail -8 gpurun_out/r2c4_gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-ptap --no-cpu > gpurun_out/r2c4_bench_256.json 2> gpurun_out/r2c4_bench_256.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c4_bench_256.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
P
timeout 600 python tools/host_profile.py 256 > gpurun_out/r2c4_host_profile.txt 2>&1
head -60 gpurun_out/r2c4_host_profile.txt
