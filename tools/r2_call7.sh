#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c7_gpu_tests.log 2>&1
tail -12 gpurun_out/r2c7_gpu_tests.log
timeout 900 python examples/scordelis_lo.py 64 1.0 > gpurun_out/r2c7_shell_64.log 2>&1
tail -4 gpurun_out/r2c7_shell_64.log
timeout 1500 python examples/scordelis_lo.py 256 1.0 > gpurun_out/r2c7_shell_256.log 2>&1
tail -4 gpurun_out/r2c7_shell_256.log
timeout 600 python bench.py --workload annulus --nel 128 --steps 2 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c7_annulus_128.json 2> gpurun_out/r2c7_annulus_128.err
tail -3 gpurun_out/r2c7_annulus_128.err
python - <<'P'
import json
for f in ['r2c7_annulus_128']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
    print(d['parity'])
P
