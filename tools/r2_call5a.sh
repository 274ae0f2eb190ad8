#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_solvers.py -q -k "config2" > gpurun_out/r2c5_cfg3.log 2>&1
tail -12 gpurun_out/r2c5_cfg3.log
timeout 900 python bench.py --no-ptap --no-cpu > gpurun_out/r2c5_bench_256.json 2> gpurun_out/r2c5_bench_256.err
tail -3 gpurun_out/r2c5_bench_256.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c5_bench_256.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e'], d['config']['cg_iterations'], d['gpu_launches'])
print(d['parity']); print(d.get('ptap')); print(d.get('cpu_baseline')); print(d['fp64_peak_tflops_measured'])
for r in d['rooflines']: print("%-50s %7.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
