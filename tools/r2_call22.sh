#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c22_tests.log 2>&1
tail -4 gpurun_out/r2c22_tests.log
for gb in 20 12 8; do
TIGAR_B200_GSF_GB=$gb timeout 600 python bench.py --no-ptap --no-cpu > gpurun_out/r2c22_bench_gb$gb.json 2> gpurun_out/r2c22_bench_gb$gb.err
tail -2 gpurun_out/r2c22_bench_gb$gb.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c22_bench_gb$gb.json').read().strip().splitlines()[-1])
print($gb, d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['gpu_launches'], d['parity'])
for r in d['rooflines'][:5]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
done
