#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_bspline.py tests/test_gpu_gsf.py tests/test_gpu_solvers.py -m gpu -q -x > gpurun_out/r2c21_tests.log 2>&1
tail -4 gpurun_out/r2c21_tests.log
for gb in 20 40 80; do
TIGAR_B200_GSF_GB=$gb timeout 600 python bench.py --no-ptap --no-cpu > gpurun_out/r2c21_bench_gb$gb.json 2> gpurun_out/r2c21_bench_gb$gb.err
tail -2 gpurun_out/r2c21_bench_gb$gb.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c21_bench_gb$gb.json').read().strip().splitlines()[-1])
print($gb, d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['gpu_launches'])
for r in d['rooflines'][:5]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
done
