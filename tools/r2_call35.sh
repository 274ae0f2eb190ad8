#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_gsf.py tests/test_zz_gpu_multifield.py -m gpu -q -x > gpurun_out/r2c35_tests.log 2>&1
tail -3 gpurun_out/r2c35_tests.log
timeout 600 python bench.py --no-ptap --no-cpu > gpurun_out/r2c35_bench.json 2> gpurun_out/r2c35_bench.err
tail -2 gpurun_out/r2c35_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c35_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['gpu_launches'], d['parity'])
for r in d['rooflines'][:5]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
