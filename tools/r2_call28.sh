#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_zz_gpu_multifield.py tests/test_gpu_gsf.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/r2c28_tests.log 2>&1
tail -5 gpurun_out/r2c28_tests.log
timeout 900 python bench.py --mode matfree --nel 512 --steps 1 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c28_matfree_512.json 2> gpurun_out/r2c28_matfree_512.err
tail -3 gpurun_out/r2c28_matfree_512.err
python - <<'P'
import json
for f in ['r2c28_matfree_512']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['config']['workload'], d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
    print(d['parity'])
    for r in d['rooflines'][:10]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
