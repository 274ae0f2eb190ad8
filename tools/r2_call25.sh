#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c25_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c25_gpu_tests.log
timeout 900 python bench.py --no-cpu > gpurun_out/r2c25_bench_256.json 2> gpurun_out/r2c25_bench_256.err
tail -3 gpurun_out/r2c25_bench_256.err
timeout 900 python bench.py --mode matfree --nel 512 --steps 1 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c25_matfree_512.json 2> gpurun_out/r2c25_matfree_512.err
tail -3 gpurun_out/r2c25_matfree_512.err
python - <<'P'
import json
for f in ['r2c25_bench_256','r2c25_matfree_512']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['config']['workload'], d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
    print(d['parity'])
    print(d['roofline']['kernel'], d['roofline']['frac'], d.get('ptap_fused',{}).get('frac'), d.get('ptap',{}).get('frac'))
    for r in d['rooflines'][:6]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
