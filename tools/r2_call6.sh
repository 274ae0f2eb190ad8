#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c6_gpu_tests.log 2>&1
tail -8 gpurun_out/r2c6_gpu_tests.log
timeout 600 python bench.py --no-ptap --no-cpu > gpurun_out/r2c6_bench_256.json 2> gpurun_out/r2c6_bench_256.err
tail -3 gpurun_out/r2c6_bench_256.err
timeout 600 python bench.py --mode matfree --nel 128 --steps 2 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c6_matfree_128.json 2> gpurun_out/r2c6_matfree_128.err
tail -3 gpurun_out/r2c6_matfree_128.err
timeout 900 python bench.py --mode matfree --nel 512 --steps 1 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c6_matfree_512.json 2> gpurun_out/r2c6_matfree_512.err
tail -3 gpurun_out/r2c6_matfree_512.err
python - <<'P'
import json
for f in ['r2c6_bench_256','r2c6_matfree_128','r2c6_matfree_512']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
    print(d['parity'])
    for r in d['rooflines']: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gsf|tigar_qp" -c 7 -o gpurun_out/r2c6_asm256 \
    python bench.py --steps 1 --warmup 0 --no-ptap --no-cpu > gpurun_out/r2c6_ncu_asm.log 2>&1
ls -la gpurun_out/r2c6_asm256.ncu-rep
