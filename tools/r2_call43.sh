#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print(\"SMOKE_OK\")" > gpurun_out/r2c43_smoke.log 2>&1; tail -3 gpurun_out/r2c43_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c43_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c43_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2c43_bench_256_full.json 2> gpurun_out/r2c43_bench_256_full.err
tail -3 gpurun_out/r2c43_bench_256_full.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c43_bench_256_full.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
print(d['roofline']['kernel'], d['roofline']['frac'], d.get('ptap_fused',{}).get('frac'), d.get('ptap',{}).get('frac'))
for r in d['rooflines'][:5]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c43_launches_256.csv \
    python bench.py --steps 1 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c43_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c43_launches_256.csv "python bench.py --steps 1 --warmup 1 --no-ptap --no-cpu (4 steps in the list: warm-up, timed, e2e warm-up, e2e)" > gpurun_out/r2c43_launches_256_summary.txt
head -24 gpurun_out/r2c43_launches_256_summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c43_reference_arm.json 2> gpurun_out/r2c43_reference_arm.err
tail -c 700 gpurun_out/r2c43_reference_arm.json
