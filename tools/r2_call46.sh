#!/bin/bash
# same-box A/B: library of the working tree vs a variant library (TIGAR_B200_LIB)
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_gsf.py -m gpu -q -x 2>&1 | tail -2
for v in new prev new prev; do
if [ $v = new ]; then unset TIGAR_B200_LIB; else export TIGAR_B200_LIB=$PWD/gpurun_variants/libtigar_$v.so; fi
timeout 600 python bench.py --no-ptap --no-cpu --steps 5 > gpurun_out/r2c40_bench_$v.json 2> gpurun_out/r2c40_bench_$v.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c40_bench_$v.json').read().strip().splitlines()[-1])
print("$v", d['ms_per_step'], d['stage_ms'])
for r in d['rooflines'][:4]: print("  %-50s %8.2f ms/step  hbm %.3f" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac']))
P
done
